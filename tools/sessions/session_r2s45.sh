O=gpurun_out/r2s45; mkdir -p $O
BLR_VAR_SMALL_MAX=128 timeout 600 python -m pytest tests/test_gpu_boundaries.py tests/test_gpu_unaligned.py -m gpu -q --timeout 300 -x -k "prediction or marginals or unaligned_inference" > $O/tests.log 2>&1; echo tests exit $?; tail -3 $O/tests.log
for m in 128 64; do BLR_VAR_SMALL_MAX=$m BLR_BENCH_DS=66,72,80,96,112,128 timeout 300 python tools/bench_small_d.py 2>> $O/mid.err | grep mean_and_var | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('VAR_SMALL_MAX=$m', d['config'], 'ms %.3f'%d['ms'], 'TF %.1f'%d['tflops_triangular'])"; done

O=gpurun_out/r2s44; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_unaligned.py tests/test_gpu_boundaries.py -m gpu -q --timeout 300 -x > $O/tests.log 2>&1; echo tests exit $?; tail -4 $O/tests.log
for m in 1 0; do BLR_MID_RING=$m BLR_BENCH_DS=67,81,95 timeout 300 python tools/bench_small_d.py 2>> $O/mid.err | grep posterior | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('MID_RING=$m', d['config'], 'ms %.3f'%d['ms'])"; done
BLR_SANITIZE_SET=case:9 timeout 300 compute-sanitizer --tool memcheck python tests/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|max rel"

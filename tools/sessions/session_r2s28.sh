O=gpurun_out/r2s28; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_unaligned.py tests/test_gpu_boundaries.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_illcond.py -m gpu -q --timeout 300 -x > $O/tests.log 2>&1; echo tests exit $?; tail -15 $O/tests.log
BLR_BENCH_DS=66,72,80,96,112,120,127,128 timeout 300 python tools/bench_small_d.py > $O/mid_d.jsonl 2> $O/mid_d.err; echo mid_d exit $?; tail -3 $O/mid_d.err
python - <<'PY'
import json
for l in open("gpurun_out/r2s28/mid_d.jsonl"):
    d=json.loads(l); print(d["config"], "ms %.3f"%d["ms"], "hbm %.2f"%d["frac_of_measured_hbm"], "TF %.1f"%d.get("tflops_triangular", d.get("gram_tflops",0)), "gram_ms %.3f"%d.get("gram_ms",0))
PY

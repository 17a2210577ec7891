O=gpurun_out/r2s26; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_boundaries.py tests/test_gpu_parity.py -m gpu -q --timeout 300 > $O/tests.log 2>&1; echo tests exit $?; tail -3 $O/tests.log
timeout 300 python tools/bench_small_d.py > $O/small_d.jsonl 2> $O/small_d.err; echo small_d exit $?
BLR_BENCH_DS=66,72,80,96,100,112,120,127,128,192 timeout 300 python tools/bench_small_d.py > $O/mid_d.jsonl 2> $O/mid_d.err; echo mid_d exit $?
for f in direct whitened; do for shape in "256 1048576" "1024 262144" "64 1048576"; do BLR_FORM=$f timeout 120 python tools/step_breakdown.py $shape 20 >> $O/form_$f.jsonl 2>> $O/form.err; done; done
python - <<'PY'
import json
for f in ("small_d","mid_d"):
    for l in open(f"gpurun_out/r2s26/{f}.jsonl"):
        d=json.loads(l); print(d["config"], "ms %.3f"%d["ms"], "hbm %.2f"%d["frac_of_measured_hbm"], "TF %.1f"%d.get("tflops_triangular", d.get("gram_tflops",0)), "gram_ms %.3f"%d.get("gram_ms",0))
for f in ("direct","whitened"):
    for l in open(f"gpurun_out/r2s26/form_{f}.jsonl"):
        d=json.loads(l); print(f, d["D"], d["N"], "api %.3f"%d["api_posterior_and_logpdf_ms"], d["device_phases_ms"])
PY

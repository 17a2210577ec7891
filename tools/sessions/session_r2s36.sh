O=gpurun_out/r2s36; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_boundaries.py -m gpu -q --timeout 300 -x -k "mid_ring or seams" > $O/tests.log 2>&1; echo tests exit $?; tail -8 $O/tests.log
for m in 1 0; do BLR_MID_RING=$m BLR_BENCH_DS=66,72,80,88,96 timeout 300 python tools/bench_small_d.py > $O/mid_ring$m.jsonl 2>> $O/mid.err; done
for m in 1; do BLR_BENCH_MW=1 BLR_MID_RING=$m BLR_BENCH_DS=72,96 timeout 300 python tools/bench_small_d.py > $O/mid_ring${m}_mw.jsonl 2>> $O/mid.err; done
python - <<'PY'
import json
for f in ("mid_ring1","mid_ring0","mid_ring1_mw"):
    for l in open(f"gpurun_out/r2s36/{f}.jsonl"):
        d=json.loads(l)
        if "posterior" in d["config"]: print(f, d["config"], "ms %.3f"%d["ms"], "gram_ms %.3f"%d.get("gram_ms",0), "TF %.1f"%d.get("gram_tflops",0))
PY
tail -3 $O/mid.err

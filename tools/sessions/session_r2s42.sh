O=gpurun_out/r2s42; mkdir -p $O
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench20.json 2> $O/bench20.err; echo bench20 exit $?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s42/bench20.json"))
print("value %.5g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], "cpu %.4g"%d["cpu_baseline"]["value"], d["clocks"], d["parity"]["ok"])
PY
BLR_BENCH_DS=96 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_mid_ring -s 2 -c 1 -o $O/prof_gram_mid -f python tools/bench_small_d.py > $O/ncu_mid.log 2>&1; echo ncu_mid exit $?
BLR_BENCH_DS=96 timeout 600 ncu --set full --clock-control none --import-source on -k regex:var_tma_kernel -s 2 -c 1 -o $O/prof_var_mid -f python tools/bench_small_d.py > $O/ncu_var.log 2>&1; echo ncu_var exit $?
ls -la $O

#!/bin/bash
# One gpurun call = one session: run the requested stages, never abort on a failed stage, log under gpurun_out/.
# usage: tools/gpu_session.sh <tag> stage [stage ...]      stages: tests smoke calib bench_small bench ncu_list ncu_full
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv > "$OUT/gpu.csv" 2>&1
nproc > "$OUT/nproc.txt"; free -g >> "$OUT/nproc.txt"
for stage in "$@"; do
  echo "=== stage $stage $(date +%T)"
  case $stage in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > "$OUT/tests.log" 2>&1; echo "tests exit $?"; tail -15 "$OUT/tests.log";;
    tests_all)
      timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/tests.log" 2>&1; echo "tests exit $?"; tail -40 "$OUT/tests.log";;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -5 "$OUT/smoke.log";;
    calib)
      timeout 300 python -c "
import blr_b200 as b, json
c = b.Context(0)
print(json.dumps(c.calibrate()))
" > "$OUT/calib.log" 2>&1; echo "calib exit $?"; tail -3 "$OUT/calib.log";;
    bench_small)
      timeout 600 python bench.py --n-obs 1048576 --dim 256 --steps 5 --warmup 3 --e2e-obs 262144 --cpu-sample 65536 > "$OUT/bench_small.json" 2> "$OUT/bench_small.err"; echo "bench_small exit $?"; tail -c 3000 "$OUT/bench_small.json"; tail -5 "$OUT/bench_small.err";;
    bench)
      timeout 1200 python bench.py --steps 3 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"; tail -c 3000 "$OUT/bench.json"; tail -5 "$OUT/bench.err";;
    bench_ref)
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench_ref exit $?"; tail -c 1500 "$OUT/bench_ref.json";;
    ncu_list)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
        python bench.py --n-obs 2097152 --steps 1 --warmup 1 --no-cpu --no-calibrate --e2e-obs 65536 > "$OUT/ncu_list.log" 2>&1; echo "ncu_list exit $?";;
    ncu_full)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gram_tma -s 1 -c 1 -o "$OUT/prof_gram" -f \
        python bench.py --n-obs 2097152 --steps 1 --warmup 1 --no-cpu --no-calibrate --e2e-obs 65536 > "$OUT/ncu_full.log" 2>&1; echo "ncu_full exit $?";;
    multi_tests)
      timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 600 > "$OUT/multi_tests.log" 2>&1; echo "multi_tests exit $?"; tail -15 "$OUT/multi_tests.log";;
    bench_multi)
      NG=$(nvidia-smi -L | wc -l)
      for n in ${NLIST:-1 2 4 8}; do
        if [ $n -le $NG ]; then
          if [ $n -eq 1 ]; then
            timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-e2e > "$OUT/bench_g$n.json" 2> "$OUT/bench_g$n.err"
          else
            timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
              bench.py --gpus $n --steps 3 --warmup 3 --no-cpu --no-e2e > "$OUT/bench_g$n.json" 2> "$OUT/bench_g$n.err"
          fi
          echo "bench_multi n=$n exit $?"; tail -c 1200 "$OUT/bench_g$n.json"; tail -3 "$OUT/bench_g$n.err"
        fi
      done;;
    solve_timing)
      for cfg in "--n-obs 1048576 --dim 256" "--n-obs 1048576 --dim 1024" "--n-obs 262144 --dim 4096"; do
        echo "cfg=$cfg"
        timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/solve.err" \
          | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'gram', d['roofline']['kernel_ms'], 'solve', d['roofline']['solve_ms'], 'TF', d['roofline']['achieved'])"
      done 2>&1 | tee "$OUT/solve_timing.log";;
    configs)
      timeout 1200 python tools/bench_configs.py > "$OUT/configs.jsonl" 2> "$OUT/configs.err"; echo "configs exit $?"; cat "$OUT/configs.jsonl"; tail -5 "$OUT/configs.err";;
    ncu_pred)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"var_tma|rand_tma" -s 2 -c 2 -o "$OUT/prof_pred" -f \
        python tools/run_predict.py 512 2097152 64 > "$OUT/ncu_pred.log" 2>&1; echo "ncu_pred exit $?"; tail -3 "$OUT/ncu_pred.log";;
    kt_sweep)
      for cfg in "--n-obs 1048576 --dim 256" "--n-obs 2097152 --dim 1024"; do
        echo "KT=16 STAGES=6 cfg=$cfg"
        BLR_GRAM_KT=16 BLR_GRAM_STAGES=6 timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/kt.err" \
          | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
      done
      for kt in 16 32; do
        for cfg in "--n-obs 1048576 --dim 256" "--n-obs 2097152 --dim 1024"; do
          echo "KT=$kt cfg=$cfg"
          BLR_GRAM_KT=$kt timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/kt.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
        done
      done 2>&1 | tee "$OUT/kt_sweep.log"
      BLR_GRAM_KT=32 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "statistics or posterior_and_logpdf" > "$OUT/tests_kt32.log" 2>&1; tail -3 "$OUT/tests_kt32.log";;
    period_sweep)
      for po in ${PERIODS:-0 2048 4096 8192 1000000000}; do
        for cfg in "--n-obs 1048576 --dim 256" "--n-obs 2097152 --dim 1024"; do
          echo "PERIOD_OBS=$po cfg=$cfg"
          BLR_GRAM_PERIOD_OBS=$po timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/period.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['clocks'].get('power_w_max'))"
        done
      done 2>&1 | tee "$OUT/period_sweep.log";;
    tests_fast)
      timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --ignore tests/test_gpu_fullsize.py > "$OUT/tests.log" 2>&1; echo "tests exit $?"; tail -15 "$OUT/tests.log";;
    fullsize)
      timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s --timeout 900 --durations=5 > "$OUT/fullsize.log" 2>&1; echo "fullsize exit $?"; tail -25 "$OUT/fullsize.log";;
    cfg4_multi)
      NG=$(nvidia-smi -L | wc -l)
      if [ $NG -gt 1 ]; then
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 \
          tools/bench_cfg4.py > "$OUT/cfg4_g$NG.json" 2> "$OUT/cfg4_g$NG.err"
      else
        timeout 900 python tools/bench_cfg4.py > "$OUT/cfg4_g$NG.json" 2> "$OUT/cfg4_g$NG.err"
      fi
      echo "cfg4 exit $?"; cat "$OUT/cfg4_g$NG.json"; tail -3 "$OUT/cfg4_g$NG.err";;
    var_ab)
      for c in 0 1; do
        echo "BLR_VAR_CFG=$c"
        BLR_VAR_CFG=$c timeout 600 python tools/bench_cfg4.py --max-log2-per-gpu 23 --n-fit 262144 2>> "$OUT/var_ab.err" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mean_var', d['mean_and_var'], 'rand', d['rand']['ms'])"
        BLR_VAR_CFG=$c timeout 900 python -m pytest tests -m gpu -q -k "mean_var or golden or cfg4 or reference_suite" --timeout 600 2>&1 | tail -2
      done 2>&1 | tee "$OUT/var_ab.log";;
    var_dbg)
      for c in "0 0" "0 1" "1 0" "1 1" "0 0" "0 1"; do
        set -- $c
        echo "BLR_VAR_CFG=$1 BLR_VAR_DBG=$2"
        BLR_VAR_CFG=$1 BLR_VAR_DBG=$2 timeout 600 python tools/bench_cfg4.py --max-log2-per-gpu 23 --n-fit 262144 2>> "$OUT/var_dbg.err" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mean_var ms', d['mean_and_var']['ms'], 'TF', d['mean_and_var']['tflops_per_gpu'])"
      done 2>&1 | tee "$OUT/var_dbg.log";;
    ncu_solve)
      for d in 256 1024; do
        timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file "$OUT/solve_launches_D$d.csv" \
          python bench.py --n-obs 262144 --dim $d --steps 1 --warmup 1 --no-cpu --no-calibrate --no-e2e > "$OUT/ncu_solve_D$d.log" 2>&1; echo "ncu_solve D=$d exit $?"
      done;;
    probe_mix)
      for v in 0 1 2 3 0 1; do
        BLR_PROBE_VARIANT=$v timeout 300 python -c "
import blr_b200 as b, ctypes as C
c = b.Context(0)
v = C.c_double(); c.check(c.lib.blr_calibrate_gram_inner(c.handle, C.byref(v)))
p = C.c_double(); c.check(c.lib.blr_calibrate_dmma(c.handle, C.byref(p)))
print('variant $v mix TF', round(v.value, 3), 'pure DMMA TF', round(p.value, 3), 'frac', round(v.value / p.value, 4))
"
      done 2>&1 | tee "$OUT/probe_mix.log";;
    cs_ab)
      timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "consumer_tilings" --timeout 600 2>&1 | tail -5
      for c in 0 1 0 1; do
        for cfg in "--n-obs 1048576 --dim 256" "--n-obs 2097152 --dim 1024" "--n-obs 262144 --dim 4096"; do
          echo "BLR_GRAM_CS=$c cfg=$cfg"
          BLR_GRAM_CS=$c timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/cs.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['clocks'].get('power_w_max'))"
        done
      done 2>&1 | tee "$OUT/cs_ab.log";;
    cs_weight)
      for w in ${WEIGHTS:-40 41 42 43}; do
        for cfg in "--n-obs 1048576 --dim 256" "--n-obs 2097152 --dim 1024" "--n-obs 1048576 --dim 2048"; do
          echo "BLR_GRAM_CS=1 BLR_DIAG_WEIGHT=$w cfg=$cfg"
          BLR_GRAM_CS=1 BLR_DIAG_WEIGHT=$w timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/csw.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
        done
      done 2>&1 | tee "$OUT/cs_weight.log"
      for cfg in "--n-obs 1048576 --dim 2048"; do
        echo "BLR_GRAM_CS=0 cfg=$cfg"
        BLR_GRAM_CS=0 timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/csw.err" \
          | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
      done 2>&1 | tee -a "$OUT/cs_weight.log";;
    diag_probe)
      for c in 0 1; do
        echo "BLR_GRAM_CS=$c D=128"
        BLR_GRAM_CS=$c timeout 600 python bench.py --n-obs 4194304 --dim 128 --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/dp.err" \
          | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
      done 2>&1 | tee "$OUT/diag_probe.log"
      BLR_GRAM_CS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_tma -s 1 -c 1 -o "$OUT/prof_diag" -f \
        python bench.py --n-obs 1048576 --dim 128 --steps 1 --warmup 1 --no-cpu --no-calibrate --no-e2e > "$OUT/ncu_diag.log" 2>&1; echo "ncu exit $?";;
    l2pf_sweep)
      for pf in ${PFS:-0 2 4 8}; do
        for w in ${WEIGHTS:-36 40}; do
          for cfg in "--n-obs 4194304 --dim 128" "--n-obs 1048576 --dim 256" "--n-obs 2097152 --dim 1024"; do
            echo "L2PF=$pf DIAG_WEIGHT=$w cfg=$cfg"
            BLR_GRAM_L2PF=$pf BLR_DIAG_WEIGHT=$w timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/pf.err" \
              | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
          done
        done
      done 2>&1 | tee "$OUT/l2pf_sweep.log";;
    shape_sweep)
      for cfg in "--n-obs 4194304 --dim 128" "--n-obs 1048576 --dim 256" "--n-obs 1048576 --dim 384" "--n-obs 1048576 --dim 512" "--n-obs 1048576 --dim 640" "--n-obs 1048576 --dim 768" "--n-obs 2097152 --dim 1024" "--n-obs 1048576 --dim 2048" "--n-obs 262144 --dim 4096"; do
        for c in ${CSS:-0 1}; do
          echo "BLR_GRAM_CS=$c cfg=$cfg"
          BLR_GRAM_CS=$c timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/shape.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
        done
      done 2>&1 | tee "$OUT/shape_sweep.log";;
    unit_ab)
      timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py -m gpu -q --timeout 600 2>&1 | tail -3
      for u in 1 0; do
        for cfg in "--n-obs 1048576 --dim 256" "--n-obs 2097152 --dim 1024"; do
          echo "BLR_GRAM_UNIT=$u scalar-noise cfg=$cfg"
          BLR_GRAM_UNIT=$u timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate --scalar-noise 2>> "$OUT/unit.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
        done
      done 2>&1 | tee "$OUT/unit_ab.log";;
    small_d)
      timeout 600 python tools/bench_small_d.py > "$OUT/small_d.jsonl" 2> "$OUT/small_d.err"; echo "small_d exit $?"; cat "$OUT/small_d.jsonl"; tail -3 "$OUT/small_d.err";;
    rff_multi)
      NG=$(nvidia-smi -L | wc -l)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
        tools/bench_rff.py > "$OUT/rff_g$NG.json" 2> "$OUT/rff_g$NG.err"; echo "rff exit $?"; cat "$OUT/rff_g$NG.json"; tail -3 "$OUT/rff_g$NG.err";;
    dmma_probe)
      timeout 300 python -c "
import blr_b200 as b, ctypes as C
c = b.Context(0)
for w in (1,2,3,4,8,12,16):
    row=[]
    for na in (1,2,4,8,16,32):
        v=C.c_double(); c.check(c.lib.blr_calibrate_dmma_cfg(c.handle, w, na, C.byref(v))); row.append(round(v.value,2))
    print('warps/SM', w, 'n_acc 1,2,4,8,16,32 ->', row)
" > "$OUT/dmma_probe.log" 2>&1; echo "dmma_probe exit $?"; cat "$OUT/dmma_probe.log";;
    sweep_diag)
      for w in 36 38 40 42; do
        for cfg in "--n-obs 1048576 --dim 256" "--n-obs 4194304 --dim 1024"; do
          echo "diag_weight=$w cfg=$cfg"
          BLR_DIAG_WEIGHT=$w timeout 600 python bench.py $cfg --steps 5 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/sweep.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'])"
        done
      done 2>&1 | tee "$OUT/sweep_diag.log";;
    tests_new)
      timeout 2400 python -m pytest tests/test_gpu_contracts.py tests/test_gpu_shapes.py tests/test_gpu_illcond.py tests/test_highprec_pins.py -m gpu -q -s --timeout 900 > "$OUT/tests_new.log" 2>&1; echo "tests_new exit $?"; grep -E "^\[|passed|failed|Error|error" "$OUT/tests_new.log" | tail -80;;
    tests_contracts)
      timeout 1200 python -m pytest tests/test_gpu_contracts.py -m gpu -q -s -x --timeout 600 > "$OUT/tests_contracts.log" 2>&1; echo "tests_contracts exit $?"; tail -30 "$OUT/tests_contracts.log";;
    sanitize)
      for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
        BLR_SANITIZE_SET=${SAN_SET:-all} timeout 1500 compute-sanitizer --tool $tool --print-limit ${PRINT_LIMIT:-20} python tests/sanitize_small.py > "$OUT/sanitize_$tool.log" 2>&1
        echo "sanitize $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max rel err|Error|hazard" "$OUT/sanitize_$tool.log" | tail -20
      done;;
    racecheck_cases)
      # one racecheck run per TMA case with an unbounded report, reduced to the distinct (write site, read site) pairs
      for c in ${RC_CASES:-3 4 5 6 7}; do
        BLR_SANITIZE_SET=case:$c timeout 900 compute-sanitizer --tool racecheck --print-limit 100000 python tests/sanitize_small.py > "$OUT/racecheck_case$c.log" 2>&1
        echo "racecheck case $c exit $?: $(grep -E 'RACECHECK SUMMARY|max rel err' "$OUT/racecheck_case$c.log" | tr '\n' ' ')"
        grep -E "Race reported|and (Read|Write) access" "$OUT/racecheck_case$c.log" | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//; s/^=+ +//' | sort | uniq -c | sort -rn > "$OUT/racecheck_case${c}_sites.txt"
        head -12 "$OUT/racecheck_case${c}_sites.txt"
        gzip -f "$OUT/racecheck_case$c.log"
      done;;
    racecheck_strict)
      # the same cases against the strict-arrive build (every lane arrives on the ring mbarriers itself; csrc/Makefile `strict`)
      for c in ${RC_CASES:-4 7 8}; do
        LIBBLR_CUDA=$PWD/bayesianlinearregressors.jl_b200/csrc/libblr_cuda_strict.so BLR_SANITIZE_SET=case:$c timeout 900 \
          compute-sanitizer --tool racecheck --print-limit 100000 python tests/sanitize_small.py > "$OUT/racecheck_strict_case$c.log" 2>&1
        echo "racecheck STRICT case $c exit $?: $(grep -E 'RACECHECK SUMMARY|max rel err' "$OUT/racecheck_strict_case$c.log" | tr '\n' ' ')"
        grep -E "Race reported|and (Read|Write) access" "$OUT/racecheck_strict_case$c.log" | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//; s/^=+ +//' | sort | uniq -c | sort -rn > "$OUT/racecheck_strict_case${c}_sites.txt"
        head -6 "$OUT/racecheck_strict_case${c}_sites.txt"
        gzip -f "$OUT/racecheck_strict_case$c.log"
      done;;
    breakdown)
      for cfg in "256 1048576" "1024 1048576" "64 4194304"; do
        timeout 300 python tools/step_breakdown.py $cfg 50 2>> "$OUT/breakdown.err" | tee -a "$OUT/breakdown.jsonl"
      done;;
    traffic)
      timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gram_tma --csv \
        --log-file "$OUT/traffic.csv" python tools/gram_traffic.py > "$OUT/traffic.log" 2>&1; echo "traffic exit $?"; tail -8 "$OUT/traffic.log"; grep -c gram_tma "$OUT/traffic.csv";;
    dxd_ab)
      for mode in fused legacy; do
        for cfg in "--n-obs 1048576 --dim 256" "--n-obs 1048576 --dim 1024" "--n-obs 262144 --dim 2048" "--n-obs 262144 --dim 4096"; do
          echo "BLR_DXD=$mode cfg=$cfg"
          BLR_DXD=$mode timeout 600 python bench.py $cfg --steps 10 --warmup 3 --no-cpu --no-e2e --no-calibrate 2>> "$OUT/dxd.err" \
            | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'gram', d['roofline']['kernel_ms'], 'solve', d['roofline']['solve_ms'], 'TF', d['roofline']['achieved'], 'logpdf', d['logpdf'])"
        done
      done 2>&1 | tee "$OUT/dxd_ab.log";;
    ncu_dxd)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:dxd_fused -s 2 -c 1 -o "$OUT/prof_dxd" -f \
        python bench.py --n-obs 262144 --dim 1024 --steps 1 --warmup 2 --no-cpu --no-calibrate --no-e2e > "$OUT/ncu_dxd.log" 2>&1; echo "ncu_dxd exit $?";;
    dxd_trace)
      for cfg in "--n-obs 65536 --dim 256" "--n-obs 65536 --dim 1024" "--n-obs 32768 --dim 4096"; do
        d=$(echo $cfg | awk '{print $4}')
        BLR_DXD_TRACE="$OUT/dxd_trace_D$d.txt" timeout 300 python bench.py $cfg --steps 2 --warmup 1 --no-cpu --no-e2e --no-calibrate > /dev/null 2>> "$OUT/trace.err"
        echo "trace D=$d lines $(wc -l < $OUT/dxd_trace_D$d.txt)"
      done;;
    bench_cfgs)
      for cfgname in ${CFGS:-cfg2 cfg4 cfg5}; do
        timeout 1500 python bench.py --config $cfgname --steps ${STEPS:-5} --warmup 3 > "$OUT/bench_$cfgname.json" 2> "$OUT/bench_$cfgname.err"; echo "bench $cfgname exit $?"
        tail -c 2500 "$OUT/bench_$cfgname.json"; tail -3 "$OUT/bench_$cfgname.err"
      done;;
    bench_pm)
      timeout 1500 python bench.py --prior-mean random --steps 5 --warmup 3 --no-cpu --no-e2e > "$OUT/bench_cfg3_pm.json" 2> "$OUT/bench_cfg3_pm.err"; echo "bench_pm exit $?"; tail -c 1500 "$OUT/bench_cfg3_pm.json";;
    sanitize_late)
      # kernels added late in round 2: trsm_lower_kernel / dxd_whitened, repack_colvecs + padded-odd Gram, staged var / rand, Cfg<2,128>
      for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
        BLR_SANITIZE_SET=late timeout 1200 compute-sanitizer --tool $tool --print-limit 50 python tests/sanitize_small.py > "$OUT/sanitize_late_$tool.log" 2>&1
        echo "sanitize_late $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max rel err|late set done" "$OUT/sanitize_late_$tool.log" | tail -12
      done;;
    bench20)
      timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench20.json" 2> "$OUT/bench20.err"; echo "bench20 exit $?"; tail -c 3500 "$OUT/bench20.json"; tail -5 "$OUT/bench20.err";;
    *) echo "unknown stage $stage";;
  esac
done
echo "=== done $(date +%T)"

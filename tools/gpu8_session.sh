set -u
OUT=gpurun_out/r2s09; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt; numactl -H >> $OUT/nproc.txt 2>&1
run() { # name, args...
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@" > $OUT/$name.json 2> $OUT/$name.err
  echo "$name exit $?"; tail -c 1800 $OUT/$name.json; echo; tail -2 $OUT/$name.err
}
run bench_cfg3_g8 --steps 10 --warmup 3 --no-cpu
run bench_cfg4_g8 --config cfg4 --steps 5 --warmup 3 --no-cpu
run bench_cfg5_g8 --config cfg5 --steps 5 --warmup 3 --no-cpu
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 > $OUT/multi_tests.log 2>&1; tail -3 $OUT/multi_tests.log

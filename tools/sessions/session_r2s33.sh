O=gpurun_out/r2s33; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_shapes.py -m gpu -q --timeout 500 -k "maximum_dimension" -s > $O/maxdim.log 2>&1; echo maxdim exit $?; tail -5 $O/maxdim.log
bash tools/gpu_session.sh r2s33 bench_cfgs ncu_list

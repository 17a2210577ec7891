O=gpurun_out/r2s29; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_unaligned.py tests/test_gpu_boundaries.py tests/test_gpu_parity.py -m gpu -q --timeout 300 > $O/tests.log 2>&1; echo tests exit $?; tail -15 $O/tests.log

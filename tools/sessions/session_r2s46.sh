O=gpurun_out/r2s46; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --ignore tests/test_gpu_fullsize.py > $O/tests.log 2>&1; echo tests exit $?; tail -4 $O/tests.log
for c in 13 5; do for tool in memcheck racecheck; do
  BLR_SANITIZE_SET=case:$c timeout 600 compute-sanitizer --tool $tool --print-limit 10 python tests/sanitize_small.py 2>&1 | grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|max rel err' | tr '\n' ' '; echo " [case $c $tool]"
done; done

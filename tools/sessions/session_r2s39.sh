O=gpurun_out/r2s39; mkdir -p $O
for c in 13 14; do for tool in racecheck synccheck; do
  BLR_SANITIZE_SET=case:$c timeout 600 compute-sanitizer --tool $tool --print-limit 30 python tests/sanitize_small.py > $O/san_case${c}_$tool.log 2>&1
  echo "case $c $tool exit $?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|max rel err' $O/san_case${c}_$tool.log | tr '\n' ' ')"
done; done
grep -h -E "Race reported|and (Read|Write) access" $O/san_case1*_racecheck.log | sed -E 's/\+0x[0-9a-f]+//g; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -8
timeout 600 python -m pytest tests/test_gpu_boundaries.py -m gpu -q --timeout 300 -x -k "mid_ring" 2>&1 | tail -2

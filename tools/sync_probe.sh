mkdir -p gpurun_out/r2s20
run() { # label, env...
  local label=$1; shift
  env "$@" BLR_SANITIZE_SET=case:4 timeout 120 compute-sanitizer --tool synccheck --print-limit 2000 python tests/sanitize_small.py > gpurun_out/r2s20/sync_$label.log 2>&1
  echo "== $label: $(grep -c 'Barrier error' gpurun_out/r2s20/sync_$label.log) barrier errors; $(grep -E 'ERROR SUMMARY|max rel err|launch failure' gpurun_out/r2s20/sync_$label.log | tr '\n' ' ' | cut -c1-260)"
  grep -E "Barrier is located" gpurun_out/r2s20/sync_$label.log | sort | uniq -c | head -6
  grep -E "by thread" gpurun_out/r2s20/sync_$label.log | sed -E "s/.*thread \(([0-9]+),0,0\) in block \(([0-9]+).*/\1 \2/" | awk '{print "warp", int($1/32), "block", $2}' | sort | uniq -c | head -8
}
run product
run inline_emit LIBBLR_CUDA=$PWD/bayesianlinearregressors.jl_b200/csrc/libblr_cuda_inl.so
run single_group BLR_RAND_PP=0

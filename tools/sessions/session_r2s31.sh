O=gpurun_out/r2s32; mkdir -p $O
C=bayesianlinearregressors.jl_b200/csrc
for v in "" _rff32x1 _rff16x2 _rff8x2; do
  LIBBLR_CUDA=$PWD/$C/libblr_cuda$v.so timeout 200 python tools/bench_features.py >> $O/features.jsonl 2>> $O/features.err
done
cat $O/features.jsonl; tail -3 $O/features.err


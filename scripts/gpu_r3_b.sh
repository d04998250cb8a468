#!/bin/bash
# conv layers with 8 epilogue warps: parity, A/B timings
out=gpurun_out/${1:-t2}
mkdir -p $out
( timeout 500 python -m pytest tests/test_gpu_encoder.py -m gpu -q --tb=short 2>&1 | tail -30 ) > $out/pytest_enc.log
{
for lib in scripts/exp/libnaf_nepi128.so "" scripts/exp/libnaf_nepi128.so ""; do
  export NAF_B200_LIB=$lib
  [ -z "$lib" ] && unset NAF_B200_LIB
  echo "== lib: ${lib:-regular}"
  timeout 120 python scripts/enc_bench.py 8 448 448 1
done
} > $out/enc_bench.log 2>&1
tail -5 $out/pytest_enc.log; cat $out/enc_bench.log

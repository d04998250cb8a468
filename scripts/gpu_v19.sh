#!/bin/bash
# elect.sync in the conv encoder kernels: parity + A/B against the lane == 0 build
out=gpurun_out/${1:-v19}
mkdir -p $out
( timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q --tb=short -x 2>&1 | tail -8 ) > $out/pytest.log
tail -3 $out/pytest.log
{
for lib in "" scripts/exp/libnaf_convlane0.so "" scripts/exp/libnaf_convlane0.so; do
  echo "== lib: ${lib:-regular (elect)}"
  NAF_B200_LIB=$lib timeout 120 python scripts/enc_bench.py 8 448 448 1
  NAF_B200_LIB=$lib timeout 120 python scripts/enc_bench.py 8 448 448 3
done
} > $out/enc_bench.log 2>&1
cat $out/enc_bench.log

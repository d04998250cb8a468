#!/bin/bash
# elect.sync in the forward TMA kernel: parity + A/B against the lane == 0 build
out=gpurun_out/${1:-v18}
mkdir -p $out
( timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tma.py -m gpu -q --tb=short -x 2>&1 | tail -30 ) > $out/pytest.log
tail -4 $out/pytest.log
{
for lib in "" scripts/exp/libnaf_tmalane0.so ""  scripts/exp/libnaf_tmalane0.so; do
  echo "== lib: ${lib:-regular (elect)}"
  NAF_B200_LIB=$lib timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  NAF_B200_LIB=$lib timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 1
  NAF_B200_LIB=$lib timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
done
} > $out/time_xattn.log 2>&1
cat $out/time_xattn.log

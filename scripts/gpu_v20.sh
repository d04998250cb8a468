#!/bin/bash
# forward TMA kernel: next tile's Q K^T issued after the first accumulator half (NAF_TMA_QK_MID=1) vs before the P V
out=gpurun_out/${1:-v20}
mkdir -p $out
{
for lib in "" scripts/exp/libnaf_tmaqkmid.so "" scripts/exp/libnaf_tmaqkmid.so; do
  echo "== lib: ${lib:-regular}"
  NAF_B200_LIB=$lib timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
  NAF_B200_LIB=$lib timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 1
done
} > $out/time_xattn.log 2>&1
cat $out/time_xattn.log
( NAF_B200_LIB=scripts/exp/libnaf_tmaqkmid.so timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tma.py -m gpu -q --tb=short -x 2>&1 | tail -5 ) > $out/pytest_qkmid.log
tail -3 $out/pytest_qkmid.log

#!/bin/bash
# C3 / C2 / C5 attention launch with the profiling variants of the TMA kernel
tag=${1:-exp}
out=gpurun_out/$tag
mkdir -p $out
{
for lib in "" scripts/exp/libnaf_nostore.so scripts/exp/libnaf_noq.so scripts/exp/libnaf_noexp.so scripts/exp/libnaf_nothing.so; do
  export NAF_B200_LIB=$lib
  [ -z "$lib" ] && unset NAF_B200_LIB
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 4 cell_tma 768 2048 32 7 4
done
unset NAF_B200_LIB
timeout 120 python scripts/time_xattn.py 1 cell_tma 384 224 16 7 1
timeout 120 python scripts/time_xattn.py 2 cell_tma 768 1344 24 7 4
} > $out/time_xattn.log 2>&1
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -30 ) > $out/pytest.log
cat $out/time_xattn.log; tail -3 $out/pytest.log

#!/bin/bash
# TMA kernel with two accumulator buffers (NOB = 2): parity, then A/B timings against NOB = 1 and profiling variants
tag=${1:-t3}
out=gpurun_out/$tag
mkdir -p $out
{
timeout 120 python scripts/check_xattn.py cell_tma 2 768 224 8 7
timeout 120 python scripts/check_xattn.py cell_tma 1 1024 336 12 11
timeout 120 python scripts/check_xattn.py cell_tma 1 1024 252 9 9
timeout 120 python scripts/check_xattn.py cell_tma 2 1024 192 12 5
} > $out/check.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -30 ) > $out/pytest.log
{
for lib in scripts/exp/libnaf_nob1.so "" scripts/exp/libnaf_nob2_nostore.so scripts/exp/libnaf_nob2_nothing.so scripts/exp/libnaf_nob1_nothing.so scripts/exp/libnaf_nob1.so ""; do
  export NAF_B200_LIB=$lib
  [ -z "$lib" ] && unset NAF_B200_LIB
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 1
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
done
} > $out/time_xattn.log 2>&1
cat $out/check.log; tail -3 $out/pytest.log; cat $out/time_xattn.log

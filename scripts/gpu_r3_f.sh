#!/bin/bash
# L2 cache-policy hints on the output tensor stores / query loads of the TMA kernel: A/B
tag=${1:-t6}
out=gpurun_out/$tag
mkdir -p $out
{
for lib in "" scripts/exp/libnaf_sth1.so scripts/exp/libnaf_sth2.so scripts/exp/libnaf_sth1q.so scripts/exp/libnaf_qh.so ""; do
  export NAF_B200_LIB=$lib
  [ -z "$lib" ] && unset NAF_B200_LIB
  timeout 120 python scripts/check_xattn.py cell_tma 2 768 224 8 7
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 1
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
done
} > $out/time_xattn.log 2>&1
cat $out/time_xattn.log

#!/bin/bash
# per-item bulk L2 prefetch of the query rows (loader warp, head-0 CTA pulls whole pixels): parity + A/B
tag=${1:-t5}
out=gpurun_out/$tag
mkdir -p $out
{
timeout 120 python scripts/check_xattn.py cell_tma 2 768 224 8 7
timeout 120 python scripts/check_xattn.py cell_tma 1 1024 336 12 11
} > $out/check.log 2>&1
{
for lib in scripts/exp/libnaf_qbulk0.so "" scripts/exp/libnaf_qbulk0.so ""; do
  export NAF_B200_LIB=$lib
  [ -z "$lib" ] && unset NAF_B200_LIB
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 1
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 1
done
} > $out/time_xattn.log 2>&1
cat $out/check.log; cat $out/time_xattn.log

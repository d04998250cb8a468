#!/bin/bash
tag=${1:-exp}
out=gpurun_out/$tag
mkdir -p $out
{
for lib in "" scripts/exp/libnaf_norope.so; do
  export NAF_B200_LIB=$lib
  [ -z "$lib" ] && unset NAF_B200_LIB
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 4 cell_tma 768 2048 32 7 4
done
unset NAF_B200_LIB
} > $out/time_xattn.log 2>&1
cat $out/time_xattn.log

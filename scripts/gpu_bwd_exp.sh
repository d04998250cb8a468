#!/bin/bash
out=gpurun_out/${1:-bwdexp}
mkdir -p $out
{
timeout 100 python scripts/time_bwd.py
for e in 1 2 3 4 8 16; do NAF_B200_LIB=scripts/exp/libnaf_bwdexp$e.so timeout 100 python scripts/time_bwd.py; done
} > $out/time_bwd.log 2>&1
cat $out/time_bwd.log

#!/bin/bash
out=gpurun_out/${1:-v15}
mkdir -p $out
for i in 1 2 3; do timeout 120 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "fused_path and case8" 2>&1 | tail -2; done > $out/elect.log 2>&1
for i in 1 2 3; do NAF_B200_LIB=scripts/exp/libnaf_bwdlane0.so timeout 120 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "fused_path and case8" 2>&1 | tail -2; done > $out/lane0.log 2>&1
NAF_B200_LIB=scripts/exp/libnaf_bwdlane0.so timeout 300 python -m pytest tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -3 > $out/lane0_all.log
NAF_B200_LIB=scripts/exp/libnaf_bwdlane0.so timeout 100 python scripts/time_bwd.py > $out/time_lane0.log 2>&1
echo "== elect"; cat $out/elect.log; echo "== lane0"; cat $out/lane0.log; cat $out/lane0_all.log; cat $out/time_lane0.log

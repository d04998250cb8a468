#!/bin/bash
# round 2 / session 4: backward kernel with a dedicated MMA warp: parity, timeline, timings
out=gpurun_out/${1:-v2}
mkdir -p $out
timeout 120 python scripts/check_bwd_tc.py > $out/check_bwd_tc.log 2>&1 || { echo "check_bwd_tc FAILED/TIMEOUT rc=$?"; tail -5 $out/check_bwd_tc.log; exit 1; }
tail -14 $out/check_bwd_tc.log
timeout 400 python -m pytest tests/test_gpu_backward.py -x -q -m gpu > $out/pytest_bwd.log 2>&1
tail -5 $out/pytest_bwd.log
NAF_B200_LIB=scripts/exp/libnaf_bwdtrace.so timeout 100 python scripts/trace_bwd.py > $out/trace.log 2>&1
{ timeout 100 python scripts/time_bwd.py; NAF_B200_LIB=scripts/exp/libnaf_bwdnopf.so timeout 100 python scripts/time_bwd.py; } > $out/time_bwd.log 2>&1
cat $out/time_bwd.log

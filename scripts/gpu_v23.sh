#!/bin/bash
out=gpurun_out/${1:-v23}
mkdir -p $out
( timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x 2>&1 | tail -3 ) > $out/pytest_bwd.log; cat $out/pytest_bwd.log
timeout 120 python scripts/check_bwd_tc.py > $out/check_bwd_tc.log 2>&1; tail -2 $out/check_bwd_tc.log | cut -c1-200
NAF_B200_LIB=scripts/exp/libnaf_bwdtrace.so timeout 100 python scripts/trace_bwd.py > $out/trace.log 2>&1
timeout 100 python scripts/time_bwd.py > $out/time_bwd.log 2>&1; cat $out/time_bwd.log

#!/bin/bash
# round 2 / session 4: backward kernel wrap-up: parity, sanitizers, timeline, timings, ncu capture, sweep
out=gpurun_out/${1:-v8}
mkdir -p $out
timeout 120 python scripts/check_bwd_tc.py > $out/check_bwd_tc.log 2>&1 || { echo "check_bwd_tc FAILED/TIMEOUT rc=$?"; tail -5 $out/check_bwd_tc.log; exit 1; }
tail -14 $out/check_bwd_tc.log
timeout 400 python -m pytest tests/test_gpu_backward.py -x -q -m gpu > $out/pytest_bwd.log 2>&1
tail -3 $out/pytest_bwd.log
NAF_B200_LIB=scripts/exp/libnaf_bwdtrace.so timeout 100 python scripts/trace_bwd.py > $out/trace.log 2>&1
timeout 100 python scripts/time_bwd.py > $out/time_bwd.log 2>&1
cat $out/time_bwd.log
( timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_memcheck_union_bwdtc.txt
( timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_racecheck_union_bwdtc.txt
tail -6 $out/compute_sanitizer_memcheck_union_bwdtc.txt; tail -6 $out/compute_sanitizer_racecheck_union_bwdtc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xattn_bwd_cell_tc -s 1 -c 1 -f -o $out/bwd_tc \
   python scripts/ncu_bwd.py > $out/ncu_bwd.log 2>&1
python scripts/ncu_summary.py $out/bwd_tc.ncu-rep 30 > $out/ncu_full_xattn_bwd_cell_tc_c2like_b1.txt 2>&1
rm -f $out/*.ncu-rep
head -42 $out/ncu_full_xattn_bwd_cell_tc_c2like_b1.txt
( timeout 600 python scripts/backward_sweep.py 2>&1 | tail -9 ) > $out/backward_sweep.log
head -8 $out/backward_sweep.log | cut -c1-200

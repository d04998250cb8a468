#!/bin/bash
# round 2 / session 4 wrap-up: whole GPU test suite, sanitizers and ncu capture of the backward kernel, bench (both arms)
out=gpurun_out/${1:-fin4}
mkdir -p $out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
( timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_memcheck_union_bwdtc.txt
( timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_racecheck_union_bwdtc.txt
tail -3 $out/compute_sanitizer_memcheck_union_bwdtc.txt; tail -3 $out/compute_sanitizer_racecheck_union_bwdtc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xattn_bwd_cell_tc -s 1 -c 1 -f -o $out/bwd_tc \
   python scripts/ncu_bwd.py > $out/ncu_bwd.log 2>&1
python scripts/ncu_summary.py $out/bwd_tc.ncu-rep 30 > $out/ncu_full_xattn_bwd_cell_tc_c2like_b1.txt 2>&1
rm -f $out/*.ncu-rep
head -34 $out/ncu_full_xattn_bwd_cell_tc_c2like_b1.txt
NAF_B200_LIB=scripts/exp/libnaf_bwdtrace.so timeout 100 python scripts/trace_bwd.py > $out/trace.log 2>&1
timeout 100 python scripts/time_bwd.py > $out/time_bwd.log 2>&1; cat $out/time_bwd.log
timeout 120 python scripts/check_bwd_tc.py > $out/check_bwd_tc.log 2>&1; tail -3 $out/check_bwd_tc.log | cut -c1-220
( timeout 600 python scripts/backward_sweep.py 2>&1 | tail -9 ) > $out/backward_sweep.log
head -8 $out/backward_sweep.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_c2_step.csv \
   python bench.py --steps 2 --warmup 1 --configs '' --no-cpu-baseline --no-other-paths --e2e-d2h sample > $out/launches_bench.log 2>&1
timeout 900 python bench.py > $out/bench_c2_n1.json 2> $out/bench.err
timeout 600 python bench.py --impl reference > $out/bench_reference_arm.json 2> $out/bench_ref.err
python - <<PY
import json
d = json.load(open("$out/bench_c2_n1.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("roofline", {}).get("kernel_ms"), v.get("roofline", {}).get("frac")) for k, v in d.get("configs", {}).items()})
print(json.dumps(d.get("other_paths"), indent=1)[:2500])
PY
tail -2 $out/bench_reference_arm.json | cut -c1-400

#!/bin/bash
# sanitizer runs + ncu capture of the union kernel + the full bench line
out=gpurun_out/${1:-fin}
mkdir -p $out
( timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_memcheck_new_kernels.txt
( timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_racecheck_new_kernels.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xattn_union_tc -s 1 -c 1 -f -o $out/union \
   python scripts/ncu_union.py > $out/ncu_union.log 2>&1
python scripts/ncu_summary.py $out/union.ncu-rep 30 > $out/ncu_full_xattn_union_tc_denoise.txt 2>&1
timeout 900 python bench.py > $out/bench_c2.json 2> $out/bench.err
tail -8 $out/compute_sanitizer_memcheck_new_kernels.txt; tail -8 $out/compute_sanitizer_racecheck_new_kernels.txt
head -30 $out/ncu_full_xattn_union_tc_denoise.txt
python - <<PY
import json
d = json.load(open("$out/bench_c2.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"])
print(json.dumps(d.get("other_paths"), indent=1))
print({k: (v.get("ms_per_step"), v.get("roofline", {}).get("kernel_ms")) for k, v in d.get("configs", {}).items()})
PY

#!/bin/bash
# round 2 / session 4, last state: whole GPU test suite, smoke(), bench (our arm), SASS counts
out=gpurun_out/${1:-fin5}
mkdir -p $out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > $out/smoke.log
tail -2 $out/smoke.log
timeout 900 python bench.py > $out/bench_c2_n1.json 2> $out/bench.err
python - <<PY
import json
d = json.load(open("$out/bench_c2_n1.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("roofline", {}).get("kernel_ms"), v.get("roofline", {}).get("frac")) for k, v in d.get("configs", {}).items()})
print({k: (v.get("backward", v.get("auto", {})).get("ms"), v.get("backward", v.get("auto", {})).get("frac")) for k, v in d.get("other_paths", {}).items() if isinstance(v, dict)})
PY
timeout 200 python scripts/step_breakdown.py > $out/step_breakdown_c2.txt 2>&1; tail -25 $out/step_breakdown_c2.txt

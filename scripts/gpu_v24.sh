#!/bin/bash
out=gpurun_out/${1:-v24}
mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $out/pytest_gpu.log; tail -2 $out/pytest_gpu.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 ) > $out/smoke.log; cat $out/smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 --configs '' --no-cpu-baseline > $out/bench_quick.json 2> $out/bench.err
python - <<PY
import json
d = json.load(open("$out/bench_quick.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
print({k: (v.get("backward", v.get("auto", {})).get("ms"), v.get("backward", v.get("auto", {})).get("frac")) for k, v in d.get("other_paths", {}).items() if isinstance(v, dict)})
PY

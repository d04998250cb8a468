#!/bin/bash
# full bench line with the new legs (C1 graph replay, same-GPU eager reference), smoke(), then N=2 under torchrun
out=gpurun_out/${1:-t11}
mkdir -p $out
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 ) > $out/smoke.log
timeout 900 python bench.py > $out/bench_c2_n1.json 2> $out/bench.err
tail -6 $out/smoke.log
python - <<PY
import json
d = json.load(open("$out/bench_c2_n1.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["clocks"])
print("C1", json.dumps(d["configs"]["C1"].get("graph_replay")), d["configs"]["C1"]["ms_per_step"])
print("gpu_eager_reference", json.dumps(d.get("gpu_eager_reference"))[:400])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
tail -3 $out/bench.err

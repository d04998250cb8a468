#!/bin/bash
out=gpurun_out/${1:-n2}
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
tail -3 $out/bench_n2.err
python - <<PY
import json
d = json.loads(open("$out/bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("roofline", {}).get("kernel_ms")) for k, v in d.get("configs", {}).items()})
PY

#!/bin/bash
out=gpurun_out/${1:-enc}
mkdir -p $out
( timeout 500 python -m pytest tests/test_gpu_encoder.py -m gpu -q --tb=short 2>&1 | tail -60 ) > $out/pytest_enc.log
( timeout 300 python scripts/step_breakdown.py 2>&1 | tail -50 ) > $out/step_breakdown.log
( timeout 400 python bench.py --configs '' --e2e-d2h sample --no-cpu-baseline 2>$out/bench.err | tail -1 ) > $out/bench_c2.json
tail -15 $out/pytest_enc.log; cat $out/step_breakdown.log | tail -40; python -c "
import json; d=json.load(open('$out/bench_c2.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"

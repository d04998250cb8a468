#!/bin/bash
# TMA kernel (table rows in shared memory) against the round-1 pipelined kernel on the shapes AUTO gives to the latter
tag=${1:-t4}
out=gpurun_out/$tag
mkdir -p $out
{
for algo in cell_tcws cell_tma cell_tcws cell_tma; do
  timeout 120 python scripts/time_xattn.py 2 $algo 768 1344 24 7 4
  timeout 120 python scripts/time_xattn.py 4 $algo 768 2048 32 7 4
  timeout 120 python scripts/time_xattn.py 1 $algo 384 224 16 7 1
  timeout 120 python scripts/time_xattn.py 8 $algo 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 4 $algo 1024 1036 37 11 2
done
} > $out/time_xattn.log 2>&1
( timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 ) > $out/pytest_gpu.log
cat $out/time_xattn.log; tail -4 $out/pytest_gpu.log

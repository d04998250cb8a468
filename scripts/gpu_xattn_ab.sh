#!/bin/bash
# A/B timing of the attention kernels on the BASELINE shapes + the parity tests of the attention path
tag=${1:-ab}; shift
out=gpurun_out/$tag
mkdir -p $out
( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tma.py -m gpu -q --tb=short -x "$@" 2>&1 | tail -150 ) > $out/pytest.log
tail -4 $out/pytest.log
{
for algo in cell_tcws cell_tma; do
  timeout 120 python scripts/time_xattn.py 8 $algo 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 8 $algo 768 896 32 7 1
  timeout 120 python scripts/time_xattn.py 1 $algo 384 224 16 7 1
  timeout 120 python scripts/time_xattn.py 4 $algo 1024 1036 37 11 2
  timeout 120 python scripts/time_xattn.py 2 $algo 768 1344 24 7 4
  timeout 120 python scripts/time_xattn.py 4 $algo 768 2048 32 7 4
done
} > $out/time_xattn.log 2>&1
cat $out/time_xattn.log

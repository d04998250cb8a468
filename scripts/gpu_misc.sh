#!/bin/bash
out=gpurun_out/${1:-misc}
mkdir -p $out
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | tail -60 ) > $out/pytest_parity.log
tail -30 $out/pytest_parity.log

#!/bin/bash
out=gpurun_out/${1:-v21}
mkdir -p $out
( timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x 2>&1 | tail -3 ) > $out/pytest_bwd.log; cat $out/pytest_bwd.log
( NAF_B200_LIB=scripts/exp/libnaf_bwdelectall.so timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q 2>&1 | tail -6 ) > $out/pytest_bwd_electall.log; cat $out/pytest_bwd_electall.log
NAF_B200_LIB=scripts/exp/libnaf_bwdelectall.so timeout 120 python scripts/repro_bwd.py 1 768 88 11 11 2>&1 | tail -2
NAF_B200_LIB=scripts/exp/libnaf_bwdelectall.so timeout 120 python scripts/repro_bwd.py 1 256 154 11 11 2>&1 | tail -2

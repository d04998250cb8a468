#!/bin/bash
# union-window kernel: MMAs issued under elect.sync (regular) vs by lane 0 (libnaf_unionlane0.so)
out=gpurun_out/${1:-v26}
mkdir -p $out
{ echo "== regular (elect)"; timeout 60 python scripts/check_union.py 2>&1 | tail -12; echo "== lane 0"; NAF_B200_LIB=scripts/exp/libnaf_unionlane0.so timeout 60 python scripts/check_union.py 2>&1 | tail -12; } > $out/check_union.log 2>&1
cat $out/check_union.log | cut -c1-230
( timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 ) > $out/pytest.log; cat $out/pytest.log

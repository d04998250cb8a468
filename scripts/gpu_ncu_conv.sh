#!/bin/bash
out=gpurun_out/${1:-ncuconv}
mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv128_ws_kernel -s 8 -c 1 -f -o $out/conv3x3 \
   python scripts/ncu_conv.py 1 > $out/ncu_conv.log 2>&1
python scripts/ncu_summary.py $out/conv3x3.ncu-rep 40 > $out/ncu_full_conv128_ws_3x3_f16_c2.txt 2>&1
( timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -q --tb=short -k stem 2>&1 | tail -5 ) > $out/pytest_stem.log
cat $out/ncu_full_conv128_ws_3x3_f16_c2.txt; tail -3 $out/pytest_stem.log

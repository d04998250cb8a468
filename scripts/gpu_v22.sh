#!/bin/bash
# session 4 evidence on the final library: backward tests, phase timeline, ncu of the 3x3 conv layer and of the backward kernel
out=gpurun_out/${1:-v22}
mkdir -p $out
( timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_gpu_encoder.py -m gpu -q -x 2>&1 | tail -3 ) > $out/pytest.log; cat $out/pytest.log
NAF_B200_LIB=scripts/exp/libnaf_bwdtrace.so timeout 100 python scripts/trace_bwd.py > $out/trace.log 2>&1
timeout 100 python scripts/time_bwd.py > $out/time_bwd.log 2>&1; cat $out/time_bwd.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv128_ws_kernel -s 5 -c 1 -f -o $out/conv3x3 \
   python scripts/ncu_conv.py 1 > $out/ncu_conv.log 2>&1
python scripts/ncu_summary.py $out/conv3x3.ncu-rep 30 > $out/ncu_full_conv128_ws_3x3_f16_c2.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xattn_bwd_cell_tc -s 1 -c 1 -f -o $out/bwd_tc \
   python scripts/ncu_bwd.py > $out/ncu_bwd.log 2>&1
python scripts/ncu_summary.py $out/bwd_tc.ncu-rep 30 > $out/ncu_full_xattn_bwd_cell_tc_c2like_b1.txt 2>&1
rm -f $out/*.ncu-rep
head -32 $out/ncu_full_conv128_ws_3x3_f16_c2.txt
head -24 $out/ncu_full_xattn_bwd_cell_tc_c2like_b1.txt

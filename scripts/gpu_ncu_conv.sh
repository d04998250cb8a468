#!/bin/bash
# ncu --set full of one 3x3 GroupNorm+SiLU+conv layer (fp16 in / fp16 out, C2 guidance size)
out=gpurun_out/${1:-ncuconv}
mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv128_ws_kernel -s 5 -c 1 -f -o $out/conv3x3 \
   python scripts/ncu_conv.py 1 > $out/ncu_conv.log 2>&1
python scripts/ncu_summary.py $out/conv3x3.ncu-rep 50 > $out/ncu_full_conv128_ws_3x3_f16_c2.txt 2>&1
cat $out/ncu_full_conv128_ws_3x3_f16_c2.txt

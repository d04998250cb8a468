#!/bin/bash
# ncu --set full captures of the attention kernel at C2 and C3 (third launch of each), summarised
out=gpurun_out/${1:-ncu}
mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xattn_cell_tma -s 2 -c 1 -f -o $out/tma_c2_rep2 \
   python scripts/ncu_xattn.py 8 cell_tma 2 > $out/ncu_c2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xattn_cell_tma -s 2 -c 1 -f -o $out/tma_c3_rep2 \
   python scripts/ncu_xattn.py 4 cell_tma 2 1024 1036 37 11 > $out/ncu_c3.log 2>&1
python scripts/ncu_summary.py $out/tma_c2_rep2.ncu-rep 30 > $out/ncu_full_xattn_cell_tma_c2_b8_rep2.txt 2>&1
python scripts/ncu_summary.py $out/tma_c3_rep2.ncu-rep 30 > $out/ncu_full_xattn_cell_tma_c3_b4_rep2.txt 2>&1
ls -la $out; head -40 $out/ncu_full_xattn_cell_tma_c3_b4_rep2.txt

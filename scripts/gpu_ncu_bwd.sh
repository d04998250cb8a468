#!/bin/bash
out=gpurun_out/${1:-ncubwd}
mkdir -p $out
timeout 300 python scripts/check_bwd_tc.py > $out/check.log 2>&1; tail -14 $out/check.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xattn_bwd_cell_tc -s 1 -c 1 -f -o $out/bwd_tc \
   python scripts/ncu_bwd.py > $out/ncu_bwd.log 2>&1
python scripts/ncu_summary.py $out/bwd_tc.ncu-rep 45 > $out/ncu_full_xattn_bwd_cell_tc.txt 2>&1
cat $out/ncu_full_xattn_bwd_cell_tc.txt

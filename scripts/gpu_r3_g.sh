#!/bin/bash
# x RoPE tables as chunk planes (conflict-free shared-memory reads): parity + timing + bank-conflict count
tag=${1:-t8}
out=gpurun_out/$tag
mkdir -p $out
{
timeout 120 python scripts/check_xattn.py cell_tma 2 768 224 8 7
timeout 120 python scripts/check_xattn.py cell_tma 1 1024 252 9 9
timeout 120 python scripts/check_xattn.py cell_tma 2 384 256 16 5
timeout 120 python scripts/check_xattn.py cell_tma 1 768 240 8 7
} > $out/check.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -30 ) > $out/pytest.log
{
for i in 1 2; do
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 1
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
done
} > $out/time_xattn.log 2>&1
timeout 300 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum --clock-control none -k regex:xattn_cell_tma -s 2 -c 1 python scripts/ncu_xattn.py 8 cell_tma 2 2>&1 | grep -E "l1tex|gpu__time" > $out/ncu_conflicts.log
cat $out/check.log; tail -3 $out/pytest.log; cat $out/time_xattn.log $out/ncu_conflicts.log

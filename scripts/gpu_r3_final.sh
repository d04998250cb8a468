#!/bin/bash
# round 2 / session 3 wrap-up: ncu captures (with tensor-memory counters) of the attention kernel at C2 / C3 and of the
# 3x3 conv layer, the launch list of one bench step, the full bench line
out=gpurun_out/${1:-fin3}
mkdir -p $out
TM=sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,sm__mem_tensor_reads_op_utcmma_matrix_c.sum,sm__mem_tensor_writes_op_utcmma.sum,sm__inst_executed_pipe_tmem.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg
timeout 400 ncu --set full --metrics $TM --clock-control none --import-source on -k regex:xattn_cell_tma -s 2 -c 1 -f -o $out/tma_c2_rep2 \
   python scripts/ncu_xattn.py 8 cell_tma 2 > $out/ncu_c2.log 2>&1
timeout 400 ncu --set full --metrics $TM --clock-control none --import-source on -k regex:xattn_cell_tma -s 2 -c 1 -f -o $out/tma_c3_rep2 \
   python scripts/ncu_xattn.py 4 cell_tma 2 1024 1036 37 11 > $out/ncu_c3.log 2>&1
timeout 400 ncu --set full --metrics $TM --clock-control none --import-source on -k regex:conv128_ws_kernel -s 9 -c 1 -f -o $out/conv3 \
   python scripts/enc_bench.py 8 448 448 1 > $out/ncu_conv.log 2>&1
python scripts/ncu_summary.py $out/tma_c2_rep2.ncu-rep 30 > $out/ncu_full_xattn_cell_tma_c2_b8_rep2.txt 2>&1
python scripts/ncu_summary.py $out/tma_c3_rep2.ncu-rep 30 > $out/ncu_full_xattn_cell_tma_c3_b4_rep2.txt 2>&1
python scripts/ncu_summary.py $out/conv3.ncu-rep 30 > $out/ncu_full_conv128_ws_c2.txt 2>&1
rm -f $out/*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_c2_step.csv \
   python bench.py --steps 2 --warmup 1 --configs '' --no-cpu-baseline --no-other-paths --e2e-d2h sample > $out/launches_bench.log 2>&1
timeout 900 python bench.py > $out/bench_c2_n1.json 2> $out/bench.err
timeout 600 python bench.py --impl reference > $out/bench_reference_arm.json 2> $out/bench_ref.err
head -45 $out/ncu_full_xattn_cell_tma_c2_b8_rep2.txt; head -40 $out/ncu_full_xattn_cell_tma_c3_b4_rep2.txt; head -45 $out/ncu_full_conv128_ws_c2.txt
python - <<PY
import json
d = json.load(open("$out/bench_c2_n1.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("roofline", {}).get("kernel_ms"), v.get("roofline", {}).get("frac")) for k, v in d.get("configs", {}).items()})
PY
tail -2 $out/bench_reference_arm.json | cut -c1-600

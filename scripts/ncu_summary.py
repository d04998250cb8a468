"""Summarise an .ncu-rep (first kernel): key metrics + hottest source lines.  Usage:
   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        # tensor-memory traffic (captured with --metrics on top of --set full, see scripts/gpu_r3_final.sh)
        "sm__mem_tensor_reads.sum", "sm__mem_tensor_writes.sum", "sm__mem_tensor_reads_op_utcmma_matrix_c.sum",
        "sm__mem_tensor_writes_op_utcmma.sum", "sm__inst_executed_pipe_tmem.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in keys or h.startswith("smsp__average_warps_issue_stalled") and "per_issue_active" in h and float(v or 0) > 0.15:
        print(f"{h:86s} {v:>16s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
cur, hd, L = None, None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hd = r
    elif hd and len(r) == len(hd) and r[0].strip().isdigit():
        L.append((cur, int(r[0]), r[1].strip()[:88], int(r[7] or 0), int(r[6] or 0)))
ti, ts = sum(x[3] for x in L), sum(x[4] for x in L)
print(f"--- hottest source lines (of {ti} warp-inst, {ts} samples)")
for x in sorted(L, key=lambda x: -x[4])[:nl]:
    print(f"{100*x[4]/ts:5.1f}%smp {100*x[3]/ti:5.1f}%inst {x[0]}:{x[1]:<4d} {x[2]}")

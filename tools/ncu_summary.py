"""Condense an ncu raw-page CSV into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
want = [
 "Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
 "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
 "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
 "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
for k in want:
    if k in hdr:
        i = hdr.index(k); print(f"{k:90s} {[r[i][:60] for r in data]}")
for k in hdr:
    if "average_warps_issue_stalled" in k and "not_issued" not in k:
        i = hdr.index(k); v = [round(float(r[i]), 2) for r in data]
        if max(v) >= 0.1: print(f"{k:90s} {v}")

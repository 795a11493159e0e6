"""Digest of one `ncu --set full` report: the counters DESIGN.md quotes plus the stall samples by opcode.
  python tools/ncu_digest.py <report.ncu-rep> > profiles/r2_ncu_<tag>.txt      (runs on the GPU box: reports are ~17 MB each)"""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, data = raw[0], raw[1], raw[2:]
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]
print(f"# {rep.split('/')[-1]}")
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:86s} {data[0][i][:110]} {units[i]}")
print("# warp stall reasons, cycles per issued instruction (>= 0.1)")
for k in hdr:
    if "average_warps_issue_stalled" in k and "not_issued" not in k:
        v = float(data[0][hdr.index(k)])
        if v >= 0.1:
            print(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:6.2f}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = next((r for r in src if "# Samples" in r), None)
if h:
    idx = {c: i for i, c in enumerate(h)}
    tot, byop, cnt, stall = 0, collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for r in src[src.index(h) + 1:]:
        if len(r) < len(h):
            continue
        try:
            smp, ie = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
        except ValueError:
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[idx["Source"]]).split()[0] if r[idx["Source"]].strip() else "?"
        tot += smp; byop[op] += smp; cnt[op] += ie
        for c in h:
            if c.startswith("stall_") and "Not Issued" not in c and r[idx[c]].isdigit():
                stall[op][c[6:]] += int(r[idx[c]])
    print(f"# stall samples by opcode ({tot} samples): share, warp-level executions, top stall reasons")
    for op, s in byop.most_common(12):
        top = ", ".join(f"{k} {100 * v / max(s, 1):.0f}%" for k, v in stall[op].most_common(3))
        print(f"{op:18s} {100 * s / max(tot, 1):5.1f}%  {cnt[op]:>13d}   {top}")

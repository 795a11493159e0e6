"""Stall samples and executed warp instructions of an ncu report by source region of jne_kernels.cuh / jne_rng.cuh."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# function line ranges from the source
import bisect
src = open("johansen_null_eigenspectra_b200/csrc/jne_kernels.cuh").read().splitlines()
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:__device__ __forceinline__|__global__)?.*?\b(jne_\w+)\s*\(", l)
    if m and not l.startswith(" ") and not l.startswith("//"):
        starts.append((i, m.group(1)))
def region(f, ln):
    if f != "jne_kernels.cuh": return f
    k = bisect.bisect_right([s for s, _ in starts], ln) - 1
    return starts[k][1] if k >= 0 else "?"
hdr = None; cur_line = None; cur_file = None
smp_r = collections.Counter(); ie_r = collections.Counter(); tot = 0; tie = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 8: continue
    if r[2] == "-": cur_line = int(r[0]); continue
    try: smp = int(r[idx["# Samples"]]); ie = int(r[idx["Instructions Executed"]])
    except ValueError: continue
    reg = region(cur_file, cur_line or 0)
    smp_r[reg] += smp; ie_r[reg] += ie; tot += smp; tie += ie
print("total samples", tot, "warp instructions", tie)
for k, v in smp_r.most_common():
    print(f"{k:28s} samples {100*v/tot:5.1f}%   instructions {100*ie_r[k]/tie:5.1f}%")

"""SASS rows of an ncu report with their dominant stall reasons; optional filter by source line numbers."""
import csv, subprocess, sys
rep = sys.argv[1]
lines = set(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 and sys.argv[2] else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; cur_line = None; cur_file = None
tot = 0
out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[2] == "-": cur_line = int(r[0]); continue
    idx = {h: i for i, h in enumerate(hdr)}
    try: smp = int(r[idx["# Samples"]])
    except ValueError: continue
    tot += smp
    stalls = {h[6:]: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i]) > 0}
    out.append((cur_file, cur_line, r[3], smp, int(r[idx["Instructions Executed"]]), stalls))
print("total samples", tot)
for f, ln, sass, smp, ie, st in out:
    if lines is not None and not (f == "jne_kernels.cuh" and ln in lines): continue
    if lines is None and smp < tot * 0.004: continue
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f"{f}:{(ln or 0):4d} {100*smp/tot:5.2f}% {ie:>10d}  {sass[:60]:60s} {top}")

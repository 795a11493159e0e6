"""Per-source-line sample / instruction shares from an ncu report (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 8 or r[2] != "-": continue      # source-line summary rows have Address '-'
    try: s = int(r[hdr["# Samples"]]); ie = int(r[hdr["Instructions Executed"]])
    except ValueError: continue
    k = (cur, int(r[0]))
    a = agg.setdefault(k, [r[1], 0, 0]); a[1] += s; a[2] += ie
ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values())
print(f"total samples {ts}  warp instructions {ti}")
for (f, ln), (src, s, ie) in sorted(agg.items()):
    if 100 * s / ts >= thr or 100 * ie / ti >= thr:
        print(f"{f}:{ln:4d} {100*s/ts:5.1f}% smp {100*ie/ti:5.1f}% ins  {src.strip()[:100]}")
if len(sys.argv) > 3:   # ranges "name:lo-hi,name:lo-hi" over jne_kernels.cuh
    for spec in sys.argv[3].split(","):
        nm, rg = spec.split(":"); lo, hi = map(int, rg.split("-"))
        s = sum(a[1] for (f, ln), a in agg.items() if f == "jne_kernels.cuh" and lo <= ln <= hi)
        ie = sum(a[2] for (f, ln), a in agg.items() if f == "jne_kernels.cuh" and lo <= ln <= hi)
        print(f"{nm:12s} {100*s/ts:5.1f}% smp {100*ie/ti:5.1f}% ins")
    s = sum(a[1] for (f, ln), a in agg.items() if f != "jne_kernels.cuh"); ie = sum(a[2] for (f, ln), a in agg.items() if f != "jne_kernels.cuh")
    print(f"{'other files':12s} {100*s/ts:5.1f}% smp {100*ie/ti:5.1f}% ins")

"""Instruction mix of the hottest loop (the backward branch spanning the most DMMAs) of a kernel in libjne.so."""
import collections, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "jne_run_kernelILi12ELi0ELb1E"
out = subprocess.run(["cuobjdump", "-sass", "johansen_null_eigenspectra_b200/libjne.so"], capture_output=True, text=True).stdout
keep, on = [], False
for l in out.splitlines():
    if "Function :" in l:
        on = pat in l
    if on:
        keep.append(l)
ins = []
for l in keep:
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
dm = [a for a, t in ins if "DMMA" in t]
print("total instrs", len(ins), "DMMA", len(dm))
best = None
for a, t in ins:
    m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
    if m:
        lo = int(m.group(1), 16)
        if lo < a:
            n = sum(1 for x in dm if lo <= x <= a)
            if n:
                print(f"candidate loop {lo:#x}..{a:#x} DMMA {n} len {(a-lo)//16+1}")
                if best is None or (a - lo) < (best[1] - best[0]):
                    best = (lo, a, n)
if len(sys.argv) > 2 and sys.argv[2].startswith("0x"):
    lo = int(sys.argv[2], 16); hi = int(sys.argv[3], 16); n = sum(1 for x in dm if lo <= x <= hi)
else:
    lo, hi, n = best
body = [t for x, t in ins if lo <= x <= hi]
c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
print(f"loop {lo:#x}..{hi:#x}: {len(body)} instrs, {n} DMMA")
for k, v in c.most_common():
    print(f"  {k:24s}{v}")
if "-v" in sys.argv:
    for x, t in ins:
        if lo <= x <= hi: print(f"{x:#06x}  {t}")

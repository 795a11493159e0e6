"""Instruction mix of the largest backward-branch loop of a kernel in a library: tools/sass_mix.py <symbol-substring> [lib]."""
import collections, re, subprocess, sys
pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else "johansen_null_eigenspectra_b200/libjne.so"
out = subprocess.run(["cuobjdump", "-sass", "-fun", pat, lib], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
print("total instrs", len(ins), "loops", [(hex(l), hex(h), (h - l) // 16 + 1) for l, h in loops])
lo, hi = max(loops, key=lambda p: p[1] - p[0])
body = [t for x, t in ins if lo <= x <= hi]
c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
print(f"loop {lo:#x}..{hi:#x}: {len(body)} instrs")
for k, v in c.most_common():
    print(f"  {k:24s}{v}")
if "-v" in sys.argv:
    for x, t in ins:
        if lo <= x <= hi: print(f"{x:#06x}  {t}")

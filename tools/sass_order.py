"""Instruction ORDER of the main loop of a kernel as a string of pipe classes (how finely ptxas interleaved the
generator with the FP64 work): D = DFMA/DADD/DMUL, W = IMAD.WIDE, L = LOP3, X = MUFU, C = F2F, f = other FP32, . = rest.
  python tools/sass_order.py <mangled-name-substring> [lib]"""
import re, subprocess, sys
pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "johansen_null_eigenspectra_b200/libjne.so"
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ins = None, []
for l in names.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a: loops.append((int(m.group(1), 16), a))
lo, hi = max(loops, key=lambda p: p[1] - p[0])
def cls(t):
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
    if op in ("DFMA", "DADD", "DMUL"): return "D"
    if op.startswith("IMAD.WIDE"): return "W"
    if op.startswith("LOP3"): return "L"
    if op.startswith("MUFU"): return "X"
    if op.startswith("F2F"): return "C"
    if op[0] == "F" or op.startswith("I2FP"): return "f"
    if op.startswith("DMMA"): return "T"
    return "."
s = "".join(cls(t) for a, t in ins if lo <= a <= hi)
print(len(s), "instructions in the loop")
for i in range(0, len(s), 120): print(s[i:i + 120])

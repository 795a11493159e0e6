"""Instruction-class pattern of the main loop (the backward branch spanning 40 DMMAs) of a kernel's SASS dump."""
import re, sys
from collections import Counter
lines = open(sys.argv[1]).read().splitlines()
ndm = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ins = []
for l in lines:
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
dm = [a for a, t in ins if 'DMMA' in t]
def cls(t):
    op = t.split()[0] if not t.startswith('@') else t.split()[1]
    if 'DMMA' in op: return 'M'
    if op.startswith(('DADD', 'DFMA', 'DMUL', 'DSETP')): return 'D'
    if op.startswith('IMAD.WIDE'): return 'W'
    if op.startswith('MUFU'): return 'u'
    if op.startswith(('F2F', 'I2F')): return 'c'
    if op.startswith('SHFL'): return 's'
    if op.startswith(('LDL', 'STL')): return 'L'
    return '.'
for a, t in ins:
    m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
    if m:
        lo = int(m.group(1), 16)
        if lo < a and sum(1 for x in dm if lo <= x <= a) == ndm:
            seq = [(x, y) for x, y in ins if lo <= x <= a]
            s = ''.join(cls(t) for _, t in seq)
            print("loop", hex(lo), hex(a), len(seq), "instructions")
            for i in range(0, len(s), 100): print(s[i:i + 100])
            print(Counter(s))
            break

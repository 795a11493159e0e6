"""Registers / spills / shared memory per kernel from `nvcc -Xptxas=-v` output (stdin or file)."""
import re, subprocess, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
cur = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1); spill = ""
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        spill = f"stack {m.group(1)} spill {m.group(2)}/{m.group(3)}"
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"{int(m.group(1)):4d} regs  {spill:28s} {name}")
        cur = None

"""Aggregate the warp-stall samples of an ncu report (source page) by opcode, by stall reason and by source function region."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
by_op = collections.Counter(); by_reason = collections.Counter(); exec_op = collections.Counter()
by_op_reason = collections.defaultdict(collections.Counter)
tot = 0
for r in rows:
    if not r: continue
    if "Source" in r and "# Samples" in r: hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr): continue
    try: smp = int(r[idx["# Samples"]]); ie = int(r[idx["Instructions Executed"]])
    except ValueError: continue
    sass = r[idx["Source"]]
    op = re.sub(r"^@!?U?P\d+\s+", "", sass.strip()).split()[0] if sass.strip() else "?"
    op = op.split(".")[0] if not op.startswith(("DMMA", "IMAD.WIDE", "F2F", "MUFU", "SHFL", "LDG", "LDS", "STS")) else ".".join(op.split(".")[:2])
    tot += smp; by_op[op] += smp; exec_op[op] += ie
    for h, i in idx.items():
        if h.startswith("stall_") and r[i].isdigit() and "Not Issued" not in h:
            by_reason[h[6:]] += int(r[i]); by_op_reason[op][h[6:]] += int(r[i])
print("total samples", tot)
print("by reason:", [(k, round(100 * v / tot, 1)) for k, v in by_reason.most_common(12)])
for op, v in by_op.most_common(25):
    print(f"{op:16s} {100*v/tot:5.1f}%  exec {exec_op[op]:>12d}  {[(k, round(100*x/tot,1)) for k, x in by_op_reason[op].most_common(4)]}")

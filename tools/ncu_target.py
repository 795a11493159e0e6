"""Short single-GPU target for ncu captures: a few launches of the fused kernel at the metric's config."""
import sys
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne

n = int(sys.argv[1]) if len(sys.argv) > 1 else 23680
models = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
eng = jne.Engine([0])
seeds = np.arange(1, n + 1, dtype=np.uint32)
for m in models:                     # warm-up launches (skip with ncu -s)
    eng.eigs_batch(m, 12, 10000, seeds[:592])
for m in models:
    out = eng.eigs_batch(m, 12, 10000, seeds)
    print(m, out.shape, float(out.sum(1).mean()))

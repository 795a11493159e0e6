"""Dump eigenvalues for a grid of (dim, T, models) with the library named by JNE_LIBRARY (regression aid:
two builds whose arithmetic is meant to be identical must produce identical files)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0])
out = {}
seeds = np.arange(1, 2049, dtype=np.uint32)
for dim in (1, 2, 3, 4, 5, 7, 8, 9, 11, 12, 13, 15):
    for T in (37, 1000):
        for m in range(5):
            if m == 4 and T < 3: continue
            out[f"d{dim}_T{T}_m{m}"] = eng.eigs_batch(m, dim, T, seeds)
        res = eng.eigs_batch_multi(range(5), dim, T, seeds)
        for m in range(5): out[f"d{dim}_T{T}_multi{m}"] = res[m]
rng = np.random.default_rng(5)
for dim in (2, 6, 12, 15):
    db = rng.standard_normal((64, 200, dim)) * 3.7e-3
    for m in range(5): out[f"inc_d{dim}_m{m}"] = eng.eigs_from_increments(m, db)
np.savez(sys.argv[1], **out)
print("dumped", len(out), "arrays to", sys.argv[1])

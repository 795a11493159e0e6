"""Small workload for compute-sanitizer: every kernel family once (single, multi, increments, pencil, normals)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0])
seeds = np.arange(1, 41, dtype=np.uint32)
rng = np.random.default_rng(0)
for dim, T in [(1, 9), (3, 40), (5, 33), (8, 64), (12, 100), (15, 70)]:
    for m in range(5):
        eng.eigs_batch(m, dim, T, seeds)
    eng.eigs_batch_multi(range(5), dim, T, seeds)
    eng.eigs_from_increments(3, rng.standard_normal((7, T, dim)) / np.sqrt(T))
eng.gen_normal_matrix(12, 103, 5); eng.brownian_motion_matrix(3, 50, 0.02, 5)
S1 = rng.standard_normal((9, 12, 13)); F = rng.standard_normal((9, 13, 40)); eng.pencil_eigs_batch(S1, F @ np.transpose(F, (0, 2, 1)))
print("sanitize target done")

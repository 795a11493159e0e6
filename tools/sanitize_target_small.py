"""Reduced compute-sanitizer target for the kernels changed last: lane dim 6 (generator states in shared memory) and the
unaligned-segment instance of the tensor kernels (short horizons), single-model and fused."""
import sys
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0])
seeds = np.arange(1, 41, dtype=np.uint32)
for dim, T in [(6, 33), (6, 300), (8, 64), (12, 100), (12, 200), (10, 700)]:
    for m in (0, 3, 4):
        eng.eigs_batch(m, dim, T, seeds)
    eng.eigs_batch_multi(range(5), dim, T, seeds)
print("sanitize target (small) done")

"""ncu target: the fused five-model kernel at tiny T so that the per-run epilogue dominates."""
import sys
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0])
seeds = np.arange(1, 29600 + 1, dtype=np.uint32)
eng.eigs_batch_multi(range(5), 12, 32, seeds[:1024])
out = eng.eigs_batch_multi(range(5), 12, 32, seeds)
print(out[0].shape)

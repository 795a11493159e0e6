import sys
import numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
bad = 0; worst = 0.0
for k in a.files:
    x, y = a[k], b[k]
    if x.shape != y.shape: print("SHAPE", k, x.shape, y.shape); bad += 1; continue
    if not np.array_equal(x, y):
        rel = np.abs(x - y) / np.maximum(np.abs(x), 1e-300)
        big = x > 1e-9 * x.max(axis=-1, keepdims=True)
        w = rel[big].max() if big.any() else 0.0
        worst = max(worst, w); bad += 1
        print(f"DIFF {k}: {np.count_nonzero(x != y)} of {x.size} values differ, worst rel (non-negligible) {w:.3e}, max abs {np.abs(x-y).max():.3e}")
print(f"{len(a.files)} arrays, {bad} differ, worst rel {worst:.3e}")

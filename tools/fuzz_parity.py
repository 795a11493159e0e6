"""Randomised pathwise parity sweep (run once per round for profiles/, not part of the suite): random (model, dim, T,
seeds) through jne_eigs_batch vs the oracle fed the device normals, plus fused-pass bit-identity."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
from oracle import johansen_oracle as orc
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 2026)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 400
eng = jne.Engine([0])
worst, bad, t0 = 0.0, 0, time.time()
per_dim = {}
for i in range(ncase):
    dim = int(rng.integers(1, 16))
    T = int(rng.integers(2 * dim + 8, 4000)) if rng.random() < 0.9 else int(rng.integers(4000, 40000))
    model = int(rng.integers(0, 5))
    seeds = rng.integers(0, 2**32, size=3, dtype=np.uint64).astype(np.uint32)
    got = eng.eigs_batch(model, dim, T, seeds)
    ref = np.stack([orc.eigs_from_normals(eng.gen_normal_matrix(dim, T, int(s)), model) for s in seeds])
    tol = 1e-9 * np.abs(ref) + 1e-12 * ref.max(axis=1, keepdims=True)
    r = float(np.max(np.abs(got - ref) / tol))
    worst = max(worst, r); per_dim[dim] = max(per_dim.get(dim, 0.0), r)
    multi = eng.eigs_batch_multi(range(5), dim, T, seeds)
    same = np.array_equal(multi[model], got)
    if r > 1.0 or not same:
        bad += 1
        print("FAIL", model, dim, T, seeds.tolist(), r, same, flush=True)
print(f"{ncase} random cases (dim 1..15, T up to 40000, models 0-4, 3 random u32 seeds each): failures {bad}, worst err/tol {worst:.3g}, {time.time()-t0:.0f} s")
print("worst err/tol by dim:", {d: float(f"{v:.2g}") for d, v in sorted(per_dim.items())})

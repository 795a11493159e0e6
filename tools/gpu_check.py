"""Quick GPU-side parity probe (development aid; the real checks live in tests/)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
from oracle import johansen_oracle as orc
from oracle import philox_ref

eng = jne.Engine([0])
print(jne.version())
z = eng.gen_normal_matrix(12, 103, 7)
zr = philox_ref.normal_matrix(12, 103, 7)
print("normals max abs diff vs numpy philox ref:", np.abs(z - zr).max(), "mean", z.mean(), "std", z.std())
rng = np.random.default_rng(1)
worst = 0
for model in range(5):
    for dim in (1, 2, 3, 5, 8, 9, 12, 13, 15):
        for T in (7, 103, 1000):
            if model == 4 and T < 3: continue
            n = 6
            db = rng.standard_normal((n, T, dim)) / np.sqrt(T)
            try:
                got = eng.eigs_from_increments(model, db)
            except jne.JneError as e:
                print("ERR", model, dim, T, e); continue
            ref = orc.eigs_batch_from_increments(db, model)
            tol = 1e-9 * np.abs(ref) + 1e-12 * ref.max(axis=1, keepdims=True)
            err = np.abs(got - ref) / tol
            worst = max(worst, err.max())
            if err.max() > 1:
                print("MISMATCH model", model, "dim", dim, "T", T, "max err/tol", err.max())
                print(got[0]); print(ref[0])
print("increments parity worst err/tol:", worst)
# rng path vs oracle fed with device normals
worst = 0
for model in range(5):
    for dim in (1, 4, 12):
        T = 200
        seeds = np.arange(1, 5, dtype=np.uint32)
        got = eng.eigs_batch(model, dim, T, seeds)
        for i, s in enumerate(seeds):
            zz = eng.gen_normal_matrix(dim, T, int(s))
            ref = orc.eigs_from_normals(zz, model)
            tol = 1e-9 * np.abs(ref) + 1e-12 * ref.max()
            worst = max(worst, (np.abs(got[i] - ref) / tol).max())
print("rng-path pathwise parity worst err/tol:", worst)
for model in (0, 4):
    n = 1 << 18
    seeds = np.arange(1, n + 1, dtype=np.uint32)
    eng.eigs_batch(model, 12, 1000, seeds[:1024])
    t0 = time.time(); out = eng.eigs_batch(model, 12, 10000, seeds); dt = time.time() - t0
    fl = jne.flops_per_run(model, 12, 10000)
    print(f"model {model}: {n/dt:.0f} runs/s e2e, {n/dt*fl/1e12:.2f} TFLOP/s; trace mean {out.sum(1).mean():.3f}")
print("peak dfma", eng.fp64_peak_tflops(0), "dmma", eng.fp64_peak_tflops(1))

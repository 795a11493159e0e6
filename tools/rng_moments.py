#!/usr/bin/env python
"""Moments and tail quantiles of the device normal stream (GPU box): the production transform (FP32 Box-Muller on MUFU,
32-bit uniforms) against the validation build's FP64 transform on 64-bit uniforms (jne_rng.cuh, -DJNE_RNG_F64) on the
SAME uniform words, element by element, plus both against the exact N(0,1) values.
    python tools/rng_moments.py [--n-log2 28] > profiles/r2_rng_moments.txt
The two streams share their leading 32 bits, so the differences of the sample moments are paired: their Monte Carlo
error is that of z32 - z64 (~1e-6 per element), not that of the moments themselves."""
import argparse, os, subprocess, sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

WORKER = r'''
import sys, numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0])
n_log2, out = int(sys.argv[1]), sys.argv[2]
dim, steps = 8, 1 << 21                       # 2^24 normals per call
calls = max(1, (1 << n_log2) // (dim * steps))
acc = np.zeros(8); absmax = 0.0
qs = np.array([0.5, 0.9, 0.99, 0.999, 0.9999, 0.99999, 0.999999])
zs = []
for c in range(calls):
    z = eng.gen_normal_matrix(dim, steps, 1000 + c).ravel()
    zs.append(z.astype(np.float64))
z = np.concatenate(zs)
np.save(out, z)
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-log2", type=int, default=27)
    a = ap.parse_args()
    import importlib.util
    spec = importlib.util.spec_from_file_location("jne_build", ROOT / "johansen_null_eigenspectra_b200" / "build.py")
    jb = importlib.util.module_from_spec(spec); spec.loader.exec_module(jb)
    f64_lib = jb.build_variant("rng_f64")
    tmp = Path("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    z = {}
    for tag, lib in (("fp32", jb.LIB_PATH), ("f64", f64_lib)):
        f = tmp / f"jne_rngmom_{tag}_{os.getpid()}.npy"
        r = subprocess.run([sys.executable, "-c", WORKER, str(a.n_log2), str(f)], cwd=ROOT, env=dict(os.environ, JNE_LIBRARY=str(lib)),
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr[-3000:])
        z[tag] = np.load(f); f.unlink()
    x, y = z["fp32"], z["f64"]
    n = x.size
    from scipy import stats
    print(f"# {n} normals (2^{np.log2(n):.0f}), seeds 1000.., dim 8: production FP32 transform (x) vs FP64 validation transform (y), same uniform words")
    print(f"max |x - y| = {np.max(np.abs(x - y)):.3e}   rms(x - y) = {np.sqrt(np.mean((x - y) ** 2)):.3e}   max |x| = {np.max(np.abs(x)):.4f}   max |y| = {np.max(np.abs(y)):.4f}")
    print("moment        production        validation        exact   paired difference (x - y)   its standard error")
    for name, f, exact in (("E[z]", lambda v: v, 0.0), ("E[z^2]", lambda v: v * v, 1.0), ("E[z^3]", lambda v: v ** 3, 0.0), ("E[z^4]", lambda v: v ** 4, 3.0),
                           ("E[z^6]", lambda v: v ** 6, 15.0)):
        fx, fy = f(x), f(y)
        d = fx - fy
        print(f"{name:8s} {fx.mean():+17.10f} {fy.mean():+17.10f} {exact:8.1f}   {d.mean():+.4e}                 {d.std() / np.sqrt(n):.1e}"
              f"   (MC se of the moment itself {fy.std() / np.sqrt(n):.1e})")
    qs = np.array([0.9, 0.99, 0.999, 0.9999, 0.99999, 0.999999])
    qx, qy = np.quantile(np.abs(x), qs), np.quantile(np.abs(y), qs)
    qe = stats.norm.ppf(0.5 + qs / 2)
    print("quantiles of |z|:   q        production   validation   exact N(0,1)")
    for q, a_, b_, e_ in zip(qs, qx, qy, qe):
        print(f"               {q:10.6f} {a_:12.6f} {b_:12.6f} {e_:12.6f}")
    ks = stats.kstest(x[: 1 << 24], "norm")
    print(f"one-sample KS of the production stream against N(0,1) on 2^24 values: D = {ks.statistic:.2e}, p = {ks.pvalue:.3f}")


if __name__ == "__main__":
    main()

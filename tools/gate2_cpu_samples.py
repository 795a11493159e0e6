#!/usr/bin/env python
"""Gate-(2) CPU side: the trace / max-eig distributions of the reference path driven by f64 normals.

TEST INFRASTRUCTURE (uses oracle/).  For seeds 1..n the C restatement draws one Brownian path per seed from
f64 ziggurat normals (xoshiro256++, oracle/jne_oracle.c: jne_oracle_fast_multi_stats), evaluates all five models
on it and keeps trace = sum(lambda) and max = lambda_1 (src/simulation_analyzers.rs:25-40).  The 2 x 5 samples are
reduced to a grid of K + 1 exact order statistics each (ranks round(j (n - 1) / K)), so that a two-sample
Kolmogorov-Smirnov statistic against a GPU sample can be evaluated without shipping the raw samples:
    D_grid = max_j |F_gpu(g_j) - (rank_j + 1) / n|  <=  D  <=  D_grid + 1 / K.
The file also records mean, variance and the quantiles 0.5 / 0.9 / 0.95 / 0.99 / 0.999 with their Monte Carlo
standard errors (sqrt(q (1 - q) / n) / f(x_q), f from a central difference of the order statistics).

Runs on CPU only; resumable (per-chunk .npy files under --work).  Example (this container, 6 threads, ~70 min):
    python tools/gate2_cpu_samples.py --dim 12 --T 10000 --n 2000000 --threads 6 \
        --out tests/golden/gate2_cpu_dim12_T10000.npz
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

QS = (0.5, 0.9, 0.95, 0.99, 0.999)


def quantile_with_se(sorted_x: np.ndarray, q: float):
    """get_percentile_value of src/simulation_analyzers.rs:4-18 plus a standard error."""
    n = sorted_x.size
    rank = q * (n - 1)
    lo, hi = int(np.floor(rank)), int(np.ceil(rank))
    w = rank - lo
    val = sorted_x[lo] * (1 - w) + sorted_x[hi] * w
    h = max(int(2.0 * np.sqrt(n * q * (1 - q))), 10)            # ~2 binomial sigmas of rank on each side
    a, b = max(lo - h, 0), min(hi + h, n - 1)
    dens = (b - a) / n / max(sorted_x[b] - sorted_x[a], 1e-300)
    return float(val), float(np.sqrt(q * (1 - q) / n) / dens)


def reduce_sample(x: np.ndarray, K: int):
    xs = np.sort(x)
    n = xs.size
    ranks = np.unique(np.rint(np.arange(K + 1) * ((n - 1) / K)).astype(np.int64))
    qv = [quantile_with_se(xs, q) for q in QS]
    return {"grid": xs[ranks], "ranks": ranks, "mean": xs.mean(), "var": xs.var(ddof=1),
            "q": np.array([v for v, _ in qv]), "q_se": np.array([s for _, s in qv])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, required=True)
    ap.add_argument("--T", type=int, required=True)
    ap.add_argument("--n", type=int, required=True)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--chunk", type=int, default=50000)
    ap.add_argument("--K", type=int, default=16384)
    ap.add_argument("--out", required=True)
    ap.add_argument("--first-seed", type=int, default=1, help="seeds first-seed .. first-seed + n - 1")
    ap.add_argument("--work", default="/tmp/jne_gate2_work")
    args = ap.parse_args()

    from oracle import c_oracle
    lib = c_oracle.load()
    work = Path(args.work) / (f"d{args.dim}_T{args.T}" + (f"_s{args.first_seed}" if args.first_seed != 1 else ""))
    work.mkdir(parents=True, exist_ok=True)
    t0 = time.time()
    parts = []
    for a in range(0, args.n, args.chunk):
        b = min(a + args.chunk, args.n)
        f = work / f"chunk_{a}_{b}.npy"
        if not f.exists():
            st = c_oracle.fast_multi_stats(lib, args.dim, args.T, np.arange(args.first_seed + a, args.first_seed + b, dtype=np.uint32), args.threads)
            np.save(f, st)
            print(f"[{time.time() - t0:7.0f} s] seeds {a + 1}..{b} done", flush=True)
        parts.append(np.load(f))
    st = np.concatenate(parts)                       # (n, 5, 2)
    out = {"dim": args.dim, "T": args.T, "n": args.n, "K": args.K, "qs": np.array(QS), "first_seed": args.first_seed,
           "generator": "xoshiro256++ seeded by the run seed, 256-layer ziggurat, f64 (oracle/jne_oracle.c)"}
    for m in range(5):
        for k, name in enumerate(("trace", "max")):
            for key, v in reduce_sample(st[:, m, k], args.K).items():
                out[f"m{m}_{name}_{key}"] = v
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out} ({os.path.getsize(args.out) / 1e6:.2f} MB) in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Parity gate (2) at the BASELINE configurations, full size (GPU box).

  ks   two-sample KS + quantile table of the device stream against the committed CPU grids
       (tests/golden/gate2_cpu_dim*_T*.npz: the C restatement of the reference path on f64 ziggurat normals)
  ab   paired A/B of the production stream (FP32 Box-Muller, 32-bit uniforms, |z| <= 6.76) against the validation
       stream of jne_rng.cuh (-DJNE_RNG_F64: 64-bit uniforms, FP64 transform) on the SAME seeds: the two streams share
       their leading 32 bits, so the statistics are compared run by run, then as distributions.

    python tools/validate_gate2.py ks --n 2000000 > profiles/r2_gate2_ks.txt
    python tools/validate_gate2.py ab --dim 12 --T 10000 --n 10000000 > profiles/r2_gate2_ab_dim12.txt
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def cmd_ks(args):
    import johansen_null_eigenspectra_b200 as jne
    from tests import gate2_common as g2
    eng = jne.Engine([0])
    for dim, T in ((5, 5000), (12, 10000)):
        if not g2.cpu_grid_path(dim, T).exists():
            print(f"# dim {dim} T {T}: CPU grid missing, skipped")
            continue
        t0 = time.time()
        n_cpu = int(np.load(g2.cpu_grid_path(dim, T))["n"])
        rows = g2.compare(eng, dim, T, args.n or n_cpu, first_seed=args.first_seed)
        print(f"# dim {dim}, T {T}: GPU seeds {args.first_seed}..{args.first_seed + rows[0]['n_gpu'] - 1} vs CPU f64-ziggurat sample of {rows[0]['n_cpu']} runs "
              f"({time.time() - t0:.1f} s); KS alpha = {g2.ALPHA}")
        print(g2.format_rows(rows))
        print("# quantiles (GPU | CPU | relative difference), q = " + ", ".join(str(q) for q in g2.QS))
        for r in rows:
            print(f"  m{r['model']} {r['stat']:5s} " + " ".join(f"{v:10.5f}" for v in r["q_gpu"]) + " | " +
                  " ".join(f"{v:10.5f}" for v in r["q_cpu"]) + " | " + " ".join(f"{v:+.1e}" for v in r["rel"]))
        worst = max(rows, key=lambda r: r["D_upper"] / r["D_crit"])
        print(f"# worst D_upper / D_crit = {worst['D_upper'] / worst['D_crit']:.3f} (model {worst['model']} {worst['stat']}); "
              f"all pass: {all(r['p_value'] > g2.ALPHA for r in rows)}; max |z| = {max(float(np.abs(r['z']).max()) for r in rows):.2f}\n")
    eng.close()


def cmd_ab_worker(args):
    """Runs in a subprocess with JNE_LIBRARY pointing at one of the two builds; writes (n, 5, 2) statistics."""
    import torch
    import johansen_null_eigenspectra_b200 as jne
    eng = jne.Engine([0])
    widths = [jne.num_eigs(m, args.dim) for m in range(5)]
    st = torch.cuda.current_stream()
    res = torch.empty((args.n, 5, 2), dtype=torch.float64, device="cuda")
    t0 = time.time()
    chunk = 1 << 19
    for a in range(0, args.n, chunk):
        m = min(chunk, args.n - a)
        seeds = torch.arange(1 + a, 1 + a + m, dtype=torch.int64, device="cuda").to(torch.int32)
        out = torch.empty((m, sum(widths)), dtype=torch.float64, device="cuda")
        eng.eigs_batch_multi_device(range(5), args.dim, args.T, seeds.data_ptr(), m, out.data_ptr(), st.cuda_stream)
        eng.check_async()
        off = 0
        for k, w in enumerate(widths):
            res[a:a + m, k, 0] = out[:, off:off + w].sum(dim=1)
            res[a:a + m, k, 1] = out[:, off]
            off += w
    torch.cuda.synchronize()
    np.save(args.out, res.cpu().numpy())
    print(f"{jne.version()} [{os.environ.get('JNE_LIBRARY', 'libjne.so')}]: {args.n} seeds in {time.time() - t0:.1f} s", flush=True)
    eng.close()


def cmd_ab(args):
    from scipy import stats
    import importlib.util
    spec = importlib.util.spec_from_file_location("jne_build", ROOT / "johansen_null_eigenspectra_b200" / "build.py")
    jb = importlib.util.module_from_spec(spec); spec.loader.exec_module(jb)
    f64_lib = jb.build_variant("rng_f64")
    tmp = Path("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    files = {}
    for tag, lib in (("fp32", jb.LIB_PATH), ("f64", f64_lib)):
        files[tag] = tmp / f"jne_ab_{tag}_{os.getpid()}.npy"
        env = dict(os.environ, JNE_LIBRARY=str(lib))
        r = subprocess.run([sys.executable, __file__, "ab-worker", "--dim", str(args.dim), "--T", str(args.T), "--n", str(args.n),
                            "--out", str(files[tag])], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr[-3000:])
        print("# " + r.stdout.strip())
    a, b = np.load(files["fp32"]), np.load(files["f64"])
    for f in files.values():
        f.unlink()
    qs = (0.5, 0.9, 0.95, 0.99, 0.999, 0.9999)
    print(f"# paired A/B, dim {args.dim}, T {args.T}, seeds 1..{args.n}: production stream (FP32 Box-Muller, 32-bit uniforms) vs "
          "validation stream (64-bit uniforms, FP64 transform)")
    print("model stat  max|d|/value  rms(d)/sd   mean rel diff   KS D       p       | relative quantile differences q50 q90 q95 q99 q99.9 q99.99")
    for m in range(5):
        for k, name in enumerate(("trace", "max")):
            x, y = a[:, m, k], b[:, m, k]
            d = x - y
            ks = stats.ks_2samp(x, y)
            qx, qy = np.quantile(x, qs), np.quantile(y, qs)
            print(f"{m:^5d} {name:5s} {np.max(np.abs(d) / y):.3e}    {np.sqrt(np.mean(d * d)) / y.std():.3e}   "
                  f"{x.mean() / y.mean() - 1:+.3e}     {ks.statistic:.2e}  {ks.pvalue:.4f}  | " +
                  " ".join(f"{v:+.1e}" for v in qx / qy - 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["ks", "ab", "ab-worker"])
    ap.add_argument("--dim", type=int, default=12)
    ap.add_argument("--T", type=int, default=10000)
    ap.add_argument("--n", type=int, default=0, help="GPU seeds (ks: 0 = as many as the CPU sample; ab: required)")
    ap.add_argument("--out", default="")
    ap.add_argument("--first-seed", type=int, default=1)
    args = ap.parse_args()
    {"ks": cmd_ks, "ab": cmd_ab, "ab-worker": cmd_ab_worker}[args.cmd](args)


if __name__ == "__main__":
    main()

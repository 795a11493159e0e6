"""Parity gate (2) helpers shared by tests/test_gate2_gpu.py and tools/validate_gate2.py.

BASELINE.json north_star: "GPU-RNG runs match the reference's trace and max-eigenvalue quantiles within Monte Carlo
error (two-sample KS at stated alpha)".  The CPU side (f64 ziggurat normals, the C restatement of the reference path)
is sampled once by tools/gate2_cpu_samples.py and committed as a grid of exact order statistics
(tests/golden/gate2_cpu_dim*_T*.npz); the GPU side is produced here, on the device, through the C ABI.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
ALPHA = 1e-3
QS = (0.5, 0.9, 0.95, 0.99, 0.999)

# MacKinnon-Haug-Michelis (1999) asymptotic 95 % critical values, cases I-V <-> models 0-4, dim 1..12
MHM95_TRACE = {
    0: [4.129906, 12.32090, 24.27596, 40.17493, 60.06141, 83.93712, 111.7805, 143.6691, 179.5098, 219.4016, 263.2603, 311.1288],
    1: [9.164546, 20.26184, 35.19275, 54.07904, 76.97277, 103.8473, 134.6780, 169.5991, 208.4374, 251.2650, 298.1594, 348.9784],
    2: [3.841466, 15.49471, 29.79707, 47.85613, 69.81889, 95.75366, 125.6154, 159.5297, 197.3709, 239.2354, 285.1425, 334.9837],
    3: [12.51798, 25.87211, 42.91525, 63.87610, 88.80380, 117.7082, 150.5585, 187.4701, 228.2979, 273.1889, 322.0692, 374.9076],
    4: [3.841466, 18.39771, 35.01090, 55.24578, 79.34145, 107.3466, 139.2753, 175.1715, 215.1232, 259.0294, 306.8944, 358.7184],
}
MHM95_MAX = {
    0: [4.129906, 11.22480, 17.79730, 24.15921, 30.43961, 36.63019, 42.77219, 48.87720, 54.96577, 61.03407, 67.07555, 73.09094],
    1: [9.164546, 15.89210, 22.29962, 28.58808, 34.80587, 40.95680, 47.07897, 53.18784, 59.24000, 65.30016, 71.33542, 77.38180],
    2: [3.841466, 14.26460, 21.13162, 27.58434, 33.87687, 40.07757, 46.23142, 52.36261, 58.43354, 64.50472, 70.53513, 76.57843],
    3: [12.51798, 19.38704, 25.82321, 32.11832, 38.33101, 44.49720, 50.59985, 56.70519, 62.75215, 68.81206, 74.83748, 80.87025],
    4: [3.841466, 17.14769, 24.25202, 30.81507, 37.16359, 43.41977, 49.58633, 55.72819, 61.80550, 67.90393, 73.94036, 79.97193],
}


def cpu_grid_path(dim: int, T: int) -> Path:
    return GOLDEN / f"gate2_cpu_dim{dim}_T{T}.npz"


def gpu_statistics(engine, dim: int, T: int, n: int, first_seed: int = 1, chunk: int = 1 << 20):
    """trace and max-eig of seeds first_seed .. first_seed + n - 1 for all five models from the fused pass, reduced and
    SORTED on the device.  Returns {model: (trace_sorted, max_sorted)} as torch float64 CUDA tensors."""
    import torch
    import johansen_null_eigenspectra_b200 as jne
    widths = [jne.num_eigs(m, dim) for m in range(5)]
    st = torch.cuda.current_stream()
    tr = torch.empty((5, n), dtype=torch.float64, device="cuda")
    mx = torch.empty((5, n), dtype=torch.float64, device="cuda")
    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        seeds = torch.arange(first_seed + a, first_seed + a + m, dtype=torch.int64, device="cuda").to(torch.int32)
        out = torch.empty((m, sum(widths)), dtype=torch.float64, device="cuda")
        engine.eigs_batch_multi_device(range(5), dim, T, seeds.data_ptr(), m, out.data_ptr(), st.cuda_stream)
        engine.check_async()
        off = 0
        for k, w in enumerate(widths):
            tr[k, a:a + m] = out[:, off:off + w].sum(dim=1)
            mx[k, a:a + m] = out[:, off]              # rows are descending (src/johansen_statistics.rs:45)
            off += w
    return {k: (torch.sort(tr[k]).values, torch.sort(mx[k]).values) for k in range(5)}


def ks_against_grid(sorted_gpu, grid: np.ndarray, ranks: np.ndarray, n_cpu: int):
    """Two-sample KS of a sorted GPU sample (torch CUDA tensor) against the CPU sample represented by its order
    statistics grid[j] = x_(ranks[j]).  Returns (D_lower, D_upper, p_value of D_upper, effective n)."""
    import torch
    from scipy import stats
    n_gpu = sorted_gpu.numel()
    g = torch.from_numpy(np.ascontiguousarray(grid)).to(sorted_gpu.device)
    # ECDFs just right of each grid point (<=) and just left of it (<)
    f_gpu_hi = torch.searchsorted(sorted_gpu, g, right=True).double() / n_gpu
    f_gpu_lo = torch.searchsorted(sorted_gpu, g, right=False).double() / n_gpu
    r = torch.from_numpy(ranks.astype(np.float64)).to(sorted_gpu.device)
    f_cpu_hi = (r + 1.0) / n_cpu                        # CPU ECDF at its own order statistic (ties have measure zero)
    f_cpu_lo = r / n_cpu
    d_lower = float(torch.maximum((f_gpu_hi - f_cpu_hi).abs().max(), (f_gpu_lo - f_cpu_lo).abs().max()))
    gap = float(np.max(np.diff(ranks))) / n_cpu          # between grid points the CPU ECDF moves by at most this much
    d_upper = d_lower + gap
    ne = n_gpu * n_cpu / (n_gpu + n_cpu)
    return d_lower, d_upper, float(stats.kstwobign.sf(d_upper * np.sqrt(ne))), ne


def quantile_of_sorted(sorted_gpu, q: float) -> float:
    """get_percentile_value of src/simulation_analyzers.rs:4-18 on a sorted device sample."""
    n = sorted_gpu.numel()
    rank = q * (n - 1)
    lo, hi = int(np.floor(rank)), int(np.ceil(rank))
    w = rank - lo
    return float(sorted_gpu[lo]) * (1 - w) + float(sorted_gpu[hi]) * w


def compare(engine, dim: int, T: int, n_gpu: int, first_seed: int = 1):
    """One row per (model, statistic): KS of the GPU sample against the committed CPU grid and the quantile
    differences in units of their Monte Carlo standard error."""
    ref = np.load(cpu_grid_path(dim, T))
    n_cpu = int(ref["n"])
    stats_gpu = gpu_statistics(engine, dim, T, n_gpu, first_seed)
    rows = []
    for m in range(5):
        for k, name in enumerate(("trace", "max")):
            sg = stats_gpu[m][k]
            d_lo, d_up, p, ne = ks_against_grid(sg, ref[f"m{m}_{name}_grid"], ref[f"m{m}_{name}_ranks"], n_cpu)
            qc, se_c = ref[f"m{m}_{name}_q"], ref[f"m{m}_{name}_q_se"]
            qg = np.array([quantile_of_sorted(sg, q) for q in QS])
            se = se_c * np.sqrt(1.0 + n_cpu / n_gpu)    # the GPU sample has the same density: its SE scales with 1/sqrt(n)
            mhm = (MHM95_TRACE if name == "trace" else MHM95_MAX)[m][dim - 1] if dim <= 12 else float("nan")
            rows.append({
                "model": m, "stat": name, "n_gpu": n_gpu, "n_cpu": n_cpu, "D_lower": d_lo, "D_upper": d_up, "p_value": p,
                "D_crit": float(np.sqrt(-0.5 * np.log(ALPHA / 2.0)) / np.sqrt(ne)),
                "q_gpu": qg, "q_cpu": qc, "z": (qg - qc) / se, "rel": qg / qc - 1.0,
                "mean_gpu": float(sg.mean()), "mean_cpu": float(ref[f"m{m}_{name}_mean"]),
                "q95_vs_mhm_gpu": qg[2] / mhm - 1.0, "q95_vs_mhm_cpu": qc[2] / mhm - 1.0,
                "q95_se_rel": float(se_c[2] / qc[2]),
            })
    return rows


def format_rows(rows) -> str:
    out = ["model stat   n_gpu     n_cpu     D_lower   D_upper   D_crit    p(D_up)  | z-scores of q50 q90 q95 q99 q99.9 |"
           " q95/MHM-1: gpu      cpu      (MC se)"]
    for r in rows:
        z = " ".join(f"{v:+5.2f}" for v in r["z"])
        out.append(f"{r['model']:^5d} {r['stat']:5s} {r['n_gpu']:9d} {r['n_cpu']:9d} {r['D_lower']:.3e} {r['D_upper']:.3e} "
                   f"{r['D_crit']:.3e} {r['p_value']:8.4f} | {z} | {r['q95_vs_mhm_gpu']:+.5f} {r['q95_vs_mhm_cpu']:+.5f} "
                   f"({r['q95_se_rel']:.5f})")
    return "\n".join(out)

"""Parity gate (2) at the BASELINE configurations (-m gpu): the trace / max-eig distributions of the device stream
(stream JNE2: Philox-keyed xoshiro128++ substreams, FP32 Box-Muller) against the C restatement of the reference path driven by f64 ziggurat normals
(src/rng_matrix.rs:26-34), at c2 (dim 5, T 5 000) and c4 (dim 12, T 10 000), all five models.

CPU side: tests/golden/gate2_cpu_dim*_T*.npz, made by tools/gate2_cpu_samples.py (4 * 10^6 runs at c2, 10^7 at c4; order-
statistics grid of 16 385 points, so the KS statistic is bracketed to 6e-5).  GPU side: as many seeds as the CPU side
through jne_eigs_batch_multi_device, reduced and sorted on the device.  Stated alpha = 1e-3 per statistic for the KS test
(D_upper, the conservative end of the bracket, is the one tested); the five quantiles 0.5 .. 0.999 must agree within
4.5 Monte Carlo standard errors of the difference.
"""
import numpy as np
import pytest

from tests import gate2_common as g2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,T,n_gpu", [(5, 5000, 4_000_000), (12, 10000, 10_000_000)])
def test_ks_and_quantiles_vs_cpu_f64_normals(engine, dim, T, n_gpu):
    assert g2.cpu_grid_path(dim, T).exists(), "CPU grid missing: run tools/gate2_cpu_samples.py"
    rows = g2.compare(engine, dim, T, n_gpu=n_gpu)
    print("\n" + g2.format_rows(rows))
    for r in rows:
        what = f"dim {dim} T {T} model {r['model']} {r['stat']}"
        assert r["p_value"] > g2.ALPHA, f"KS rejects at alpha = {g2.ALPHA}: {what}: D in [{r['D_lower']:.3e}, {r['D_upper']:.3e}]"
        assert np.all(np.abs(r["z"]) < 4.5), f"quantiles differ: {what}: z = {r['z']}"
        # both sides carry the same finite-T bias against the asymptotic table (a few 1e-3 at T = 10^4)
        assert abs(r["q95_vs_mhm_gpu"] - r["q95_vs_mhm_cpu"]) < 6.0 * r["q95_se_rel"], what

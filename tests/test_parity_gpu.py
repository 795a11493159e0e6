"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libjne.so), against the
CPU oracle on the same seeded inputs and against the committed golden fixtures.  (-m gpu)

Tolerances (BASELINE.json north_star / BASELINE.md section 5):
  gate (1) shared increments: |d| <= 1e-9 |lambda| + 1e-12 lambda_max   (eig_tol)
  gate (2) GPU RNG: two-sample KS at alpha = 1e-3 on trace and max-eig; chi^2(1) law; MHM quantiles
"""
import numpy as np
import pytest
from scipy import stats

from oracle import c_oracle, johansen_oracle as orc, philox_ref
from tests.conftest import eig_tol
from tests.golden import make_golden as mg

pytestmark = pytest.mark.gpu


def assert_close(got, ref, what=""):
    tol = eig_tol(ref)
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{what}: max err/tol {np.max(np.abs(got - ref) / tol):.3g}"


# ---- gate (1): shared increments ------------------------------------------------------------
def test_golden_fixtures(engine):
    g = np.load(mg.HERE / "eigs_from_increments.npz")
    for key in g.files:
        m, d, t = (int(s[1:]) for s in key.split("_"))
        db = mg.increments(mg.case_seed(m, d, t), mg.N, t, d)
        assert_close(engine.eigs_from_increments(m, db), g[key], key)


def test_golden_full_size(engine):
    """dim 12, T 10 000 (the metric's configuration), all five models."""
    g = np.load(mg.HERE / "full_dim12_steps10000.npz")
    for m in range(5):
        db = mg.increments(mg.case_seed(m, 12, 10000), 2, 10000, 12)
        assert_close(engine.eigs_from_increments(m, db), g[f"m{m}"], f"model {m}")


def test_config_c1_shared_increments(engine):
    """BASELINE.json configs[0]: model 0, dim 2, T 1000, 100 000 runs on shared increments
    (chunked; the C port of the oracle is the checker, itself pinned to the numpy oracle)."""
    lib = c_oracle.load()
    c1 = np.load(mg.HERE / "c1_model0_dim2_steps1000.npz")["eigs"]
    rng = np.random.default_rng(mg.SEED)
    worst = 0.0
    for chunk in range(10):
        db = rng.standard_normal((10000, 1000, 2)) * np.sqrt(1.0 / 1000)
        got = engine.eigs_from_increments(0, db)
        if chunk == 0:
            assert_close(got[:64], c1, "c1 golden")
        ref = np.stack([c_oracle.eigs_from_increments(lib, db[i], 0) for i in range(db.shape[0])])
        worst = max(worst, np.max(np.abs(got - ref) / eig_tol(ref)))
    assert worst <= 1.0, worst


@pytest.mark.parametrize("model", range(5))
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 7, 8, 9, 11, 12, 13, 15])
def test_all_dims_vs_oracle(engine, model, dim):
    rng = np.random.default_rng(1000 * model + dim)
    for T in (dim + 5, 64, 257, 1001):
        db = rng.standard_normal((5, T, dim)) / np.sqrt(T)
        assert_close(engine.eigs_from_increments(model, db), orc.eigs_batch_from_increments(db, model),
                     f"model {model} dim {dim} T {T}")


def test_ragged_step_counts(engine):
    """Segment stitching: every T mod 16 residue, incl. T smaller than one segment block."""
    rng = np.random.default_rng(77)
    for T in list(range(8, 41)) + [99, 100, 101, 102, 103]:
        for model in (0, 3, 4):
            db = rng.standard_normal((3, T, 3)) / np.sqrt(T)
            assert_close(engine.eigs_from_increments(model, db), orc.eigs_batch_from_increments(db, model),
                         f"model {model} T {T}")


def test_unscaled_increments_and_long_horizon(engine):
    """Eigenvalues are invariant to the increment scale only through factor = T; check a
    non-unit variance input and T = 100 000 (config c5: accumulation precision)."""
    rng = np.random.default_rng(5)
    db = rng.standard_normal((2, 500, 4)) * 3.7
    assert_close(engine.eigs_from_increments(2, db), orc.eigs_batch_from_increments(db, 2))
    for model in (3, 4):
        db = rng.standard_normal((1, 100000, 12)) / np.sqrt(100000)
        assert_close(engine.eigs_from_increments(model, db), orc.eigs_batch_from_increments(db, model),
                     f"T=1e5 model {model}")


# ---- the eigen-solve alone (reference a8: GeneralizedEigen::new -> dggev) ---------------------
def test_pencil_solver_vs_dggev(engine):
    from scipy.linalg import lapack
    rng = np.random.default_rng(9)
    for p, d in [(1, 1), (2, 2), (3, 2), (5, 5), (12, 12), (13, 12), (16, 15), (16, 16)]:
        n = 40
        S1 = rng.standard_normal((n, d, p))
        F = rng.standard_normal((n, p, 3 * p + 2))
        S2 = F @ np.transpose(F, (0, 2, 1)) / (3 * p)
        got = engine.pencil_eigs_batch(S1, S2)
        for i in range(n):
            ar, ai, be, *_ = lapack.dggev(np.asfortranarray(S1[i].T @ S1[i]), np.asfortranarray(S2[i]))
            ref = np.sort(np.hypot(ar, ai) / be)[::-1]
            assert_close(got[i], ref, f"p {p} d {d}")


# ---- gate (2): the GPU random stream -----------------------------------------------------------
def test_device_normals_match_philox_restatement(engine):
    """Uniform words are bit-exact; MUFU lg2/sqrt/sin/cos leave <= 5e-6 absolute on the normals."""
    for dim, T, seed in [(12, 103, 7), (1, 1, 1), (5, 4001, 0xFFFFFFFF), (15, 64, 123456789)]:
        z = engine.gen_normal_matrix(dim, T, seed)
        assert z.shape == (dim, T)
        assert np.abs(z - philox_ref.normal_matrix(dim, T, seed)).max() < 5e-6


def test_reference_rng_tests_on_device_stream(engine):
    import johansen_null_eigenspectra_b200 as jne
    # src/tests/rng_matrix_test/gen_normal_matrix_test.rs:7-16 -- 200 x 300 normals, seed 42, CDF within 1e-2
    z = engine.gen_normal_matrix(200, 300, 42)
    qs = np.arange(1, 100) / 100
    emp = np.searchsorted(np.sort(z.ravel()), stats.norm.ppf(qs)) / z.size
    assert np.abs(emp - qs).max() < 1e-2
    # brownian_motion_test.rs:9-44 shape dim x (steps+1); :47-96 increments / sqrt(dt) normal within 0.05
    dim, steps, dt = 3, 1000, 0.01
    bm = engine.brownian_motion_matrix(dim, steps, dt, 42)
    assert bm.shape == (dim, steps + 1) and np.all(bm[:, 0] == 0.0)
    inc = np.diff(bm, axis=1) / np.sqrt(dt)
    emp = np.searchsorted(np.sort(inc.ravel()), stats.norm.ppf(qs)) / inc.size
    assert np.abs(emp - qs).max() < 0.05
    # :99-125 same seed => identical; :128-153 different seeds => different
    assert np.array_equal(bm, engine.brownian_motion_matrix(dim, steps, dt, 42))
    assert not np.array_equal(bm, engine.brownian_motion_matrix(dim, steps, dt, 43))
    # cumulative sum of sqrt(dt) z, naive left to right (src/matrix_utils.rs:51-63)
    z = engine.gen_normal_matrix(dim, steps, 42)
    assert np.array_equal(bm[:, 1:], np.cumsum(z * np.sqrt(dt), axis=1))


@pytest.mark.parametrize("model", range(5))
def test_rng_path_pathwise_vs_oracle(engine, model):
    """jne_eigs_batch == oracle(reference algorithm) fed the SAME normals the device stream produces:
    checks the fused in-register generation + accumulation end to end, to gate-(1) tolerance."""
    cases = [(1, 50), (2, 103), (5, 1000), (12, 400), (12, 10000), (15, 257)]
    if model in (3, 4):
        cases.append((12, 100000))       # config c5: long horizon (accumulation precision of the one-pass moments)
    for dim, T in cases:
        seeds = np.array([1, 2, 4294967295], dtype=np.uint32)
        got = engine.eigs_batch(model, dim, T, seeds)
        ref = np.stack([orc.eigs_from_normals(engine.gen_normal_matrix(dim, T, int(s)), model) for s in seeds])
        assert_close(got, ref, f"model {model} dim {dim} T {T}")


@pytest.mark.parametrize("model", [2, 3, 4])
def test_trend_weight_kernels_vs_oracle(engine, model):
    """The AUX kernels (dim <= 6 and 9..12: trend moments through the tensor pipe, table-driven weights) against the
    oracle on the device normals, for every dim they serve (and dim 7 as a scalar-sum neighbour), ragged step counts
    (masked tail blocks, empty segments),
    and both sides of the table-size limit (T = 2^22 uses the table, 2^22 + 1 falls back to the scalar sums)."""
    # (T >= 2 dim + 6: with fewer steps than rows F F' is singular and the reference itself fails)
    cases = [(d, T) for d in (1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12) for T in (31, 33, 67, 250)] + [(5, 5000), (6, 1001)]
    cases += [(d, 14) for d in (1, 2, 3, 4)] + [(12, 2999), (4, 10000)]      # T = 14: two empty segments
    for dim, T in cases:
        seeds = np.array([7, 8], dtype=np.uint32)
        got = engine.eigs_batch(model, dim, T, seeds)
        ref = np.stack([orc.eigs_from_normals(engine.gen_normal_matrix(dim, T, int(s)), model) for s in seeds])
        assert_close(got, ref, f"model {model} dim {dim} T {T}")
        multi = engine.eigs_batch_multi(range(5), dim, T, seeds)
        assert np.array_equal(multi[model], got), f"fused pass differs from the single-model kernel: model {model} dim {dim} T {T}"
    for T in ((1 << 22), (1 << 22) + 1):
        seeds = np.array([11], dtype=np.uint32)
        got = engine.eigs_batch(model, 2, T, seeds)
        ref = np.stack([orc.eigs_from_normals(engine.gen_normal_matrix(2, T, int(s)), model) for s in seeds])
        tol = 1e-8 * np.abs(ref) + 1e-11 * ref.max()      # T = 4.2e6: the oracle's own two-pass sums carry ~1e-10
        assert np.all(np.abs(got - ref) <= tol), f"model {model} T {T}: {np.max(np.abs(got - ref) / tol):.3g}"


def test_determinism_and_geometry_independence(engine):
    """Same (model, dim, steps, seed) => bit-identical record, whatever the batch, order or entry point
    (stronger than src/tests/data_storage/integration/resumable.rs:62-70)."""
    import torch
    seeds = np.arange(1, 3001, dtype=np.uint32)
    a = engine.eigs_batch(3, 5, 300, seeds)
    perm = np.random.default_rng(0).permutation(seeds.size)
    b = engine.eigs_batch(3, 5, 300, seeds[perm])
    assert np.array_equal(a[perm], b)
    c = np.concatenate([engine.eigs_batch(3, 5, 300, seeds[i:i + 7]) for i in range(0, 70, 7)])
    assert np.array_equal(a[:70], c)
    # resume-like subset (src/data_storage/progress.rs:56-61): arbitrary missing seeds
    sub = seeds[[1, 3, 500, 2999]]
    assert np.array_equal(engine.eigs_batch(3, 5, 300, sub), a[[1, 3, 500, 2999]])
    # async pair and device-pointer entry give the same bits
    out = np.empty_like(a)
    engine.wait(engine.submit(3, 5, 300, seeds, out))
    assert np.array_equal(out, a)
    ds = torch.from_numpy(seeds.astype(np.int64)).to(torch.int32).cuda().contiguous()   # same 32 bits
    do = torch.empty((seeds.size, 6), dtype=torch.float64, device="cuda")
    engine.eigs_batch_device(3, 5, 300, ds.data_ptr(), seeds.size, do.data_ptr(), torch.cuda.current_stream().cuda_stream)
    engine.check_async()
    assert np.array_equal(do.cpu().numpy(), a)
    # the Brownian path does not depend on the model (common random numbers, src/rng_matrix.rs:11)
    m0 = engine.eigs_batch(0, 4, 200, seeds[:50])
    _, S2, _ = engine.eigs_batch_debug(0, 4, 200, seeds[:50])
    _, S2b, _ = engine.eigs_batch_debug(1, 4, 200, seeds[:50])
    assert np.allclose(S2[:, :4, :4], S2b[:, :4, :4], rtol=1e-14, atol=0)
    assert m0.shape == (50, 4)


@pytest.mark.parametrize("model", range(5))
def test_ks_vs_oracle_with_independent_generator(engine, model):
    """Gate (2): trace and max-eig of GPU-RNG runs vs the oracle driven by numpy's PCG64;
    two-sample KS, alpha = 1e-3 per statistic (dim 3, T 200, 4000 vs 4000 runs)."""
    dim, T, n = 3, 200, 4000
    gpu = engine.eigs_batch(model, dim, T, np.arange(1, n + 1, dtype=np.uint32))
    rng = np.random.default_rng(31337 + model)
    cpu = np.stack([orc.eigs_from_normals(rng.standard_normal((dim, T)), model) for _ in range(n)])
    assert stats.ks_2samp(gpu.sum(axis=1), cpu.sum(axis=1)).pvalue > 1e-3
    assert stats.ks_2samp(gpu[:, 0], cpu[:, 0]).pvalue > 1e-3


@pytest.mark.parametrize("model", [2, 4])
def test_chi2_law_dim1_gpu(engine, model):
    """Exact law: models 2 and 4 at dim 1 give lambda ~ chi^2(1) for every T."""
    ev = engine.eigs_batch(model, 1, 1000, np.arange(1, 200001, dtype=np.uint32))[:, 0]
    assert stats.kstest(ev, stats.chi2(1).cdf).pvalue > 1e-3
    assert abs(np.quantile(ev, 0.95) - 3.841) < 0.05


def test_mhm_critical_values_gpu(engine):
    """95 % quantiles vs MacKinnon-Haug-Michelis, T = 2000, 200 000 runs:
    MC error ~0.3 %, finite-T bias < 1 %; 2.5 % window."""
    import johansen_null_eigenspectra_b200 as jne
    seeds = np.arange(1, 200001, dtype=np.uint32)
    table = {(0, 2): (12.32, 11.22), (1, 2): (20.26, 15.89), (2, 2): (15.49, 14.26),
             (3, 2): (25.87, 19.39), (4, 2): (18.40, 17.15), (0, 3): (24.28, 17.80)}
    for (model, dim), (tr, mx) in table.items():
        ev = engine.eigs_batch(model, dim, 2000, seeds)
        assert abs(orc.percentiles(ev.sum(axis=1), (0.95,))[0] - tr) / tr < 0.025, (model, dim)
        assert abs(orc.percentiles(ev[:, 0], (0.95,))[0] - mx) / mx < 0.025, (model, dim)


def test_trace_critical_values_all_dims(engine):
    """The product's purpose end to end: the 95 % quantiles of the trace and maximum-eigenvalue statistics from the fused
    pass at the metric's horizon (T = 10 000), models 0-4, dim 1..12, against the published asymptotic critical values of
    MacKinnon, Haug & Michelis (1999) (cases I-V, as printed by standard econometrics packages).  200 000 runs per cell: Monte Carlo
    error ~0.15 %, finite-T bias up to -0.2 % (profiles/r1_validation_trace_quantiles.txt has the 10^6-run table:
    every cell within 0.19 %); window 0.7 %."""
    import torch
    import johansen_null_eigenspectra_b200 as jne
    mhm95 = {
        0: [4.129906, 12.32090, 24.27596, 40.17493, 60.06141, 83.93712, 111.7805, 143.6691, 179.5098, 219.4016, 263.2603, 311.1288],
        1: [9.164546, 20.26184, 35.19275, 54.07904, 76.97277, 103.8473, 134.6780, 169.5991, 208.4374, 251.2650, 298.1594, 348.9784],
        2: [3.841466, 15.49471, 29.79707, 47.85613, 69.81889, 95.75366, 125.6154, 159.5297, 197.3709, 239.2354, 285.1425, 334.9837],
        3: [12.51798, 25.87211, 42.91525, 63.87610, 88.80380, 117.7082, 150.5585, 187.4701, 228.2979, 273.1889, 322.0692, 374.9076],
        4: [3.841466, 18.39771, 35.01090, 55.24578, 79.34145, 107.3466, 139.2753, 175.1715, 215.1232, 259.0294, 306.8944, 358.7184],
    }
    mhm95_max = {
        0: [4.129906, 11.22480, 17.79730, 24.15921, 30.43961, 36.63019, 42.77219, 48.87720, 54.96577, 61.03407, 67.07555, 73.09094],
        1: [9.164546, 15.89210, 22.29962, 28.58808, 34.80587, 40.95680, 47.07897, 53.18784, 59.24000, 65.30016, 71.33542, 77.38180],
        2: [3.841466, 14.26460, 21.13162, 27.58434, 33.87687, 40.07757, 46.23142, 52.36261, 58.43354, 64.50472, 70.53513, 76.57843],
        3: [12.51798, 19.38704, 25.82321, 32.11832, 38.33101, 44.49720, 50.59985, 56.70519, 62.75215, 68.81206, 74.83748, 80.87025],
        4: [3.841466, 17.14769, 24.25202, 30.81507, 37.16359, 43.41977, 49.58633, 55.72819, 61.80550, 67.90393, 73.94036, 79.97193],
    }
    n, T = 200_000, 10_000
    st = torch.cuda.current_stream()
    seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
    for dim in range(1, 13):
        widths = [jne.num_eigs(m, dim) for m in range(5)]
        out = torch.empty((n, sum(widths)), dtype=torch.float64, device="cuda")
        engine.eigs_batch_multi_device(range(5), dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream)
        engine.check_async()
        off = 0
        for m in range(5):
            q = float(torch.quantile(out[:, off:off + widths[m]].sum(dim=1), 0.95))
            qmax = float(torch.quantile(out[:, off], 0.95))           # rows are descending: first entry = maximum
            off += widths[m]
            assert abs(q / mhm95[m][dim - 1] - 1.0) < 0.007, (m, dim, q, mhm95[m][dim - 1])
            assert abs(qmax / mhm95_max[m][dim - 1] - 1.0) < 0.007, (m, dim, qmax, mhm95_max[m][dim - 1])


def test_full_size_properties(engine):
    """BASELINE metric configuration (dim 12, T 10 000): size-independent properties on 20 000 runs per model."""
    seeds = np.arange(1, 20001, dtype=np.uint32)
    for model in range(5):
        ev = engine.eigs_batch(model, 12, 10000, seeds)
        p = 13 if model in (1, 3) else 12
        assert ev.shape == (seeds.size, p)
        assert np.all(np.isfinite(ev)) and np.all(ev >= 0) and np.all(np.diff(ev, axis=1) <= 0)
        if model in (1, 3):   # rank-d pencil: the extra eigenvalue is numerical noise (Appendix A-5)
            assert np.all(ev[:, -1] < 1e-9 * ev[:, 0])
        assert ev.sum(axis=1).mean() > 100          # trace grows ~2 d^2


# ---- error behaviour (SURVEY.md section 8b "Errors") ----------------------------------------
def test_error_codes(engine):
    import johansen_null_eigenspectra_b200 as jne
    s = np.array([1], dtype=np.uint32)
    for args in [(5, 2, 100), (0, 0, 100), (0, 2, 0), (4, 2, 1)]:
        with pytest.raises(jne.JneError) as e:
            engine.eigs_batch(*args, s)
        assert e.value.status == -1
    with pytest.raises(jne.JneError) as e:
        engine.eigs_batch(0, 16, 100, s)
    assert e.value.status == -4
    with pytest.raises(jne.JneError) as e:      # T < p: singular S2 -> NaN (the reference panics at :45)
        engine.eigs_batch(0, 8, 5, s)
    assert e.value.status == -3
    assert engine.eigs_batch(0, 2, 100, np.array([], dtype=np.uint32)).shape == (0, 2)   # empty input
    assert engine.eigs_batch(0, 2, 100, s).shape == (1, 2)                               # still usable


def test_python_mirror_of_reference_call_sites(engine):
    """calculate_eigenvalues / calculate_eigenvalues_parallel keep the reference's shapes:
    integration/basic_api.rs:35 (2 eigenvalues for model 0 dim 2), multiple_models.rs:32-37 (finite)."""
    import johansen_null_eigenspectra_b200 as jne
    ev = jne.calculate_eigenvalues(2, 103, 1, jne.JohansenModel.NoInterceptNoTrend)
    assert len(ev) == 2 and ev[0] >= ev[1] and all(np.isfinite(ev))
    got = {}
    jne.calculate_eigenvalues_parallel(2, 103, range(1, 6), 0, lambda s, e: got.__setitem__(s, e), engine=engine, batch=2)
    assert sorted(got) == [1, 2, 3, 4, 5] and got[1] == ev
    for m in jne.JohansenModel:
        assert all(np.isfinite(jne.calculate_eigenvalues(2, 103, 7, m)))


# ---- fused multi-model evaluation -----------------------------------
def test_multi_model_is_bit_identical_to_per_model(engine):
    """One pass over the path for several models == the per-model calls, bit for bit."""
    seeds = np.arange(1, 2001, dtype=np.uint32)
    for dim, T in [(1, 40), (3, 103), (5, 1000), (12, 500), (15, 64)]:
        multi = engine.eigs_batch_multi(range(5), dim, T, seeds)
        assert sorted(multi) == [0, 1, 2, 3, 4]
        for m in range(5):
            assert multi[m].shape == (seeds.size, dim + 1 if m in (1, 3) else dim)
            assert np.array_equal(multi[m], engine.eigs_batch(m, dim, T, seeds)), (dim, T, m)
    sub = engine.eigs_batch_multi([1, 4], 4, 200, seeds[:77])          # arbitrary subsets of models
    assert sorted(sub) == [1, 4]
    assert np.array_equal(sub[1], engine.eigs_batch(1, 4, 200, seeds[:77]))
    assert np.array_equal(sub[4], engine.eigs_batch(4, 4, 200, seeds[:77]))
    one = engine.eigs_batch_multi([3], 4, 200, seeds[:5])               # a single model through the multi entry
    assert np.array_equal(one[3], engine.eigs_batch(3, 4, 200, seeds[:5]))


def test_multi_model_full_size_vs_oracle(engine):
    """dim 12, T 10 000: every model's block of the fused pass against the oracle fed the device normals."""
    seeds = np.array([7, 8], dtype=np.uint32)
    multi = engine.eigs_batch_multi(range(5), 12, 10000, seeds)
    for i, s in enumerate(seeds):
        z = engine.gen_normal_matrix(12, 10000, int(s))
        for m in range(5):
            assert_close(multi[m][i], orc.eigs_from_normals(z, m), f"model {m}")


def test_increment_scale_covariance(engine):
    """Scaling the caller's increments by c scales every eigenvalue by c^2 (S1'S1 ~ c^4, S2 ~ c^2).  Tiny and huge
    scales must neither overflow the FP32-seeded reciprocals of the solver nor lose accuracy (the reference's own
    S1'S1 would underflow at c = 1e-140; the whitened formulation here never forms c^4 terms)."""
    rng = np.random.default_rng(12)
    db = rng.standard_normal((3, 300, 5)) / np.sqrt(300)
    for model in (0, 3, 4):
        ref = orc.eigs_batch_from_increments(db, model)
        for scale in (1e-140, 1e-30, 1.0, 1e30, 1e140):
            got = engine.eigs_from_increments(model, db * scale)
            assert_close(got / scale / scale, ref, f"model {model} scale {scale}")


def test_aux_kernels_vs_scalar_sum_kernels():
    """The AUX kernels of the tensor family take the trend moments out of the MMA (table-driven weights in operand
    slots that were padding); JNE_AUX=0 selects the scalar-sum kernels (FP64 sums per lane, summation by parts)
    everywhere.  Same stream, same products otherwise: the records agree to the gate-1 tolerance for every tensor
    layout (JNE_LANE=0: 4, 8 and 12 rows), partial CTAs, ragged T, all models, fused and single-model entries."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path.insert(0, ".")
        import johansen_null_eigenspectra_b200 as jne
        eng = jne.Engine([0])
        out = {}
        for dim, T, n in [(1, 9, 7), (4, 37, 50), (5, 300, 3500), (8, 64, 6000), (9, 1001, 333), (12, 103, 3100), (12, 4000, 64)]:
            seeds = np.arange(5, 5 + n, dtype=np.uint32)
            res = eng.eigs_batch_multi(range(5), dim, T, seeds)
            for m in range(5):
                out[f"multi_{dim}_{T}_{m}"] = res[m]
            for m in (0, 3, 4):
                if m == 4 and T < 3: continue
                out[f"single_{dim}_{T}_{m}"] = eng.eigs_batch(m, dim, T, seeds[: min(n, 500)])
        np.savez(sys.argv[1], **out)
    ''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for aux in ("0", "1"):
        path = f"/tmp/jne_aux_{aux}.npz"
        env = dict(os.environ, JNE_AUX=aux, JNE_LANE="0")   # tensor family for every dim
        r = subprocess.run([sys.executable, "-c", code, path], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[aux] = np.load(path)
    assert set(res["0"].files) == set(res["1"].files)
    for k in res["0"].files:
        a, b = res["1"][k], res["0"][k]
        tol = 1e-9 * np.abs(b) + 1e-12 * b.max(axis=1, keepdims=True)
        assert np.all(np.abs(a - b) <= tol), k


def test_lane_family_vs_tensor_family():
    """dim <= 6 runs on the lane family (csrc/jne_kernels_lane.cuh: one thread per run, FP64 FMA on registers, the sums
    left to right) and dim 9 on its group kernel (3 lanes per run); JNE_LANE=0 sends the same dims through
    the tensor family (one warp per run, DMMA tiles, four time segments, trend moments through the MMA).  Both consume
    the same random stream, so their records agree to rounding -- well inside the gate-1 tolerance -- for every model,
    ragged T (masked tail blocks, empty segments on the tensor side), single-model and fused entry points."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path.insert(0, ".")
        import johansen_null_eigenspectra_b200 as jne
        eng = jne.Engine([0])
        out = {}
        cases = [(1, 9, 70), (2, 1000, 300), (3, 37, 50), (4, 10000, 40), (5, 5000, 200), (6, 103, 3000),
                 (9, 103, 2000), (10, 1001, 300), (9, 10000, 45), (10, 10000, 30), (9, 31, 41), (10, 33, 50)]
        cases += [(d, T, 9) for d in (1, 2, 3, 4, 5, 6, 9, 10) for T in (2 * d + 6, 67, 250)]
        for dim, T, n in cases:
            seeds = np.arange(11, 11 + n, dtype=np.uint32)
            res = eng.eigs_batch_multi(range(5), dim, T, seeds)
            for m in range(5):
                out[f"{dim}_{T}_{m}"] = res[m]
                assert np.array_equal(res[m], eng.eigs_batch(m, dim, T, seeds))
        np.savez(sys.argv[1], **out)
    ''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, lane in (("lane", "1"), ("tensor", "0")):
        path = f"/tmp/jne_lane_{tag}.npz"
        r = subprocess.run([sys.executable, "-c", code, path], cwd=root, env=dict(os.environ, JNE_LANE=lane),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[tag] = np.load(path)
    differ = False
    for k in res["lane"].files:
        a, b = res["lane"][k], res["tensor"][k]
        tol = 1e-10 * np.abs(b) + 1e-13 * b.max(axis=1, keepdims=True)      # a tenth of the gate-1 tolerance
        assert np.all(np.abs(a - b) <= tol), (k, float(np.max(np.abs(a - b) / tol)))
        differ |= not np.array_equal(a, b)
    assert differ, "JNE_LANE=0 did not select a different kernel family"


def test_host_path_chunking_is_invisible(engine):
    """jne_eigs_batch cuts a batch into chunks of whole kernel waves on two streams; the records must not depend on
    it: compare with the one-launch device-pointer entry, over several chunks and a ragged last one."""
    import torch
    n = 20011
    seeds = np.arange(1, n + 1, dtype=np.uint32)
    for model, dim, T in ((2, 3, 40), (4, 12, 24)):
        host = engine.eigs_batch(model, dim, T, seeds)
        ds = torch.from_numpy(seeds.astype(np.int64)).to(torch.int32).cuda().contiguous()
        do = torch.empty((n, host.shape[1]), dtype=torch.float64, device="cuda")
        engine.eigs_batch_device(model, dim, T, ds.data_ptr(), n, do.data_ptr(), torch.cuda.current_stream().cuda_stream)
        engine.check_async()
        assert np.array_equal(do.cpu().numpy(), host)
    multi = engine.eigs_batch_multi(range(5), 6, 16, seeds)
    for m in range(5):
        assert np.array_equal(multi[m], engine.eigs_batch(m, 6, 16, seeds))


def test_device_normals_moments_and_tails(engine):
    """The FP32 / MUFU Box-Muller stream (32-bit radius uniform, |z| <= 6.76) as a N(0,1) sample: mean, variance,
    skewness, kurtosis and the |z| quantiles up to 1 - 1e-5 on 2^25 normals against the exact values, within 4.5
    Monte Carlo standard errors; and -- when the validation build (jne_rng.cuh, -DJNE_RNG_F64: 64-bit uniforms, FP64
    transform of the SAME uniform words) is present -- element by element against it: the transform's error is ~1e-6
    per normal and its variance deficit (-2.53e-7 before the radius-constant calibration, profiles/r2_rng_moments_before_calibration.txt)
    is gone to 3e-8.  The reference's own test is a CDF check to 1e-2 on 60 000 normals
    (src/tests/rng_matrix_test/gen_normal_matrix_test.rs:7-16)."""
    import os, subprocess, sys
    from johansen_null_eigenspectra_b200 import build as jbuild
    dim, steps, calls = 8, 1 << 20, 4
    z = np.concatenate([engine.gen_normal_matrix(dim, steps, 4242 + c).ravel() for c in range(calls)])
    n = z.size
    assert abs(z.mean()) < 4.5 / np.sqrt(n)
    assert abs((z ** 2).mean() - 1.0) < 4.5 * np.sqrt(2.0 / n)
    assert abs((z ** 3).mean()) < 4.5 * np.sqrt(15.0 / n)
    assert abs((z ** 4).mean() - 3.0) < 4.5 * np.sqrt(96.0 / n)
    a = np.sort(np.abs(z))
    for q in (0.9, 0.99, 0.999, 0.9999, 0.99999):
        x = stats.norm.ppf(0.5 + q / 2)
        se = np.sqrt(q * (1 - q) / n) / (2 * stats.norm.pdf(x))
        assert abs(a[int(q * (n - 1))] - x) < 4.5 * se + 2e-6, q
    assert np.abs(z).max() < 6.77
    f64 = jbuild.VARIANTS["rng_f64"][1]
    if not f64.exists():
        return
    code = ("import sys, numpy as np; sys.path.insert(0, '.'); import johansen_null_eigenspectra_b200 as jne; e = jne.Engine([0]); "
            f"np.save(sys.argv[1], np.concatenate([e.gen_normal_matrix({dim}, {steps}, 4242 + c).ravel() for c in range({calls})]))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code, "/tmp/jne_z_f64.npy"], cwd=root, env=dict(os.environ, JNE_LIBRARY=str(f64)),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    y = np.load("/tmp/jne_z_f64.npy")
    assert np.abs(z - y).max() < 2e-3 and np.sqrt(np.mean((z - y) ** 2)) < 1e-6      # same blocks, FP32 vs FP64 transform
    d2 = z * z - y * y
    assert abs(d2.mean()) < 3e-8 + 4.5 * d2.std() / np.sqrt(n), d2.mean()               # variance: calibrated

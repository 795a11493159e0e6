"""CPU tests of the oracle: the reference's own pins, golden fixtures, known answers.
(-m "not gpu")"""
import numpy as np
import pytest
from scipy import stats

from oracle import c_oracle, johansen_oracle as orc, philox_ref
from tests.conftest import eig_tol
from tests.golden import make_golden as mg


# ---- the reference's exact unit pins --------------------------------------------------------
# src/tests/matrix_utils_test/dmatrix_cumsum_test.rs:5-34
@pytest.mark.parametrize("order,expected", [
    ("RowWise", [[1, 3, 6], [4, 9, 15]]),
    ("ColumnWise", [[1, 2, 3], [5, 7, 9]]),
    ("ColumnMajor", [[1, 7, 15], [5, 12, 21]]),
    ("RowMajor", [[1, 3, 6], [10, 15, 21]]),
])
def test_cumsum_orders(order, expected):
    m = np.array([[1.0, 2, 3], [4, 5, 6]])
    assert np.array_equal(orc.dmatrix_cumsum(m, order), np.array(expected, dtype=float))


# src/tests/matrix_utils_test/sum_of_outer_products_test.rs:5-159
@pytest.mark.parametrize("sequential", [False, True])
def test_sum_of_outer_products_pins(sequential):
    f = lambda a, b: orc.sum_of_outer_products(a, b, sequential=sequential)
    a = np.array([[1.0, 3], [2, 4]]); b = np.array([[5.0, 7], [6, 8]])        # basic (:5-27)
    exp = np.outer(a[:, 0], b[:, 0]) + np.outer(a[:, 1], b[:, 1])
    assert np.abs(f(a, b) - exp).max() < 1e-10
    a = np.array([[1.0, 4], [2, 5], [3, 6]]); b = np.array([[7.0, 9], [8, 10]])  # 3x2 result (:29-56)
    r = f(a, b)
    assert r.shape == (3, 2)
    assert np.abs(r - (np.outer(a[:, 0], b[:, 0]) + np.outer(a[:, 1], b[:, 1]))).max() < 1e-10
    a = np.array([[2.0], [3]]); b = np.array([[4.0], [5]])                    # single column (:58-71)
    assert np.abs(f(a, b) - np.outer(a[:, 0], b[:, 0])).max() < 1e-10
    assert np.abs(f(np.zeros((3, 2)), np.zeros((2, 2)))).max() < 1e-10          # zeros (:73-86)
    assert np.abs(f(np.eye(2), np.eye(2)) - np.eye(2)).max() < 1e-10            # identity-like (:88-110)
    i, j = np.meshgrid(np.arange(5), np.arange(10), indexing="ij")              # "large" (:112-136)
    a = (i + j).astype(float); b = (i * j + 1).astype(float)
    r = f(a, b)
    assert r.shape == (5, 5) and r[0, 0] >= np.outer(a[:, 0], b[:, 0])[0, 0]
    a = np.array([[1e-10, 3e-10], [2e-10, 4e-10]]); b = np.array([[5e-10, 7e-10], [6e-10, 8e-10]])  # (:138-159)
    assert np.abs(f(a, b) - (np.outer(a[:, 0], b[:, 0]) + np.outer(a[:, 1], b[:, 1]))).max() < 1e-25


# ---- indirect pins of the eigenvalue path -----------------------------------------------------
@pytest.mark.parametrize("model", range(5))
@pytest.mark.parametrize("dim", [1, 2, 5])
def test_count_order_finite(model, dim):
    # count: src/data_storage/thread_manager.rs:40-44, integration/basic_api.rs:35
    # finite: integration/multiple_models.rs:32-37; descending: src/johansen_statistics.rs:45
    z = np.random.default_rng(5).standard_normal((dim, 103))
    ev = orc.eigs_from_normals(z, model)
    assert ev.shape == (dim + 1 if model in (1, 3) else dim,)
    assert np.all(np.isfinite(ev)) and np.all(np.diff(ev) <= 0) and np.all(ev >= 0)


def test_f_matrix_semantics():
    """SURVEY.md Appendix A items 1-7."""
    rng = np.random.default_rng(3)
    b = rng.standard_normal((4, 50)).cumsum(axis=1)
    T = 50
    tau = np.arange(1, T + 1) / T
    assert np.array_equal(orc.construct_f_matrix(b, 0), b)
    f1 = orc.construct_f_matrix(b, 1)
    assert f1.shape == (5, T) and np.all(f1[4] == 1.0) and np.array_equal(f1[:4], b)
    f2 = orc.construct_f_matrix(b, 2)
    assert f2.shape == (4, T)
    assert np.allclose(f2[:3], b[:3] - b[:3].mean(axis=1, keepdims=True))
    assert np.allclose(f2[3], tau - 0.5) and abs(f2[3].mean() - 1 / (2 * T)) < 1e-15  # trend NOT demeaned
    f3 = orc.construct_f_matrix(b, 3)
    assert f3.shape == (5, T) and np.allclose(f3[:4].sum(axis=1), 0, atol=1e-12) and np.allclose(f3[4], tau - 0.5)
    f4 = orc.construct_f_matrix(b, 4)
    assert f4.shape == (4, T)
    zz = np.stack([np.ones(T), tau])
    assert np.abs(f4 @ zz.T).max() < 1e-9          # residuals orthogonal to [1, tau]
    # dim 1: models 2 and 4 have a deterministic F
    b1 = b[:1]
    assert np.allclose(orc.construct_f_matrix(b1, 2), (tau - 0.5)[None])
    assert orc.construct_f_matrix(b1, 4).shape == (1, T)


def test_increments_and_normals_entry_agree():
    z = np.random.default_rng(11).standard_normal((3, 200))
    for model in range(5):
        a = orc.eigs_from_normals(z, model)
        b = orc.eigs_from_increments(z * np.sqrt(1 / 200), model)
        assert np.all(np.abs(a - b) <= eig_tol(a))


def test_dggev_equals_symmetric_definite_solver():
    """The pencil is symmetric-definite: dggev must agree with Cholesky-whitening + eigh
    (SURVEY.md section 8c (iii)); this is what licenses the GPU's Cholesky + Jacobi solve."""
    import scipy.linalg as sl
    rng = np.random.default_rng(2)
    for model in range(5):
        z = rng.standard_normal((6, 400))
        dt = 1 / 400
        bm = orc.brownian_motion_from_normals(z, dt)
        prev, dbm = bm[:, :400], bm[:, 1:] - bm[:, :400]
        fm = orc.construct_f_matrix(prev, model)
        s1 = dbm @ fm.T
        s2 = fm @ fm.T * dt
        ref = orc.calculate_eigenvalues_from_matrices(prev, dbm, dt, model)
        w = sl.eigh(s1.T @ s1, s2, eigvals_only=True)[::-1]
        k = 6  # genuine eigenvalues
        assert np.all(np.abs(w[:k] - ref[:k]) <= 1e-10 * np.abs(ref[:k]))
        if model in (1, 3):
            assert abs(ref[-1]) < 1e-10 * ref[0]   # spurious ~0 eigenvalue (Appendix A-5)


# ---- golden fixtures ----------------------------------------------------------------------
def test_golden_fixtures_match_oracle():
    g = np.load(mg.HERE / "eigs_from_increments.npz")
    assert len(g.files) >= 50
    for key in g.files:
        m, d, t = (int(s[1:]) for s in key.split("_"))
        db = mg.increments(mg.case_seed(m, d, t), mg.N, t, d)
        got = orc.eigs_batch_from_increments(db, m)
        assert np.all(np.abs(got - g[key]) <= eig_tol(g[key])), key


def test_c_oracle_matches_numpy_oracle():
    lib = c_oracle.load()
    g = np.load(mg.HERE / "eigs_from_increments.npz")
    for key in g.files:
        m, d, t = (int(s[1:]) for s in key.split("_"))
        db = mg.increments(mg.case_seed(m, d, t), mg.N, t, d)
        for i in range(mg.N):
            got = c_oracle.eigs_from_increments(lib, db[i], m)
            assert np.all(np.abs(got - g[key][i]) <= eig_tol(g[key][i])), key
    c1 = np.load(mg.HERE / "c1_model0_dim2_steps1000.npz")["eigs"]
    db = mg.increments(mg.SEED, 64, 1000, 2)
    for i in range(64):
        assert np.all(np.abs(c_oracle.eigs_from_increments(lib, db[i], 0) - c1[i]) <= eig_tol(c1[i]))


def test_optimised_cpu_variant_matches_faithful_port():
    """The "optimised CPU" baseline (one pass, raw moments + Schur complements, dsygv) is a different algorithm for
    the same quantity: same increments => same eigenvalues as the faithful port, to gate-(1) tolerance."""
    lib = c_oracle.load()
    rng = np.random.default_rng(21)
    for model in range(5):
        for d, T in [(1, 30), (2, 103), (5, 400), (12, 1000), (15, 64)]:
            db = rng.standard_normal((T, d)) / np.sqrt(T)
            ref = orc.eigs_from_increments(db.T, model)
            got = c_oracle.fast_from_increments(lib, db, model)
            assert np.all(np.abs(got - ref) <= eig_tol(ref)), (model, d, T)
    a = c_oracle.fast_batch(lib, 3, 4, 200, np.arange(1, 9), threads=2)
    assert a.shape == (8, 5) and np.array_equal(a, c_oracle.fast_batch(lib, 3, 4, 200, np.arange(1, 9), threads=1))


def test_c_oracle_rng_is_standard_normal_and_reproducible():
    """The reference's own RNG tests, applied to the port of its generator:
    src/tests/rng_matrix_test/gen_normal_matrix_test.rs:7-16 (CDF within 1e-2 at 99 quantiles),
    brownian_motion_test.rs:99-153 (same seed same matrix, different seeds differ)."""
    lib = c_oracle.load()
    z = c_oracle.gen_normal_matrix(lib, 200, 300, 42, ncpu=8)
    qs = np.arange(1, 100) / 100
    emp = np.searchsorted(np.sort(z.ravel()), stats.norm.ppf(qs)) / z.size
    assert np.abs(emp - qs).max() < 1e-2
    assert np.array_equal(z, c_oracle.gen_normal_matrix(lib, 200, 300, 42, ncpu=8))
    assert not np.array_equal(z, c_oracle.gen_normal_matrix(lib, 200, 300, 43, ncpu=8))
    # the stream depends on the physical core count (src/rng_matrix.rs:16-20), SURVEY section 0 item 5
    assert not np.array_equal(z, c_oracle.gen_normal_matrix(lib, 200, 300, 42, ncpu=16))
    ev = c_oracle.eigs_batch(lib, 0, 2, 103, np.arange(1, 6), threads=2, ncpu=8)
    ev2 = c_oracle.eigs_batch(lib, 0, 2, 103, np.arange(1, 6), threads=1, ncpu=8)
    assert ev.shape == (5, 2) and np.array_equal(ev, ev2)      # resumable.rs:56-71 (same seed => same record)


# ---- external known answers -----------------------------------------------------------------
@pytest.mark.parametrize("model", [2, 4])
def test_chi2_law_dim1(model):
    """Models 2 and 4 at dim 1 have a deterministic F, so lambda ~ chi^2(1) exactly for every T."""
    rng = np.random.default_rng(100 + model)
    ev = np.array([orc.eigs_from_normals(rng.standard_normal((1, 50)), model)[0] for _ in range(3000)])
    assert stats.kstest(ev, stats.chi2(1).cdf).pvalue > 1e-3


def test_mhm_critical_values_model0_dim2():
    """95 % trace / max-eig quantiles vs MacKinnon-Haug-Michelis (12.32 / 11.22);
    T = 400, 4000 runs -> Monte Carlo error ~3 %, finite-T bias ~1 %."""
    rng = np.random.default_rng(7)
    ev = np.array([orc.eigs_from_normals(rng.standard_normal((2, 400)), 0) for _ in range(4000)])
    tr95 = orc.percentiles(ev.sum(axis=1), (0.95,))[0]
    mx95 = orc.percentiles(ev[:, 0], (0.95,))[0]
    assert abs(tr95 - 12.32) / 12.32 < 0.06
    assert abs(mx95 - 11.22) / 11.22 < 0.06


# ---- device-stream restatement -----------------------------------------------------------------
def test_philox_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in kat:   # Random123 kat_vectors, philox4x32 10 rounds
        assert tuple(int(x) for x in philox_ref.philox4x32_10(*ctr, *key)) == exp


def test_xoshiro128pp_known_answers():
    """The substream generator of the device stream against the published reference vector of xoshiro128++
    (state {1, 2, 3, 4}; the vector rand_xoshiro 0.7.0 -- the reference's own generator crate -- tests its
    Xoshiro128PlusPlus with, produced by Vigna's C implementation)."""
    st = [np.array([v], dtype=np.uint32) for v in (1, 2, 3, 4)]
    got = [int(philox_ref.xoshiro128pp(st)[0]) for _ in range(10)]
    assert got == [641, 1573767, 3222811527, 3517856514, 836907274, 4247214768, 3867114732, 1355841295, 495546011,
                   621204420]


def test_stream_words_layout():
    """Block b of row r is block (b & 31) >> 1 of substream (r, b >> 5, b & 1): check the vectorised restatement
    against a scalar walk of one substream, across an epoch boundary."""
    seed, row = 12345, 3
    w = philox_ref.stream_words(5, 600, seed)          # 150 blocks: epochs 0..4
    for b in (0, 1, 2, 31, 32, 33, 95, 149):
        e, h, j = b >> 5, b & 1, (b & 31) >> 1
        st = [np.array([int(x)], dtype=np.uint32) for x in philox_ref.philox4x32_10(e, row, h, 0, seed, philox_ref.KEY1)]
        for _ in range(4 * j):
            philox_ref.xoshiro128pp(st)
        assert [int(philox_ref.xoshiro128pp(st)[0]) for _ in range(4)] == [int(x) for x in w[row, b]]


def test_stream_has_no_structure_across_blocks_halves_and_epochs():
    """The stream interleaves two generators per row (the halves) and re-keys them every 128 steps.  Serial correlations
    of the normals and of their squares at the lags that structure could show at (neighbouring steps, the four-step
    block, the other half, the epoch and its neighbours), correlations between rows, and the byte frequencies of the
    words: all at the level of independent N(0, 1) / uniform draws.  (The words are bit-exact with the device, so this
    CPU test speaks for the device stream.)"""
    zs = [philox_ref.normal_matrix(8, 1 << 15, seed) for seed in (1, 2, 3)]
    n = zs[0].shape[1]
    for f in (lambda v: v, lambda v: v * v):
        for lag in (1, 2, 4, 8, 127, 128, 129, 256):
            r = np.array([np.corrcoef(f(z[i, :-lag]), f(z[i, lag:]))[0, 1] for z in zs for i in range(8)]) * np.sqrt(n)
            assert abs(r.mean()) < 4.0 / np.sqrt(r.size), (lag, r.mean())      # no systematic correlation (4 sigma of the mean)
            assert np.abs(r).max() < 4.5, (lag, np.abs(r).max())
    for z in zs:
        c = np.corrcoef(z)
        np.fill_diagonal(c, 0.0)
        assert np.abs(c).max() * np.sqrt(n) < 4.5
    w = philox_ref.stream_words(8, 1 << 15, 7).ravel()
    for shift in (0, 8, 16, 24):
        cnt = np.bincount((w >> np.uint32(shift)) & np.uint32(0xFF), minlength=256)
        chi2 = ((cnt - w.size / 256) ** 2 / (w.size / 256)).sum()
        assert stats.chi2.sf(chi2, 255) > 1e-4, (shift, chi2)


def test_philox_normal_matrix_is_standard_normal():
    z = philox_ref.normal_matrix(200, 300, 42)
    qs = np.arange(1, 100) / 100
    emp = np.searchsorted(np.sort(z.ravel()), stats.norm.ppf(qs)) / z.size
    assert np.abs(emp - qs).max() < 1e-2          # the reference's own criterion
    assert np.array_equal(z[:5, :40], philox_ref.normal_matrix(5, 40, 42))   # prefix property


def test_port_generators_match_published_vectors():
    """The CPU port's restatement of the reference's third-party generators (rand_xoshiro 0.7.0, not under
    /root/reference) against their published known-answer vectors: SplitMix64 from seed 1234567 (Vigna's
    splitmix64.c) and xoshiro256++ from the state {1, 2, 3, 4} (the `reference` test in rand_xoshiro's
    xoshiro256plusplus.rs); seed_from_u64 fills the state with four SplitMix64 outputs."""
    import ctypes as C
    lib = c_oracle.load()
    out = np.zeros(10, dtype=np.uint64)
    lib.jne_oracle_splitmix64(C.c_uint64(1234567), C.c_size_t(5), C.c_void_p(out.ctypes.data))
    assert out[:5].tolist() == [6457827717110365317, 3203168211198807973, 9817491932198370423,
                                4593380528125082431, 16408922859458223821]
    state = np.array([1, 2, 3, 4], dtype=np.uint64)
    lib.jne_oracle_xoshiro_from_state(C.c_void_p(state.ctypes.data), C.c_size_t(10), C.c_void_p(out.ctypes.data))
    assert out.tolist() == [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
                            9973669472204895162, 14011001112246962877, 12406186145184390807,
                            15849039046786891736, 10450023813501588000]
    st = np.zeros(4, dtype=np.uint64)
    lib.jne_oracle_xoshiro_seed_from_u64(C.c_uint64(1234567), C.c_void_p(st.ctypes.data))
    assert st.tolist() == [6457827717110365317, 3203168211198807973, 9817491932198370423, 4593380528125082431]

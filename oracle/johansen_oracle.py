"""CPU oracle for the Johansen null-eigenspectra hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy restatement of the reference's per-run algorithm
(Kuan-Lun/johansen-null-eigenspectra v0.8.0, citations relative to /root/reference).
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
path (``johansen_null_eigenspectra_b200``) never imports anything from ``oracle/``.

PARITY STATUS: "parity unpinned".  The reference cannot be built here (no cargo/rustc,
git-pinned nalgebra not vendored, no system LAPACK) and it holds NO golden vector or
known-answer test for ``construct_f_matrix`` / ``calculate_eigenvalues*``.
What IS pinned, and checked in tests/test_oracle.py:
  * the reference's exact unit pins for cumsum (src/tests/matrix_utils_test/dmatrix_cumsum_test.rs:5-34)
    and sum_of_outer_products (src/tests/matrix_utils_test/sum_of_outer_products_test.rs:5-159);
  * indirect pins: eigenvalue count per model (src/data_storage/thread_manager.rs:40-44,
    src/tests/data_storage/integration/basic_api.rs:35), finiteness for all five models
    (integration/multiple_models.rs:32-37), descending order (src/johansen_statistics.rs:45);
  * external known answers: chi^2(1) law for models 2 and 4 at dim 1, and the published
    Johansen / MacKinnon-Haug-Michelis 95 % critical values.

Third-party arithmetic not under /root/reference (named, with pinned versions from Cargo.lock):
  * nalgebra 0.33.2 / nalgebra-lapack 0.25.0 @ git fa7afd5e ``GeneralizedEigen::new`` ->
    lapack 0.19.0 / lapack-sys 0.14.0 -> system LAPACK ``dggev('V','V')``
    (call site src/johansen_statistics.rs:35-44).  Restated with scipy.linalg.lapack.dggev:
    the same LAPACK routine, from scipy's bundled OpenBLAS.
  * rand 0.9.1 / rand_distr 0.5.1 / rand_xoshiro 0.7.0 (src/rng_matrix.rs:23-32): the
    Xoshiro256++ + ziggurat stream is NOT reproduced here (machine-dependent by
    construction, src/rng_matrix.rs:16-20).  This oracle takes normals or increments as
    INPUT; oracle/jne_oracle.c restates the published RNG algorithms for the timed CPU baseline.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack as _lapack

__all__ = [
    "num_eigs",
    "dmatrix_cumsum",
    "sum_of_outer_products",
    "brownian_motion_from_normals",
    "construct_f_matrix",
    "calculate_eigenvalues_from_matrices",
    "eigs_from_normals",
    "eigs_from_increments",
    "eigs_batch_from_normals",
    "eigs_batch_from_increments",
    "percentiles",
]


def num_eigs(model: int, dim: int) -> int:
    """Eigenvalues per run: dim+1 for models 1 and 3, else dim
    (src/data_storage/thread_manager.rs:40-44)."""
    if model not in (0, 1, 2, 3, 4):
        raise ValueError(f"model must be 0..4, got {model}")
    return dim + 1 if model in (1, 3) else dim


def dmatrix_cumsum(matrix: np.ndarray, order: str) -> np.ndarray:
    """The four cumulative-sum orders of src/matrix_utils.rs:13-65.  Only ``RowWise``
    is on the hot path (src/rng_matrix.rs:51,140).  np.cumsum is a naive left-to-right
    float64 accumulation, like the reference's ``acc += v`` loops."""
    m = np.asarray(matrix, dtype=np.float64)
    nrows, ncols = m.shape
    if order == "RowWise":  # each row accumulated along the columns (:51-63)
        return np.cumsum(m, axis=1)
    if order == "ColumnWise":  # each column accumulated along the rows (:39-50)
        return np.cumsum(m, axis=0)
    if order == "ColumnMajor":  # whole matrix, column-major element order (:15-24)
        return np.cumsum(m.flatten(order="F")).reshape((nrows, ncols), order="F")
    if order == "RowMajor":  # whole matrix, row-major element order (:25-38)
        return np.cumsum(m.flatten(order="C")).reshape((nrows, ncols), order="C")
    raise ValueError(order)


def sum_of_outer_products(a: np.ndarray, b: np.ndarray, sequential: bool = False) -> np.ndarray:
    """sum_t a[:,t] b[:,t]'  ->  a_rows x b_rows  (src/matrix_utils.rs:67-85).

    The reference reduces with rayon in a scheduling-dependent order, so no summation
    order is canonical.  ``sequential=True`` adds the T outer products left to right;
    the default uses one matrix product (differences are O(eps) per element)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape[1] == b.shape[1]
    if sequential:
        acc = np.zeros((a.shape[0], b.shape[0]))
        for t in range(a.shape[1]):
            acc += np.outer(a[:, t], b[:, t])
        return acc
    return a @ b.T


def brownian_motion_from_normals(z: np.ndarray, delta_t: float) -> np.ndarray:
    """``brownian_motion_matrix`` with ``AlongColumns`` and a zero start column
    (src/rng_matrix.rs:57-141) given the d x T normal matrix that ``gen_normal_matrix``
    would have produced.  z[r, c] is the flat element c*d + r of the reference's
    column-major buffer (src/rng_matrix.rs:36).  Returns d x (T+1)."""
    z = np.asarray(z, dtype=np.float64)
    d, _ = z.shape
    zfull = np.concatenate([np.zeros((d, 1)), z], axis=1)  # make_z_matrix :95-111
    scaled = zfull * np.sqrt(delta_t)                       # :138-139
    return dmatrix_cumsum(scaled, "RowWise")                # :140


def construct_f_matrix(bm_previous: np.ndarray, model: int) -> np.ndarray:
    """F_{t-1} per model (src/johansen_statistics.rs:102-197)."""
    bm_previous = np.asarray(bm_previous, dtype=np.float64)
    rows, cols = bm_previous.shape
    t = float(cols)
    if model == 0:  # :106
        return bm_previous.copy()
    if model == 1:  # :108-113
        return np.concatenate([bm_previous, np.ones((1, cols))], axis=0)
    if model == 2:  # :115-138  first rows-1 Brownian rows demeaned + trend (i+1)/T - 0.5
        x = bm_previous[: rows - 1, :].copy()
        if x.shape[0]:
            x -= x.mean(axis=1, keepdims=True)
        y = (np.arange(1, cols + 1, dtype=np.float64) / t - 0.5)[None, :]
        return np.concatenate([x, y], axis=0)
    if model == 3:  # :140-162  all rows demeaned + trend row (NOT demeaned)
        x = bm_previous - bm_previous.mean(axis=1, keepdims=True)
        y = (np.arange(1, cols + 1, dtype=np.float64) / t - 0.5)[None, :]
        return np.concatenate([x, y], axis=0)
    if model == 4:  # :164-195  residual of [B_{1..d-1}; tau^2] on [1; tau]
        y = (np.arange(1, cols + 1, dtype=np.float64) / t)[None, :]
        x_with_y2 = np.concatenate([bm_previous[: rows - 1, :], y ** 2], axis=0)
        zm = np.concatenate([np.ones((1, cols)), y], axis=0)
        zzt = zm @ zm.T
        # explicit 2x2 inverse, as nalgebra's try_inverse does for 2x2 (:191-192)
        det = zzt[0, 0] * zzt[1, 1] - zzt[0, 1] * zzt[1, 0]
        if det == 0.0:
            raise np.linalg.LinAlgError("singular Z Z' (steps < 2); the reference panics here (:192)")
        zzt_inv = np.array([[zzt[1, 1], -zzt[0, 1]], [-zzt[1, 0], zzt[0, 0]]]) / det
        projection = x_with_y2 @ zm.T @ zzt_inv @ zm
        return x_with_y2 - projection
    raise ValueError(f"model must be 0..4, got {model}")


def calculate_eigenvalues_from_matrices(bm_previous, dbm, delta_t: float, model: int) -> np.ndarray:
    """src/johansen_statistics.rs:24-47: S1 = sum dB F' (d x p), S2 = dt * sum F F' (p x p),
    LAPACK dggev('V','V') on the pencil (S1'S1, S2), |alpha|/beta, sorted descending."""
    fm = construct_f_matrix(bm_previous, model)
    s1 = sum_of_outer_products(dbm, fm)                 # :32
    s2 = sum_of_outer_products(fm, fm) * delta_t        # :33
    a = np.asfortranarray(s1.T @ s1)
    b = np.asfortranarray(s2)
    alphar, alphai, beta, _vl, _vr, _work, info = _lapack.dggev(a, b, compute_vl=1, compute_vr=1)
    if info != 0:
        raise np.linalg.LinAlgError(f"dggev info={info}")
    with np.errstate(divide="ignore", invalid="ignore"):
        ev = np.hypot(alphar, alphai) / beta            # :40-44  val.0.norm() / val.1
    if np.isnan(ev).any():
        raise FloatingPointError("NaN eigenvalue; the reference panics in partial_cmp().unwrap() (:45)")
    return np.sort(ev)[::-1].copy()                     # :45 descending


def eigs_from_normals(z: np.ndarray, model: int) -> np.ndarray:
    """``calculate_eigenvalues`` (src/johansen_statistics.rs:59-85) with the d x T matrix
    of standard normals supplied instead of generated."""
    z = np.asarray(z, dtype=np.float64)
    _, steps = z.shape
    delta_t = 1.0 / float(steps)                        # :70
    bm = brownian_motion_from_normals(z, delta_t)       # :71-78
    bm_current = bm[:, 1 : steps + 1]                   # :80
    bm_previous = bm[:, 0:steps]                        # :81
    dbm = bm_current - bm_previous                      # :82 (re-derived by subtraction)
    return calculate_eigenvalues_from_matrices(bm_previous, dbm, delta_t, model)


def eigs_from_increments(db: np.ndarray, model: int) -> np.ndarray:
    """Same path, but starting from the already scaled increments dB = sqrt(dt) z (d x T):
    B_0 = 0, B_t = B_{t-1} + dB_t by the naive cumsum (src/matrix_utils.rs:51-63), then
    dB re-derived by subtraction as the reference does (src/johansen_statistics.rs:82).
    This is the oracle side of parity gate (1) (shared increments)."""
    db = np.asarray(db, dtype=np.float64)
    d, steps = db.shape
    delta_t = 1.0 / float(steps)
    bm = dmatrix_cumsum(np.concatenate([np.zeros((d, 1)), db], axis=1), "RowWise")
    bm_previous = bm[:, 0:steps]
    dbm = bm[:, 1 : steps + 1] - bm_previous
    return calculate_eigenvalues_from_matrices(bm_previous, dbm, delta_t, model)


def eigs_batch_from_normals(z: np.ndarray, model: int) -> np.ndarray:
    """z: (n, T, d) C-order == per-run column-major d x T (src/rng_matrix.rs:36)."""
    z = np.asarray(z, dtype=np.float64)
    n, _, d = z.shape
    out = np.empty((n, num_eigs(model, d)))
    for i in range(n):
        out[i] = eigs_from_normals(z[i].T, model)
    return out


def eigs_batch_from_increments(db: np.ndarray, model: int) -> np.ndarray:
    """db: (n, T, d) C-order == per-run column-major d x T."""
    db = np.asarray(db, dtype=np.float64)
    n, _, d = db.shape
    out = np.empty((n, num_eigs(model, d)))
    for i in range(n):
        out[i] = eigs_from_increments(db[i].T, model)
    return out


def percentiles(values, ps=(0.5, 0.75, 0.8, 0.85, 0.9, 0.95, 0.975, 0.99)) -> np.ndarray:
    """Linear-interpolated percentiles at rank p*(n-1), the parity observable of
    src/simulation_analyzers.rs:4-18 (trace = sum of eigenvalues, max-eig = largest)."""
    v = np.sort(np.asarray(values, dtype=np.float64))
    n = v.size
    out = []
    for p in ps:
        idx = p * (n - 1)
        lo = int(np.floor(idx))
        hi = int(np.ceil(idx))
        w = idx - lo
        out.append(v[lo] if lo == hi else v[lo] * (1.0 - w) + v[hi] * w)   # simulation_analyzers.rs:12-17
    return np.array(out)

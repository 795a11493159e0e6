"""ctypes binding of include/jne.h plus the Python mirror of the reference's call sites."""
from __future__ import annotations

import ctypes as C
import enum
import os
from pathlib import Path
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB_PATH = Path(os.environ.get("JNE_LIBRARY", _PKG / "libjne.so"))


class JneError(RuntimeError):
    """Raised for any negative jne_status; carries the library's message."""

    def __init__(self, status: int, message: str):
        super().__init__(f"jne status {status}: {message}")
        self.status = status


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m johansen_null_eigenspectra_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback for the hot path."
        )
    lib = C.CDLL(str(_LIB_PATH))
    u8, u32, u64, i64, dbl = C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64, C.c_double
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    sig = {
        "jne_version": (C.c_char_p, []),
        "jne_device_count": (C.c_int, []),
        "jne_init": (C.c_int, [ip, C.c_int, C.POINTER(vp)]),
        "jne_shutdown": (None, [vp]),
        "jne_last_error": (C.c_char_p, [vp]),
        "jne_num_eigs": (C.c_int, [u8, u32]),
        "jne_eigs_batch": (C.c_int, [vp, u8, u32, u32, vp, u64, vp]),
        "jne_eigs_batch_multi": (C.c_int, [vp, u32, u32, u32, vp, u64, vp]),
        "jne_eigs_batch_multi_device": (C.c_int, [vp, u32, u32, u32, vp, u64, vp, vp]),
        "jne_eigs_batch_multi_stream": (C.c_int, [vp, u32, u32, u32, vp, u64, vp, vp]),
        "jne_multi_width": (C.c_int, [u32, u32]),
        "jne_submit": (i64, [vp, u8, u32, u32, vp, u64, vp]),
        "jne_wait": (C.c_int, [vp, i64]),
        "jne_eigs_batch_device": (C.c_int, [vp, u8, u32, u32, vp, u64, vp, vp]),
        "jne_check_async": (C.c_int, [vp]),
        "jne_eigs_from_increments": (C.c_int, [vp, u8, u32, u32, vp, u64, vp]),
        "jne_gen_normal_matrix": (C.c_int, [vp, u32, u32, u32, vp]),
        "jne_brownian_motion_matrix": (C.c_int, [vp, u32, u32, dbl, u32, vp]),
        "jne_pencil_eigs_batch": (C.c_int, [vp, u32, u32, vp, vp, u64, vp]),
        "jne_eigs_batch_debug": (C.c_int, [vp, u8, u32, u32, vp, u64, vp, vp]),
        "jne_percentiles_device": (C.c_int, [vp, vp, u64, u32, u32, vp, u32, vp, vp, vp]),
        "jne_simulate_percentiles": (C.c_int, [vp, u8, u32, u32, u32, u64, vp, u32, vp, vp]),
        "jne_simulate_percentiles_multi": (C.c_int, [vp, u32, u32, u32, u32, u64, vp, u32, vp, vp]),
        "jne_run_model_simulation": (C.c_int, [vp, u8, u32, u32, u64, C.c_char_p, C.c_int, ip, C.c_int, C.POINTER(u64)]),
        "jne_fp64_peak_tflops": (C.c_int, [vp, C.c_int, dbl, C.POINTER(dbl)]),
        "jne_launch_count": (u64, [vp]),
        "jne_flops_per_run": (dbl, [u8, u32, u32]),
        "jne_jacobi_table": (C.c_int, [u32, vp, u32]),
        "jne_trend_weight_table": (C.c_int64, [u32, vp, u64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def version() -> str:
    return lib.jne_version().decode()


def num_eigs(model: int, dim: int) -> int:
    r = lib.jne_num_eigs(int(model), int(dim))
    if r < 0:
        raise JneError(r, "invalid model/dim")
    return r


def flops_per_run(model: int, dim: int, steps: int) -> float:
    return float(lib.jne_flops_per_run(int(model), int(dim), int(steps)))


def jacobi_table(ne: int) -> np.ndarray:
    """The device Jacobi's step table for an ne x ne problem as (ne - 1, ne/2 + blocks) uint32 words (host-only)."""
    n = lib.jne_jacobi_table(int(ne), None, 0)
    if n < 0:
        raise JneError(n, "jne_jacobi_table: ne must be even, 2..16")
    words = np.empty(n, dtype=np.uint32)
    lib.jne_jacobi_table(int(ne), words.ctypes.data, n)
    return words.reshape(ne - 1, -1)


def trend_weight_table(steps: int) -> np.ndarray:
    """The AUX kernels' trend-weight table for a run of `steps` steps as (seg_len, 4 weights, 4 segments) doubles
    (host-only; see include/jne.h)."""
    n = lib.jne_trend_weight_table(int(steps), None, 0)
    if n < 0:
        raise JneError(int(n), "jne_trend_weight_table: steps must be 1..2^22")
    tab = np.empty(n, dtype=np.float64)
    lib.jne_trend_weight_table(int(steps), tab.ctypes.data, n)
    return tab.reshape(-1, 4, 4)


class JohansenModel(enum.IntEnum):
    """src/johansen_models.rs:6-51 -- number <-> variant is part of the .dat contract."""

    NoInterceptNoTrend = 0
    InterceptNoTrendWithInterceptInCoint = 1
    InterceptNoTrendUnrestrictedIntercept = 2
    InterceptTrendUnrestrictedInterceptRestrictedTrend = 3
    InterceptTrendUnrestrictedBoth = 4

    def to_number(self) -> int:
        return int(self)

    @classmethod
    def from_number(cls, n: int) -> Optional["JohansenModel"]:
        try:
            return cls(n)
        except ValueError:
            return None

    @classmethod
    def all_models(cls):
        return list(cls)

    @classmethod
    def default(cls) -> "JohansenModel":  # src/johansen_models.rs: Default = model 2
        return cls.InterceptNoTrendUnrestrictedIntercept

    def has_intercept(self) -> bool:
        return self != JohansenModel.NoInterceptNoTrend

    def has_trend(self) -> bool:
        return self in (JohansenModel.InterceptTrendUnrestrictedInterceptRestrictedTrend,
                        JohansenModel.InterceptTrendUnrestrictedBoth)

    def num_eigs(self, dim: int) -> int:
        return num_eigs(int(self), dim)


def _model_number(model) -> int:
    return int(model)


class Engine:
    """Owns a ``jne_ctx``.  ``devices=None`` uses every visible GPU; a list pins device ids."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self._ctx = C.c_void_p()
        if devices is None:
            rc = lib.jne_init(None, 0, C.byref(self._ctx))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = lib.jne_init(arr, len(devices), C.byref(self._ctx))
        if rc != 0:
            msg = lib.jne_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise JneError(rc, msg)

    # -- plumbing -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            lib.jne_shutdown(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc < 0:
            raise JneError(rc, lib.jne_last_error(self._ctx).decode())

    @property
    def launch_count(self) -> int:
        return int(lib.jne_launch_count(self._ctx))

    # -- the hot path ---------------------------------------------------------------------
    def eigs_batch(self, model, dim: int, steps: int, seeds) -> np.ndarray:
        """(n, p) float64, row i = descending eigenvalues of seeds[i]."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        p = num_eigs(_model_number(model), dim)
        out = np.empty((seeds.size, p), dtype=np.float64)
        self._check(lib.jne_eigs_batch(self._ctx, _model_number(model), dim, steps,
                                       seeds.ctypes.data, seeds.size, out.ctypes.data))
        return out

    @staticmethod
    def _mask(models) -> int:
        mask = 0
        for m in models:
            mask |= 1 << _model_number(m)
        return mask

    def eigs_batch_multi(self, models, dim: int, steps: int, seeds, out: Optional[np.ndarray] = None) -> dict:
        """One pass over each seed's Brownian path for all `models`; {model: (n, p) float64} (views of one (n, width)
        array; pass `out` to reuse a caller-owned buffer)."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        models = sorted({_model_number(m) for m in models})
        mask = self._mask(models)
        width = lib.jne_multi_width(mask, dim)
        if width < 0:
            raise JneError(width, "invalid model mask / dim")
        if out is None:
            out = np.empty((seeds.size, width), dtype=np.float64)
        assert out.shape == (seeds.size, width) and out.dtype == np.float64 and out.flags.c_contiguous
        self._check(lib.jne_eigs_batch_multi(self._ctx, mask, dim, steps, seeds.ctypes.data, seeds.size,
                                             out.ctypes.data))
        res, off = {}, 0
        for m in models:
            p = num_eigs(m, dim)
            res[m] = out[:, off:off + p]
            off += p
        return res

    def eigs_batch_multi_stream(self, models, dim: int, steps: int, seeds, sink) -> None:
        """The fused batch with its rows SENT to ``sink(first, rows)`` as they arrive (the reference's channel model,
        src/data_storage/parallel_compute.rs:14-41): ``rows`` is a (count, width) float64 view of the library's pinned
        staging memory, valid during the call only (copy what you keep); ``first`` is the index of its first seed in
        ``seeds``.  The sink runs on the library's per-device host threads (under the GIL here), for disjoint ranges,
        in no particular order; an exception or a truthy return value aborts the batch."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        models = sorted({_model_number(m) for m in models})
        mask = self._mask(models)
        width = lib.jne_multi_width(mask, dim)
        if width < 0:
            raise JneError(width, "invalid model mask / dim")
        failure = []

        @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_double))
        def _sink(_user, first, count, rows):
            try:
                view = np.ctypeslib.as_array(rows, shape=(count, width))
                return 1 if sink(int(first), view) else 0
            except BaseException as exc:      # nothing may propagate into the C frames
                failure.append(exc)
                return 1

        rc = lib.jne_eigs_batch_multi_stream(self._ctx, mask, dim, steps, seeds.ctypes.data, seeds.size,
                                             C.cast(_sink, C.c_void_p), None)
        if failure:
            raise failure[0]
        self._check(rc)

    def eigs_batch_multi_device(self, models, dim: int, steps: int, d_seeds_ptr: int, n: int, d_out_ptr: int,
                                stream_ptr: int = 0) -> None:
        self._check(lib.jne_eigs_batch_multi_device(self._ctx, self._mask(models), dim, steps,
                                                    C.c_void_p(d_seeds_ptr), n, C.c_void_p(d_out_ptr),
                                                    C.c_void_p(stream_ptr)))

    def submit(self, model, dim: int, steps: int, seeds, out: np.ndarray) -> int:
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert out.flags.c_contiguous and out.dtype == np.float64
        t = lib.jne_submit(self._ctx, _model_number(model), dim, steps, seeds.ctypes.data, seeds.size,
                           out.ctypes.data)
        self._check(t)
        return int(t)

    def wait(self, ticket: int) -> None:
        self._check(lib.jne_wait(self._ctx, ticket))

    def eigs_batch_device(self, model, dim: int, steps: int, d_seeds_ptr: int, n: int, d_out_ptr: int,
                          stream_ptr: int = 0) -> None:
        """Raw device pointers (e.g. torch ``tensor.data_ptr()``); enqueue only."""
        self._check(lib.jne_eigs_batch_device(self._ctx, _model_number(model), dim, steps,
                                              C.c_void_p(d_seeds_ptr), n, C.c_void_p(d_out_ptr),
                                              C.c_void_p(stream_ptr)))

    def check_async(self) -> None:
        self._check(lib.jne_check_async(self._ctx))

    def eigs_from_increments(self, model, dB: np.ndarray) -> np.ndarray:
        """dB: (n, steps, dim) C-order == per-run column-major dim x steps (src/rng_matrix.rs:36)."""
        dB = np.ascontiguousarray(dB, dtype=np.float64)
        if dB.ndim != 3:
            raise ValueError("dB must have shape (n, steps, dim)")
        n, steps, dim = dB.shape
        p = num_eigs(_model_number(model), dim)
        out = np.empty((n, p), dtype=np.float64)
        self._check(lib.jne_eigs_from_increments(self._ctx, _model_number(model), dim, steps,
                                                 dB.ctypes.data, n, out.ctypes.data))
        return out

    def eigs_batch_debug(self, model, dim: int, steps: int, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        p = num_eigs(_model_number(model), dim)
        out = np.empty((seeds.size, p), dtype=np.float64)
        mats = np.empty((seeds.size, 2, 16, 16), dtype=np.float64)
        self._check(lib.jne_eigs_batch_debug(self._ctx, _model_number(model), dim, steps,
                                             seeds.ctypes.data, seeds.size, out.ctypes.data, mats.ctypes.data))
        return out, mats[:, 0], mats[:, 1]

    # -- exposed pieces ---------------------------------------------------------------------
    def gen_normal_matrix(self, nrows: int, ncols: int, seed: int) -> np.ndarray:
        """dim x steps matrix (numpy view of the column-major buffer)."""
        buf = np.empty((ncols, nrows), dtype=np.float64)
        self._check(lib.jne_gen_normal_matrix(self._ctx, nrows, ncols, seed & 0xFFFFFFFF, buf.ctypes.data))
        return buf.T

    def brownian_motion_matrix(self, dim: int, steps: int, delta_t: float, seed: int) -> np.ndarray:
        buf = np.empty((steps + 1, dim), dtype=np.float64)
        self._check(lib.jne_brownian_motion_matrix(self._ctx, dim, steps, float(delta_t), seed & 0xFFFFFFFF,
                                                   buf.ctypes.data))
        return buf.T

    def pencil_eigs_batch(self, S1: np.ndarray, S2: np.ndarray) -> np.ndarray:
        """S1: (n, d, p) [sum dB F'], S2: (n, p, p).  Returns (n, p) descending."""
        S1 = np.asarray(S1, dtype=np.float64)
        S2 = np.ascontiguousarray(S2, dtype=np.float64)
        n, d, p = S1.shape
        s1t = np.ascontiguousarray(np.transpose(S1, (0, 2, 1)))  # (n, p, d) row-major == d x p column-major
        out = np.empty((n, p), dtype=np.float64)
        self._check(lib.jne_pencil_eigs_batch(self._ctx, p, d, s1t.ctypes.data, S2.ctypes.data, n, out.ctypes.data))
        return out

    # -- streaming statistics (src/simulation_analyzers.rs:42-81) ---------------------------------------
    def simulate_percentiles(self, model, dim: int, steps: int, num_runs: int, percentiles, first_seed: int = 1):
        """(trace percentiles, max-eig percentiles) of seeds first_seed..first_seed+num_runs-1; nothing but the
        percentiles leaves the GPU."""
        qs = np.ascontiguousarray(percentiles, dtype=np.float64)
        tr = np.empty(qs.size); mx = np.empty(qs.size)
        self._check(lib.jne_simulate_percentiles(self._ctx, _model_number(model), dim, steps, first_seed, num_runs,
                                                 qs.ctypes.data, qs.size, tr.ctypes.data, mx.ctypes.data))
        return tr, mx

    def simulate_percentiles_multi(self, models, dim: int, steps: int, num_runs: int, percentiles, first_seed: int = 1):
        """{model: (trace percentiles, max-eig percentiles)} for every model in `models` from one fused pass."""
        models = sorted(_model_number(m) for m in models)
        qs = np.ascontiguousarray(percentiles, dtype=np.float64)
        tr = np.empty((len(models), qs.size)); mx = np.empty((len(models), qs.size))
        self._check(lib.jne_simulate_percentiles_multi(self._ctx, self._mask(models), dim, steps, first_seed, num_runs,
                                                       qs.ctypes.data, qs.size, tr.ctypes.data, mx.ctypes.data))
        return {m: (tr[i], mx[i]) for i, m in enumerate(models)}

    def percentiles_device(self, d_eigs_ptr: int, n: int, p: int, stride: int, percentiles, stream_ptr: int = 0):
        qs = np.ascontiguousarray(percentiles, dtype=np.float64)
        tr = np.empty(qs.size); mx = np.empty(qs.size)
        self._check(lib.jne_percentiles_device(self._ctx, C.c_void_p(d_eigs_ptr), n, p, stride, qs.ctypes.data,
                                               qs.size, tr.ctypes.data, mx.ctypes.data, C.c_void_p(stream_ptr)))
        return tr, mx

    def fp64_peak_tflops(self, mode: int = 0, ms_target: float = 200.0) -> float:
        v = C.c_double()
        self._check(lib.jne_fp64_peak_tflops(self._ctx, mode, ms_target, C.byref(v)))
        return v.value


_default: Optional[Engine] = None


def default_engine() -> Engine:
    global _default
    if _default is None:
        _default = Engine()
    return _default


# ---- mirrors of the reference's functions ------------------------------------------------------

def calculate_eigenvalues(dim: int, steps: int, seed: int, model) -> list:
    """src/johansen_statistics.rs:59-85 -- one run, eigenvalues descending."""
    return default_engine().eigs_batch(model, dim, steps, [seed])[0].tolist()


def calculate_eigenvalues_parallel(dim: int, steps: int, seeds: Iterable[int], model,
                                   sender: Callable[[int, list], None], quiet: bool = True,
                                   engine: Optional[Engine] = None, batch: int = 1 << 20) -> None:
    """src/data_storage/parallel_compute.rs:14-41 -- for every seed, send (seed, eigenvalues).
    The reference walks 10 000-seed chunks through rayon; here each (large) chunk is one
    jne_submit, overlapped with delivering the previous chunk to ``sender``."""
    eng = engine or default_engine()
    seeds = np.ascontiguousarray(list(seeds) if not isinstance(seeds, np.ndarray) else seeds, dtype=np.uint32)
    p = num_eigs(_model_number(model), dim)
    prev = None
    for a in range(0, seeds.size, batch):
        chunk = seeds[a:a + batch]
        out = np.empty((chunk.size, p), dtype=np.float64)
        ticket = eng.submit(model, dim, steps, chunk, out)
        try:
            if prev is not None:
                for s, row in zip(prev[0].tolist(), prev[1]):
                    sender(s, row.tolist())
        except BaseException:
            # the worker still writes into `out`: join it before the array can be collected, and leave the engine
            # without a pending ticket
            try:
                eng.wait(ticket)
            except JneError:
                pass
            raise
        eng.wait(ticket)
        prev = (chunk, out)
    if prev is not None:
        for s, row in zip(prev[0].tolist(), prev[1]):
            sender(s, row.tolist())


def gen_normal_matrix(nrows: int, ncols: int, seed: int) -> np.ndarray:
    """src/rng_matrix.rs:11-37."""
    return default_engine().gen_normal_matrix(nrows, ncols, seed)


def brownian_motion_matrix(dim: int, steps: int, delta_t: float, seed: int) -> np.ndarray:
    """src/rng_matrix.rs:57-141 with AlongColumns and a zero start column."""
    return default_engine().brownian_motion_matrix(dim, steps, delta_t, seed)

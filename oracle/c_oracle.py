"""ctypes loader for oracle/libjne_oracle.so  --  TEST / BASELINE INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "libjne_oracle.so"


def build(force: bool = False) -> Path:
    if force or not _LIB.exists() or _LIB.stat().st_mtime < (_DIR / "jne_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB


def load() -> C.CDLL:
    build()
    lib = C.CDLL(str(_LIB))
    lib.jne_oracle_eigs_from_increments.restype = C.c_int
    lib.jne_oracle_eigs_from_increments.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.jne_oracle_calculate_eigenvalues.restype = C.c_int
    lib.jne_oracle_calculate_eigenvalues.argtypes = [C.c_int, C.c_size_t, C.c_uint32, C.c_int, C.c_size_t, C.c_void_p]
    lib.jne_oracle_eigs_batch.restype = C.c_int
    lib.jne_oracle_eigs_batch.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p]
    lib.jne_oracle_fast_from_increments.restype = C.c_int
    lib.jne_oracle_fast_from_increments.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.jne_oracle_fast_batch.restype = C.c_int
    lib.jne_oracle_fast_batch.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    lib.jne_oracle_fast_multi_stats.restype = C.c_int
    lib.jne_oracle_fast_multi_stats.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    lib.jne_oracle_gen_normal_matrix.restype = None
    lib.jne_oracle_gen_normal_matrix.argtypes = [C.c_size_t, C.c_size_t, C.c_uint64, C.c_size_t, C.c_void_p]
    return lib


def physical_cores() -> int:
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        return os.cpu_count() or 1


def num_eigs(model: int, dim: int) -> int:
    return dim + 1 if model in (1, 3) else dim


def eigs_from_increments(lib, db: np.ndarray, model: int) -> np.ndarray:
    """db: (T, d) C-order == d x T column-major."""
    db = np.ascontiguousarray(db, dtype=np.float64)
    T, d = db.shape
    out = np.empty(num_eigs(model, d))
    rc = lib.jne_oracle_eigs_from_increments(model, d, T, db.ctypes.data, out.ctypes.data)
    if rc:
        raise FloatingPointError(f"oracle rc={rc}")
    return out


def eigs_batch(lib, model: int, dim: int, steps: int, seeds, threads: int, ncpu: int = None) -> np.ndarray:
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    out = np.empty((seeds.size, num_eigs(model, dim)))
    rc = lib.jne_oracle_eigs_batch(model, dim, steps, seeds.ctypes.data, seeds.size, threads,
                                   ncpu or physical_cores(), out.ctypes.data)
    if rc:
        raise FloatingPointError(f"oracle rc={rc}")
    return out


def gen_normal_matrix(lib, nrows: int, ncols: int, seed: int, ncpu: int = None) -> np.ndarray:
    buf = np.empty((ncols, nrows))
    lib.jne_oracle_gen_normal_matrix(nrows, ncols, seed, ncpu or physical_cores(), buf.ctypes.data)
    return buf.T


def fast_from_increments(lib, db: np.ndarray, model: int) -> np.ndarray:
    """The optimised-CPU variant on caller increments; db: (T, d) C-order."""
    db = np.ascontiguousarray(db, dtype=np.float64)
    T, d = db.shape
    out = np.empty(num_eigs(model, d))
    rc = lib.jne_oracle_fast_from_increments(model, d, T, db.ctypes.data, out.ctypes.data)
    if rc:
        raise FloatingPointError(f"oracle fast rc={rc}")
    return out


def fast_batch(lib, model: int, dim: int, steps: int, seeds, threads: int) -> np.ndarray:
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    out = np.empty((seeds.size, num_eigs(model, dim)))
    rc = lib.jne_oracle_fast_batch(model, dim, steps, seeds.ctypes.data, seeds.size, threads, out.ctypes.data)
    if rc:
        raise FloatingPointError(f"oracle fast rc={rc}")
    return out


def fast_multi_stats(lib, dim: int, steps: int, seeds, threads: int) -> np.ndarray:
    """(n, 5, 2): per seed and model the trace and the largest eigenvalue, all five models from ONE f64-ziggurat path
    per seed (the gate-(2) CPU sampler)."""
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    out = np.empty((seeds.size, 5, 2))
    rc = lib.jne_oracle_fast_multi_stats(dim, steps, seeds.ctypes.data, seeds.size, threads, out.ctypes.data)
    if rc:
        raise FloatingPointError(f"oracle fast multi rc={rc}")
    return out

"""ctypes binding of include/jne_dat.h: batched EIGENVALS_V6 writer / reader / resume scan.

Mirrors the reference's storage interface for the records the hot path produces:
  AppendOnlyWriter            src/data_storage/writer.rs:28-302
  read_append_file            src/data_storage/reader.rs:23-73
  check_append_progress /
  get_remaining_seeds         src/data_storage/progress.rs:11-61
  uleb128::{encode, decode}   src/data_storage/uleb128.rs:68-140
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from .api import JneError, lib

_vp = C.c_void_p
lib.jne_uleb128_encode.restype = C.c_int
lib.jne_uleb128_encode.argtypes = [C.c_uint32, C.POINTER(C.c_uint8)]
lib.jne_uleb128_decode.restype = C.c_int
lib.jne_uleb128_decode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32)]
lib.jne_uleb128_encoded_size.restype = C.c_int
lib.jne_uleb128_encoded_size.argtypes = [C.c_uint32]
lib.jne_dat_expected_file_size.restype = C.c_uint64
lib.jne_dat_expected_file_size.argtypes = [C.c_uint64, C.c_uint32]
lib.jne_dat_open.restype = C.c_int
lib.jne_dat_open.argtypes = [C.c_char_p, C.c_uint8, C.c_uint8, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(_vp)]
lib.jne_dat_append_batch.restype = C.c_int
lib.jne_dat_append_batch.argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint32]
lib.jne_dat_flush.restype = C.c_int
lib.jne_dat_flush.argtypes = [_vp]
lib.jne_dat_finish.restype = C.c_int
lib.jne_dat_finish.argtypes = [_vp]
lib.jne_dat_abandon.restype = None
lib.jne_dat_abandon.argtypes = [_vp]
lib.jne_dat_last_error.restype = C.c_char_p
lib.jne_dat_last_error.argtypes = []
lib.jne_dat_info.restype = C.c_int
lib.jne_dat_info.argtypes = [C.c_char_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                             C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
lib.jne_dat_read.restype = C.c_int
lib.jne_dat_read.argtypes = [C.c_char_p, _vp, _vp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
lib.jne_dat_completed_bitmap.restype = C.c_int
lib.jne_dat_completed_bitmap.argtypes = [C.c_char_p, C.c_uint8, C.c_uint8, C.c_uint32, C.c_uint64, _vp,
                                         C.POINTER(C.c_uint64)]
lib.jne_dat_remaining_seeds.restype = C.c_uint64
lib.jne_dat_remaining_seeds.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64]
lib.jne_run_model_simulation.restype = C.c_int
lib.jne_run_model_simulation.argtypes = [_vp, C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint64, C.c_char_p, C.c_int,
                                         C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_uint64)]
lib.jne_run_models_simulation.restype = C.c_int
lib.jne_run_models_simulation.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_char_p), C.c_int,
                                          C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_uint64)]
lib.jne_dat_batch_begin.restype = C.c_int
lib.jne_dat_batch_begin.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, C.POINTER(_vp)]
lib.jne_dat_batch_fill.restype = C.c_int
lib.jne_dat_batch_fill.argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.c_int]
lib.jne_dat_batch_end.restype = C.c_int
lib.jne_dat_batch_end.argtypes = [_vp, C.c_int]
lib.jne_dat_append_batch_strided_mt.restype = C.c_int
lib.jne_dat_append_batch_strided_mt.argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
lib.jne_dat_append_batch_strided.restype = C.c_int
lib.jne_dat_append_batch_strided.argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint64]


def _check(rc: int) -> None:
    if rc < 0:
        raise JneError(rc, lib.jne_dat_last_error().decode())


def uleb128_encode(value: int) -> bytes:
    buf = (C.c_uint8 * 5)()
    n = lib.jne_uleb128_encode(value, buf)
    return bytes(buf[:n])


def uleb128_decode(data: bytes) -> Tuple[int, int]:
    """(value, bytes used); raises ValueError with the reference's Uleb128Error wording."""
    v = C.c_uint32()
    n = lib.jne_uleb128_decode(data, len(data), C.byref(v))
    if n < 0:
        raise ValueError({-1: "Incomplete ULEB128 encoding", -2: "ULEB128 encoding too long",
                          -3: "ULEB128 value too large for u32"}[n])
    return v.value, n


def uleb128_encoded_size(value: int) -> int:
    return lib.jne_uleb128_encoded_size(value)


def expected_file_size(num_runs: int, eigenvalues_per_run: int) -> int:
    return int(lib.jne_dat_expected_file_size(num_runs, eigenvalues_per_run))


class AppendOnlyWriter:
    """Batched counterpart of the reference's AppendOnlyWriter (src/data_storage/writer.rs)."""

    def __init__(self, path: str, model: int, dim: int, steps: int):
        self._w = _vp()
        existing = C.c_uint64()
        _check(lib.jne_dat_open(str(path).encode(), model, dim, steps, C.byref(existing), C.byref(self._w)))
        self.existing_records = existing.value

    def append_batch(self, seeds, eigs) -> None:
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        eigs = np.ascontiguousarray(eigs, dtype=np.float64)
        if eigs.ndim == 1:
            eigs = eigs.reshape(seeds.size, -1) if seeds.size else eigs.reshape(0, 1)
        assert eigs.shape[0] == seeds.size
        _check(lib.jne_dat_append_batch(self._w, seeds.ctypes.data, eigs.ctypes.data, seeds.size, eigs.shape[1]))

    def append_batch_strided(self, seeds, rows, offset: int, p: int, threads: int = 1) -> None:
        """Append columns [offset, offset + p) of the C-contiguous rows of a fused multi-model batch
        (threads > 1: records encoded by that many host threads, same bytes)."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        assert rows.ndim == 2 and rows.shape[0] == seeds.size and offset + p <= rows.shape[1]
        _check(lib.jne_dat_append_batch_strided_mt(self._w, seeds.ctypes.data, rows.ctypes.data + 8 * offset, seeds.size, p,
                                                   rows.shape[1], int(threads)))

    def batch(self, seeds, p: int) -> "RandomAccessBatch":
        """Reserve the bytes of ``len(seeds)`` records at the end of the file; fill record ranges in any order."""
        return RandomAccessBatch(self, seeds, p)

    def append_eigenvalues(self, seed: int, eigenvalues) -> None:   # the reference's per-record call
        self.append_batch([seed], np.asarray(eigenvalues, dtype=np.float64)[None, :])

    def flush(self) -> None:
        _check(lib.jne_dat_flush(self._w))

    def finish(self) -> None:
        w, self._w = self._w, _vp()
        _check(lib.jne_dat_finish(w))

    def abandon(self) -> None:
        w, self._w = self._w, _vp()
        if w:
            lib.jne_dat_abandon(w)

    def __del__(self):
        try:
            self.abandon()
        except Exception:
            pass


class RandomAccessBatch:
    """jne_dat_batch_*: the records of a batch have their place in the file before any is written (sizes follow from
    the seeds), so several producers fill disjoint ranges in any order; ``end(commit=True)`` makes the batch part of
    the file, ``end(False)`` truncates it away.  Until then the batch starts with an invalid ULEB128 (crash safety)."""

    def __init__(self, writer: AppendOnlyWriter, seeds, p: int):
        self._seeds = np.ascontiguousarray(seeds, dtype=np.uint32)     # must outlive the batch
        self.p = int(p)
        self._b = _vp()
        _check(lib.jne_dat_batch_begin(writer._w, self._seeds.ctypes.data, self._seeds.size, self.p, C.byref(self._b)))

    def fill(self, first: int, rows, offset: int = 0, threads: int = 1) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        assert rows.ndim == 2 and offset + self.p <= rows.shape[1]
        _check(lib.jne_dat_batch_fill(self._b, first, rows.shape[0], rows.ctypes.data + 8 * offset, rows.shape[1], int(threads)))

    def end(self, commit: bool = True) -> None:
        b, self._b = self._b, _vp()
        if b:
            _check(lib.jne_dat_batch_end(b, 1 if commit else 0))

    def __del__(self):
        try:
            self.end(False)
        except Exception:
            pass


def file_info(path: str) -> dict:
    m, d, s, n, p, t = C.c_uint8(), C.c_uint8(), C.c_uint32(), C.c_uint64(), C.c_uint32(), C.c_int()
    _check(lib.jne_dat_info(str(path).encode(), C.byref(m), C.byref(d), C.byref(s), C.byref(n), C.byref(p), C.byref(t)))
    return {"model": m.value, "dim": d.value, "steps": s.value, "records": n.value,
            "eigenvalues_per_run": p.value, "has_trailer": bool(t.value)}


def read_append_file(path: str):
    """(seeds uint32[n], eigenvalues float64[n, p], model, dim, steps) -- src/data_storage/reader.rs:23-73."""
    info = file_info(path)
    n, p = info["records"], max(info["eigenvalues_per_run"], 1)
    seeds = np.empty(n, dtype=np.uint32)
    eigs = np.empty((n, p), dtype=np.float64)
    got = C.c_uint64()
    _check(lib.jne_dat_read(str(path).encode(), seeds.ctypes.data, eigs.ctypes.data, n, p, C.byref(got)))
    return seeds[:got.value], eigs[:got.value], info["model"], info["dim"], info["steps"]


def check_append_progress(path: str, model: int, dim: int, steps: int, num_runs: int):
    """(completed record count, remaining seeds of 1..=num_runs ascending) -- src/data_storage/progress.rs:11-61."""
    bitmap = np.zeros((num_runs + 7) // 8, dtype=np.uint8)
    done = C.c_uint64()
    _check(lib.jne_dat_completed_bitmap(str(path).encode(), model, dim, steps, num_runs, bitmap.ctypes.data, C.byref(done)))
    k = lib.jne_dat_remaining_seeds(bitmap.ctypes.data, num_runs, None, 0)
    rem = np.empty(k, dtype=np.uint32)
    lib.jne_dat_remaining_seeds(bitmap.ctypes.data, num_runs, rem.ctypes.data, k)
    return done.value, rem


def get_filename(model: int, dim: int, steps: int, directory: str = "data") -> str:
    """src/data_storage/simulation.rs:98-122."""
    return f"{directory}/eigenvalues_model{model}_dim{dim}_steps{steps}.dat"


def run_model_simulation(model: int, dim: int, steps: int, num_runs: int, filename: str, quiet: bool = True,
                         devices: Optional[list] = None) -> dict:
    """src/data_storage/parallel_compute.rs:150-232 on the GPU path (C++: csrc/jne_host.cpp)."""
    stats = (C.c_uint64 * 3)()
    if devices is None:
        rc = lib.jne_run_model_simulation(None, model, dim, steps, num_runs, str(filename).encode(), int(quiet), None, 0, stats)
    else:
        arr = (C.c_int * len(devices))(*devices)
        rc = lib.jne_run_model_simulation(None, model, dim, steps, num_runs, str(filename).encode(), int(quiet), arr,
                                          len(devices), stats)
    if rc < 0:
        raise JneError(rc, "run_model_simulation failed (see stderr)")
    return {"completed_before": stats[0], "computed": stats[1], "total_in_file": stats[2]}


def run_models_simulation(models, dim: int, steps: int, num_runs: int, filenames: dict, quiet: bool = True,
                          devices: Optional[list] = None, engine=None) -> dict:
    """The CLI's model loop over one dim (src/main.rs:109-114) as one fused pass (C++: csrc/jne_host.cpp
    run_models_simulation).  filenames: {model: path} for every model in `models`.  Returns {model: stats}."""
    models = sorted(int(m) for m in models)
    mask = 0
    names = (C.c_char_p * 5)()
    for m in models:
        mask |= 1 << m
        names[m] = str(filenames[m]).encode()
    stats = (C.c_uint64 * 15)()
    arr = (C.c_int * len(devices))(*devices) if devices is not None else None
    rc = lib.jne_run_models_simulation(engine._ctx if engine is not None else None, mask, dim, steps, num_runs, names, int(quiet), arr,
                                       len(devices) if devices is not None else 0, stats)
    if rc < 0:
        raise JneError(rc, "run_models_simulation failed (see stderr)")
    return {m: {"completed_before": stats[3 * m], "computed": stats[3 * m + 1], "total_in_file": stats[3 * m + 2]} for m in models}

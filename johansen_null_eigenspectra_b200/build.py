"""In-tree build of the CUDA library (sm_100a only).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libjne.so"
SOURCES = [CSRC / "jne_api.cu", CSRC / "jne_dat.cpp", CSRC / "jne_host.cpp"]
HEADERS = [CSRC / "jne_kernels.cuh", CSRC / "jne_kernels_lane.cuh", CSRC / "jne_rng.cuh", CSRC / "jne_host.hpp",
           PKG_DIR.parent / "include" / "jne.h", PKG_DIR.parent / "include" / "jne_dat.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC,-pthread",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libjne.so")
    return exe


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.exists() and p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False, defines=(), out: Path = None) -> Path:
    """Compile csrc/*.cu|cpp -> libjne.so next to this file.

    `defines` / `out` build a VARIANT library beside it (never loaded by default; select it with the JNE_LIBRARY
    environment variable): e.g. defines=("JNE_RNG_F64",) the validation stream of jne_rng.cuh."""
    target = Path(out) if out is not None else LIB_PATH
    if out is None and not force and not is_stale():
        return LIB_PATH
    srcs = [str(s) for s in SOURCES if s.exists()]
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", str(PKG_DIR.parent / "include"), "-o", str(target), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return target


VARIANTS = {
    # name -> (defines, file): built on request only (tools/, the family regression tests)
    "rng_f64": (("JNE_RNG_F64",), PKG_DIR / "libjne_rng_f64.so"),
}


def build_variant(name: str, verbose: bool = False) -> Path:
    defines, path = VARIANTS[name]
    if path.exists() and all(not p.exists() or p.stat().st_mtime <= path.stat().st_mtime for p in SOURCES + HEADERS):
        return path
    return build_library(force=True, verbose=verbose, defines=defines, out=path)


if __name__ == "__main__":
    import sys
    for name in VARIANTS:
        if f"--{name}" in sys.argv:
            print(build_variant(name, verbose="-v" in sys.argv))
            break
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

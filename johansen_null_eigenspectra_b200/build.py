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
HEADERS = [CSRC / "jne_kernels.cuh", CSRC / "jne_kernels_v2.cuh", CSRC / "jne_kernels_ws.cuh", CSRC / "jne_rng.cuh", CSRC / "jne_host.hpp",
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


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu|cpp -> libjne.so next to this file."""
    if not force and not is_stale():
        return LIB_PATH
    srcs = [str(s) for s in SOURCES if s.exists()]
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(PKG_DIR.parent / "include"), "-o", str(LIB_PATH), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

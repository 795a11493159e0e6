"""The C++ host mirror (csrc/jne_host.hpp) as a reference maintainer would use it: compiled against include/ + libjne.so.
Compile+link is checked on CPU; the run is a GPU test."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "johansen_null_eigenspectra_b200"


def build_demo(tmp_path) -> Path:
    exe = tmp_path / "host_mirror_demo"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(PKG / "csrc"), str(ROOT / "tests" / "cpp" / "host_mirror_demo.cpp"),
           "-L", str(PKG), "-ljne", f"-Wl,-rpath,{PKG}", "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_host_mirror_compiles_and_links(tmp_path):
    exe = build_demo(tmp_path)
    assert exe.exists()
    if not os.path.exists("/dev/nvidia0"):
        # no GPU here: the program must fail loudly (no CPU fallback), not compute
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_host_mirror_matches_python_binding(tmp_path, engine):
    exe = build_demo(tmp_path)
    for dim, steps, n, model in [(2, 103, 5, 0), (12, 400, 7, 3)]:
        r = subprocess.run([str(exe), str(dim), str(steps), str(n), str(model)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        ref = engine.eigs_batch(model, dim, steps, np.arange(1, n + 1, dtype=np.uint32))
        rows = {int(l.split()[0]): [float.fromhex(x) for x in l.split()[1:]] for l in lines if l[0].isdigit()}
        assert sorted(rows) == list(range(1, n + 1))                       # every seed delivered exactly once
        for s in range(1, n + 1):
            assert rows[s] == ref[s - 1].tolist()                          # bit-identical to the Python binding
        single = [float.fromhex(x) for x in [l for l in lines if l.startswith("single")][0].split()[1:]]
        assert single == ref[2].tolist()
        assert "error -1" in r.stdout                                      # Model(7) -> JNE_ERR_INVALID_ARG, no abort

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _ensure_built():
    """A fresh checkout has no libjne.so (built artefacts are git-ignored): build it, and the C oracle, before any
    test imports the package.  build.py is loaded by path so that the package itself is not imported first."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("jne_build", ROOT / "johansen_null_eigenspectra_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.is_stale():
        mod.build_library()
    from oracle import c_oracle
    c_oracle.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    _ensure_built()


@pytest.fixture(scope="session")
def engine():
    """One jne context on cuda:0 for the whole GPU session (raises without a GPU: no fallback)."""
    import johansen_null_eigenspectra_b200 as jne
    eng = jne.Engine([0])
    yield eng
    eng.close()


def eig_tol(ref):
    """Parity gate (1) tolerance: |d| <= 1e-9 |lambda| + 1e-12 lambda_max per run."""
    import numpy as np
    ref = np.asarray(ref)
    return 1e-9 * np.abs(ref) + 1e-12 * np.max(np.abs(ref), axis=-1, keepdims=True)

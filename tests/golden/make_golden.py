"""Generates tests/golden/*.npz with the CPU oracle (oracle/johansen_oracle.py).

The reference holds no golden vectors for this path and cannot be built here ("parity
unpinned"), so these fixtures are ORACLE outputs: seeded increments in,
eigenvalues out.  They pin the oracle against accidental edits and give the GPU parity tests
a frozen target that does not need scipy's dggev at test time.

    python tests/golden/make_golden.py        # rewrites the fixtures (deterministic)

Inputs are regenerated from the stored PCG64 seed, not stored: dB[i] = sqrt(1/T) * z,
z = default_rng(seed).standard_normal((n, T, d)).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import johansen_oracle as orc  # noqa: E402

HERE = Path(__file__).resolve().parent
SEED = 20240601
DIMS = (1, 2, 5, 12)
STEPS = (10, 103, 1000)
N = 4


def increments(seed: int, n: int, steps: int, dim: int) -> np.ndarray:
    z = np.random.default_rng(seed).standard_normal((n, steps, dim))
    return z * np.sqrt(1.0 / steps)


def case_seed(model: int, dim: int, steps: int) -> int:
    return SEED + 1000003 * model + 10007 * dim + steps


def main() -> None:
    out = {}
    for model in range(5):
        for dim in DIMS:
            for steps in STEPS:
                p = orc.num_eigs(model, dim)
                if steps < p + 2:
                    continue  # singular S2: no finite reference answer
                db = increments(case_seed(model, dim, steps), N, steps, dim)
                out[f"m{model}_d{dim}_t{steps}"] = orc.eigs_batch_from_increments(db, model)
    np.savez_compressed(HERE / "eigs_from_increments.npz", **out)
    # config c1 (BASELINE.json configs[0]): model 0, dim 2, T 1000 -- first 64 runs of the stream
    db = increments(SEED, 64, 1000, 2)
    np.savez_compressed(HERE / "c1_model0_dim2_steps1000.npz", eigs=orc.eigs_batch_from_increments(db, 0))
    # full-size spot check: dim 12, T 10 000, all models, 2 runs each
    full = {}
    for model in range(5):
        db = increments(case_seed(model, 12, 10000), 2, 10000, 12)
        full[f"m{model}"] = orc.eigs_batch_from_increments(db, model)
    np.savez_compressed(HERE / "full_dim12_steps10000.npz", **full)
    print("wrote", sorted(p.name for p in HERE.glob("*.npz")))


if __name__ == "__main__":
    main()

"""B200-native engine for the Johansen null-eigenspectra Monte Carlo hot path.

Host-side mirror of the reference's interface for this path (Kuan-Lun/johansen-null-eigenspectra
v0.8.0; citations relative to the reference root):

  * ``calculate_eigenvalues(dim, steps, seed, model)``          src/johansen_statistics.rs:59-85
  * ``calculate_eigenvalues_parallel(dim, steps, seeds, model, sender, quiet)``
                                                                 src/data_storage/parallel_compute.rs:14-41
  * ``JohansenModel``                                            src/johansen_models.rs:6-141
  * ``gen_normal_matrix`` / ``brownian_motion_matrix``          src/rng_matrix.rs:11-37, 57-141

Everything numeric happens in ``libjne.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/jne.h``), reached through ctypes.  There is no CPU fallback: importing this package
without the built library, or creating an ``Engine`` without a B200, raises.
"""
from __future__ import annotations

from .api import (  # noqa: F401
    Engine,
    JneError,
    JohansenModel,
    brownian_motion_matrix,
    calculate_eigenvalues,
    calculate_eigenvalues_parallel,
    default_engine,
    gen_normal_matrix,
    lib,
    num_eigs,
    flops_per_run,
    jacobi_table,
    trend_weight_table,
    version,
)

__all__ = [
    "Engine", "JneError", "JohansenModel", "brownian_motion_matrix", "calculate_eigenvalues",
    "calculate_eigenvalues_parallel", "default_engine", "gen_normal_matrix", "lib", "num_eigs",
    "flops_per_run", "jacobi_table", "trend_weight_table", "version",
]

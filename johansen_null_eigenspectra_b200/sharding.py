"""Seed-list sharding for one-process-per-GPU launches.

Runs are independent pure functions of the seed (reference src/data_storage/parallel_compute.rs:32-39),
so ranks take contiguous slices of the seed list and never exchange data: no collective on the
data path.  torch.distributed is used only for the timing barrier / max-over-ranks in bench.py.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[a, b) of rank's contiguous slice; slices are disjoint, ordered and cover 0..n."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    per = -(-n // world_size)
    a = min(rank * per, n)
    return a, min(a + per, n)


def shard_seeds(seeds: np.ndarray, world_size: int, rank: int) -> np.ndarray:
    a, b = shard_bounds(len(seeds), world_size, rank)
    return seeds[a:b]


def weak_scaling_seeds(runs_per_gpu: int, world_size: int, rank: int, first_seed: int = 1) -> np.ndarray:
    """Weak scaling: every rank gets runs_per_gpu seeds; the union is first_seed..first_seed+N*runs-1
    (the reference numbers runs 1..=num_runs, src/data_storage/progress.rs:58)."""
    a = first_seed + rank * runs_per_gpu
    return np.arange(a, a + runs_per_gpu, dtype=np.uint32)

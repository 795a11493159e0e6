"""world_size-2 gloo test of the N>1 host logic used by bench.py: seed sharding with no data-path
collective, barrier + max-over-ranks timing.  (-m "not gpu")"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from johansen_null_eigenspectra_b200.sharding import shard_bounds, weak_scaling_seeds


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_bounds(n, world, rank)
        seeds = weak_scaling_seeds(50, world, rank)
        # stand-in for the per-rank device time: rank-dependent
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bounds = [None] * world
        dist.all_gather_object(bounds, (a, b, int(seeds[0]), int(seeds[-1])))
        if rank == 0:
            q.put((float(t), bounds))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_max_timing():
    world, n = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, bounds = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 2.0                                   # max over ranks, not rank 0's own time
    assert bounds[0][:2] == (0, 501) and bounds[1][:2] == (501, 1001)
    assert bounds[0][2:] == (1, 50) and bounds[1][2:] == (51, 100)   # disjoint seed ranges, union 1..100


def test_reference_arm_runs_without_gpu():
    """bench.py --impl reference times the CPU port of the reference path (oracle/jne_oracle.c) and needs no GPU:
    one JSON line with the contract's keys; under torchrun only rank 0 prints."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-runs", "32"],
                       cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "runs/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "runs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "dim 12" in d["config"]["workload"]
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-runs", "32"],
                       cwd=root, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]

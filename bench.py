#!/usr/bin/env python
"""bench.py -- Monte Carlo runs/s of the Johansen null-eigenspectra hot path on B200.

Workload (BASELINE.json metric: "Monte Carlo runs/sec at dim=12 T=10k (models 0-4)"): one STEP is
one pass of the hot path over one batch of synthetic input = `--runs` seeds for EACH of the five
models at dim 12, T 10 000.  A run is defined by (model, dim, steps, seed); there is no other input.
The reference's Brownian path depends on (dim, steps, seed) only and its CLI evaluates all five
models on the same seeds (src/rng_matrix.rs:11, src/main.rs:109), so the product path serves the five
runs of a seed from ONE pass over that seed's path (jne_eigs_batch_multi, one launch per step);
`per_model_path` reports the same step done as five independent per-model launches.

  value        device-resident: seeds and eigenvalue buffers live in HBM, launches go on torch's
               current stream, timed with CUDA events, max over ranks
  e2e          the same work through the public host-buffer API (jne_eigs_batch): seeds H2D and
               eigenvalues D2H inside the timed region
  roofline     FP64: F_alg = 2 T [p(p+1)/2 + p d] flop per run x runs per launch /
               average launch duration (CUDA events around each launch), over the measured FP64 peak
  cpu_baseline the CPU restatement of the reference path (oracle/jne_oracle.c, "port": the Rust
               crate cannot be built in this image) on all host cores, bounded sample
  --impl reference   times that CPU port instead (rank 0 only)

Multi-GPU: one process per GPU under torchrun; ranks take disjoint seed ranges (weak scaling), no
collective on the data path; torch.distributed only for the barrier and the max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Monte Carlo runs/sec at dim=12 T=10k (models 0-4)"
MODELS = (0, 1, 2, 3, 4)
NOMINAL_FP64_TFLOPS = 37.2   # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5, help="timed steps K")
    ap.add_argument("--warmup", type=int, default=3, help="untimed warm-up steps W")
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--dim", type=int, default=12)
    ap.add_argument("--T", type=int, default=10000, help="time steps per run (the reference's --steps)")
    # 133 200 = 45 full waves of 148 SMs x 5 resident CTAs x 4 runs per CTA (the fused kernel's geometry)
    ap.add_argument("--runs", type=int, default=133200, help="runs per model per step per GPU")
    ap.add_argument("--cpu-runs", type=int, default=0, help="CPU sample: runs per model (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dat", action="store_true", help="skip the end-to-end figure that includes the .dat files")
    ap.add_argument("--sustain-s", type=float, default=10.0, help="length of the sustained window (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the c2 figures and the in-process multi-device figure")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(dim, T, runs_per_model, threads):
    """Times the CPU port of the reference path (all five models).  Returns (runs/s, total runs, seconds)."""
    from oracle import c_oracle
    lib = c_oracle.load()
    ncpu = c_oracle.physical_cores()
    seeds = np.arange(1, runs_per_model + 1, dtype=np.uint32)
    t0 = time.perf_counter()
    for m in MODELS:
        c_oracle.eigs_batch(lib, m, dim, T, seeds, threads, ncpu)
    dt = time.perf_counter() - t0
    return len(MODELS) * runs_per_model / dt, len(MODELS) * runs_per_model, dt


def cpu_optimised_rate(dim, T, runs_per_model, threads):
    """The optimised CPU variant (one pass, raw moments, dsygv; oracle/jne_oracle.c) -- reported beside the faithful port."""
    from oracle import c_oracle
    lib = c_oracle.load()
    seeds = np.arange(1, runs_per_model + 1, dtype=np.uint32)
    t0 = time.perf_counter()
    for m in MODELS:
        c_oracle.fast_batch(lib, m, dim, T, seeds, threads)
    dt = time.perf_counter() - t0
    return len(MODELS) * runs_per_model / dt, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, on the host cores.
    The Rust crate cannot be built here (no cargo/rustc, un-vendored git dependencies, no system LAPACK), so this is
    the C port oracle/jne_oracle.c (cpu_baseline.kind = "port")."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    runs = args.cpu_runs or max(threads * 256, 1024)     # ~2-3 s of CPU work per step
    for _ in range(max(args.warmup, 0)):
        cpu_reference_rate(args.dim, args.T, max(threads, 8), threads)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        _, n, _ = cpu_reference_rate(args.dim, args.T, runs, threads)
        total += n
    dt = time.perf_counter() - t0
    value = total / dt
    sample = f"{runs} runs x 5 models per step, dim {args.dim}, T {args.T}, seeds 1..{runs}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "runs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"models 0-4, dim {args.dim}, T {args.T}, CPU port of the reference path, {threads} threads"},
        "cpu_baseline": {"value": value, "unit": "runs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "runs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not (ROOT / "johansen_null_eigenspectra_b200" / "libjne.so").exists() and local_rank == 0:
        import __graft_entry__          # built artefacts are git-ignored: build on a fresh checkout
        __graft_entry__.build()
    import johansen_null_eigenspectra_b200 as jne
    from johansen_null_eigenspectra_b200.sharding import weak_scaling_seeds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host-side barrier for the stretch in which rank 0 drives EVERY GPU from one process: a rank waiting in an NCCL
        # barrier spins a kernel on its GPU, and two processes on one device are time-sliced (measured: the in-process
        # figure of 2 GPUs fell to that of one)
        cpu_group = dist.new_group(backend="gloo")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = jne.Engine([local_rank])
    dim, T, R = args.dim, args.T, args.runs
    stream = torch.cuda.current_stream()
    seeds_np = weak_scaling_seeds(R, world, rank)          # disjoint seed ranges per rank
    d_seeds = torch.from_numpy(seeds_np.astype(np.int64)).to(torch.int32).cuda()
    d_out = {m: torch.empty((R, jne.num_eigs(m, dim)), dtype=torch.float64, device="cuda") for m in MODELS}
    width = sum(jne.num_eigs(m, dim) for m in MODELS)
    d_out_multi = torch.empty((R, width), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    model_flops_step = sum(jne.flops_per_run(m, dim, T) for m in MODELS) * R      # five separate runs per seed
    # what the fused kernel actually computes per seed: sum c c' (sym) + sum c z' + 5 deterministic cross
    # moments per row (sum c, w1 c, w2 c, w1 z, w2 z)
    fused_flops_run = 2.0 * T * (dim * (dim + 1) / 2 + dim * dim + 5 * dim)

    ev_pairs = []      # fused launches
    ev_pm = []         # per-model launches

    def device_step(record):
        flush.zero_()                                      # L2 flush between steps
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        eng.eigs_batch_multi_device(MODELS, dim, T, d_seeds.data_ptr(), R, d_out_multi.data_ptr(), stream.cuda_stream)
        if record:
            e1.record(stream)
            ev_pairs.append((e0, e1))

    def per_model_step(record):
        flush.zero_()
        for m in MODELS:
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            eng.eigs_batch_device(m, dim, T, d_seeds.data_ptr(), R, d_out[m].data_ptr(), stream.cuda_stream)
            if record:
                e1.record(stream)
                ev_pm.append((m, e0, e1))

    # FP64 roofline denominator: measured here (MEASURED_PEAKS.json carries no FP64 figure)
    peak_dfma = eng.fp64_peak_tflops(0, 300.0)
    peak_dmma = eng.fp64_peak_tflops(1, 300.0)
    peak = max(peak_dfma, peak_dmma)

    for _ in range(args.warmup):
        device_step(False)
    eng.check_async()
    launches0 = eng.launch_count
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        t_start.record(stream)
        for _ in range(args.steps):
            device_step(True)
        t_end.record(stream)
        barrier()
    gpu_launches = eng.launch_count - launches0   # our kernels only (torch's L2-flush fill is not ours)
    eng.check_async()
    ms = t_start.elapsed_time(t_end)
    fused_ms = [e0.elapsed_time(e1) for e0, e1 in ev_pairs]
    # the same step as five independent per-model launches (reported beside the headline, not part of it)
    per_model_step(False)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        per_model_step(True)
    p1.record(stream)
    barrier()
    eng.check_async()
    pm_ms = p0.elapsed_time(p1)
    kern_ms = {m: [] for m in MODELS}
    for m, e0, e1 in ev_pm:
        kern_ms[m].append(e0.elapsed_time(e1))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_runs = len(MODELS) * R * args.steps * world
    value = total_runs / (ms_max * 1e-3)

    # ---- sustained window: the same device-resident step back to back for >= --sustain-s seconds, clocks sampled
    #      (shows whether the short timed region above survives power management) ----
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(1, int(np.ceil(args.sustain_s / (ms / args.steps * 1e-3))))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as sus_clocks:
            s0.record(stream)
            for _ in range(n_sus):
                device_step(False)
            s1.record(stream)
            barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sustained = {"value": len(MODELS) * R * n_sus * world / (float(ts.item()) * 1e-3), "unit": "runs/s", "steps": n_sus,
                     "seconds": float(ts.item()) * 1e-3, "clocks": sus_clocks.summary()}
        eng.check_async()

    # ---- e2e: host buffers through the public API (H2D seeds + D2H eigenvalues inside) ----
    eng.eigs_batch_multi(MODELS, dim, T, seeds_np[: min(R, 4096)])
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        out = eng.eigs_batch_multi(MODELS, dim, T, seeds_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_runs / float(te.item())
    # ---- e2e incl. the .dat files: the CLI's model loop for this dim as one fused job (run_models_simulation):
    #      resume scan, GPU batches, batched EIGENVALS_V6 encode + write of five files, trailers ----
    dat_value = None
    dat_bytes = 0
    dat_error = None
    if rank == 0 and not args.no_dat:
        import shutil, tempfile
        from johansen_null_eigenspectra_b200 import dat as jdat
        n_dat = R * max(1, args.steps)
        need_bytes = n_dat * (width * 8 + 5 * 6) * 2          # five files plus head-room
        base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need_bytes else None
        tmpdir = tempfile.mkdtemp(prefix="jne_bench_", dir=base)
        try:
            names = {m: os.path.join(tmpdir, f"eigenvalues_model{m}_dim{dim}_steps{T}.dat") for m in MODELS}
            d0 = time.perf_counter()
            st = jdat.run_models_simulation(MODELS, dim, T, n_dat, names, quiet=True, engine=eng)
            dat_s = time.perf_counter() - d0
            assert all(st[m]["total_in_file"] == n_dat for m in MODELS)
            dat_bytes = sum(os.path.getsize(f) for f in names.values())
            dat_value = len(MODELS) * n_dat / dat_s
        except Exception as exc:                                # a full scratch disk must not take the bench line down
            dat_error = f"{type(exc).__name__}: {exc}"
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
    # ---- rank 0 extras: BASELINE config c2 (dim 5, T 5 000) on this GPU, and ONE in-process context over every
    #      visible GPU with host buffers (what the reference-side FFI shim creates: jne_init(NULL, 0)) ----
    extras = {}
    if rank == 0 and not args.no_extras:
        def timed(models, dim2, T2, n2, reps=3):
            ds = torch.arange(1, n2 + 1, dtype=torch.int32, device="cuda")
            do = torch.empty((n2, sum(jne.num_eigs(m, dim2) for m in models)), dtype=torch.float64, device="cuda")
            best = 1e30
            for _ in range(reps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if len(models) == 1:
                    eng.eigs_batch_device(models[0], dim2, T2, ds.data_ptr(), n2, do.data_ptr(), stream.cuda_stream)
                else:
                    eng.eigs_batch_multi_device(models, dim2, T2, ds.data_ptr(), n2, do.data_ptr(), stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            eng.check_async()
            return best
        n2 = 1 << 20
        c2 = {"workload": f"README example (BASELINE configs[1]): dim 5, T 5000, {n2} seeds per launch, device-resident, best of 3"}
        t = timed(MODELS, 5, 5000, n2)
        c2["fused_runs_per_s"] = len(MODELS) * n2 / (t * 1e-3)
        pm = {m: timed([m], 5, 5000, n2) for m in MODELS}
        c2["per_model_runs_per_s"] = {str(m): n2 / (pm[m] * 1e-3) for m in MODELS}
        c2["per_model_frac_of_fp64_peak"] = {str(m): jne.flops_per_run(m, 5, 5000) * n2 / (pm[m] * 1e-3) / 1e12 / peak for m in MODELS}
        seeds2 = np.arange(1, n2 + 1, dtype=np.uint32)
        buf2 = np.zeros((n2, sum(jne.num_eigs(m, 5) for m in MODELS)))    # touched: a caller's reused buffer, not fresh pages
        eng.eigs_batch_multi(MODELS, 5, 5000, seeds2[:200000], out=buf2[:200000])
        h0 = time.perf_counter()
        eng.eigs_batch_multi(MODELS, 5, 5000, seeds2, out=buf2)
        c2["e2e_fused_runs_per_s"] = len(MODELS) * n2 / (time.perf_counter() - h0)      # host buffers, H2D + D2H inside
        extras["c2"] = c2
    if world > 1:
        dist.barrier(group=cpu_group)   # the other ranks idle (on the host) while rank 0 drives every GPU from one process
    if rank == 0 and not args.no_extras:
        try:
            nvis = torch.cuda.device_count()
            eng_all = jne.Engine(None)
            seeds_all = np.arange(1, R * nvis + 1, dtype=np.uint32)
            buf_all = np.empty((seeds_all.size, width), dtype=np.float64)
            eng_all.eigs_batch_multi(MODELS, dim, T, seeds_all[: 4096 * nvis])
            i0 = time.perf_counter()
            reps_ip = max(1, min(args.steps, 5))
            for _ in range(reps_ip):
                eng_all.eigs_batch_multi(MODELS, dim, T, seeds_all, out=buf_all)
            ip_s = time.perf_counter() - i0
            extras["e2e_inprocess"] = {"value": len(MODELS) * seeds_all.size * reps_ip / ip_s, "unit": "runs/s", "devices": nvis,
                                       "note": "one jne context over every visible GPU (jne_init(NULL, 0)), host buffers, seed-sharded inside the library"}
            eng_all.close()
        except Exception as exc:
            extras["e2e_inprocess"] = {"value": None, "error": f"{type(exc).__name__}: {exc}"}
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.barrier(group=cpu_group)
    h2d = 4 * R
    d2h = 8 * R * width
    tp = torch.tensor([pm_ms], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    pm_value = total_runs / (float(tp.item()) * 1e-3)

    if rank == 0:
        # dominant kernel = the fused per-run kernel in multi-model mode (one launch per step)
        dur_ms = float(np.mean(fused_ms))
        achieved = fused_flops_run * R / (dur_ms * 1e-3) / 1e12
        pm_dur_ms = sum(float(np.mean(kern_ms[m])) for m in MODELS)
        per_model = {str(m): round(jne.flops_per_run(m, dim, T) * R / (float(np.mean(kern_ms[m])) * 1e-3) / 1e12, 3)
                     for m in MODELS}
        # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this very
        # launch shape (tools/ncu_target.py --n R; tools/ncu_traffic.py writes the file).  Only quoted when the capture
        # was taken at the same seeds-per-launch and configuration; never extrapolated.
        traffic, traffic_src = None, None
        tfile = ROOT / "profiles" / "r2_traffic_fused_d12_jne2.json"
        if tfile.exists():
            try:
                tj = json.loads(tfile.read_text())
                if tj.get("seeds_in_launch") == R and tj.get("dim") == dim and tj.get("steps") == T:
                    traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                    traffic_src = f"profiles/{tfile.name} ({tj.get('report')})"
            except Exception:
                traffic = None
        contraction_flops_run = 2.0 * T * (dim * (dim + 1) / 2 + dim * dim)
        line = {
            "metric": METRIC, "value": value, "unit": "runs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"models 0-4, dim {dim}, T {T}, {R} runs per model per step per GPU, seeds {int(seeds_np[0])}..",
                "runs_per_step": len(MODELS) * R * world, "rng": "stream JNE2: philox4x32-10-keyed xoshiro128++ substreams + fp32 box-muller, in registers",
                "l2": "256 MiB buffer written between steps (inputs are 4 B per run; the path is FP64-bound)",
                "parallelism": f"seed-sharded x{world}, no collective",
            },
            "roofline": {
                # the FP64 tensor pipe (DMMA.8x8x4; scalar DFMA shares the same datapath and the same cap)
                "bound": "tensor", "precision": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak,
                "frac_definition": "fused pass, its own flop count (incl. the five trend moments per row)",
                "frac_contraction_only": contraction_flops_run * R / (dur_ms * 1e-3) / 1e12 / peak,
                "frac_per_model_s8d": model_flops_step / (pm_dur_ms * 1e-3) / 1e12 / peak,
                "frac_per_model_s8d_definition": "SURVEY section 8(d): F_alg = 2T[p(p+1)/2 + p d] per run, five per-model launches "
                                                 "(per_model_path); the fused pass serves the same runs with 4.1x fewer flops",
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured in this run: register-resident DFMA / DMMA m8n8k4 chains (jne_fp64_peak_tflops); "
                               "MEASURED_PEAKS.json holds no FP64 figure",
                "peak_dfma": peak_dfma, "peak_dmma": peak_dmma, "peak_nominal": NOMINAL_FP64_TFLOPS,
                "kernel": "jne_run_kernel<12,2,rng,multi> (5 models per path)",
                "flops_per_launch": fused_flops_run * R,
                "flops_model": "2T[d(d+1)/2 + d^2 + 5d] per seed: what the fused pass computes "
                               "(sum cc', sum cz', five deterministic cross moments per row)",
                "reference_flops_per_launch": model_flops_step,
                "algorithmic_speedup": model_flops_step / (fused_flops_run * R),
                "kernel_share_of_step": dur_ms * args.steps / ms * 1.0 if ms > 0 else None,
            },
            "per_model_path": {
                "value": pm_value, "unit": "runs/s", "launches_per_step": len(MODELS),
                "achieved_tflops": model_flops_step / (pm_dur_ms * 1e-3) / 1e12,
                "frac_of_fp64_peak": model_flops_step / (pm_dur_ms * 1e-3) / 1e12 / peak,
                "achieved_tflops_per_model": per_model,
                "note": "five independent launches, one per model, F_alg = 2T[p(p+1)/2 + p d] per run",
            },
            "e2e": {"value": e2e_value, "unit": "runs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks.summary(),
        }
        if sustained is not None:
            line["sustained"] = sustained
        line.update(extras)
        if dat_error is not None:
            line["e2e_dat"] = {"value": None, "unit": "runs/s", "error": dat_error}
        if dat_value is not None:
            line["e2e_dat"] = {
                "value": dat_value, "unit": "runs/s", "file_bytes": dat_bytes,
                "note": f"rank 0: run_models_simulation of {R * max(1, args.steps)} seeds x 5 models into five EIGENVALS_V6 files "
                        "(tmpfs when it has room, else the temp dir), resume scan + GPU + batched encode/write + trailers inside",
            }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            runs = args.cpu_runs or max(threads * 1024, 4096)      # ~10 s of CPU work on the box's cores
            cpu_value, n_cpu, cpu_s = cpu_reference_rate(dim, T, runs, threads)
            line["cpu_baseline"] = {
                "value": cpu_value, "unit": "runs/s", "cores": threads, "kind": "port",
                "sample": f"{runs} runs x 5 models, dim {dim}, T {T}, {cpu_s:.1f} s on {threads} threads "
                          "(C port of the reference path incl. xoshiro256++/ziggurat and LAPACK dggev)",
            }
            opt_value, opt_s = cpu_optimised_rate(dim, T, runs, threads)
            line["cpu_optimised"] = {
                "value": opt_value, "unit": "runs/s", "cores": threads, "kind": "port-optimised",
                "sample": f"{runs} runs x 5 models, {opt_s:.1f} s: one pass, raw moments + Schur complements, LAPACK dsygv "
                          "(not the reference's algorithm; shown so the speed-up is not flattered by its temporaries)",
            }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

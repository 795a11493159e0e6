#!/usr/bin/env python
"""Multi-GPU evidence in ONE visit (`gpurun --gpus 8 -- python tools/multi_gpu_suite.py`), through the in-process
multi-device context -- what the reference-side FFI creates (jne_init(NULL, 0)):
  c3   BASELINE configs[2]: the default sweep dim 1..12, T 10 000, 10^7 runs, models 0-4, on all GPUs; one
       jne_simulate_percentiles_multi call per dim (60 cells: 90/95/99 % of trace and max-eig), wall time incl. quantiles
  c4   BASELINE configs[3]: model 4, dim 12, T 10 000, 10^7 runs -- strong scaling over 1 / 2 / 4 / 8 devices, host buffers
  dat  the fused five-model job incl. its five .dat files (run_models_simulation) on all GPUs
Usage: python tools/multi_gpu_suite.py [c3] [c4] [dat] [--runs N]"""
import os, shutil, sys, tempfile, time
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
from johansen_null_eigenspectra_b200 import dat
from tests.gate2_common import MHM95_TRACE, MHM95_MAX

args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3", "c4", "dat"]
runs = int(sys.argv[sys.argv.index("--runs") + 1]) if "--runs" in sys.argv else 10_000_000
ndev = jne.lib.jne_device_count()
print(f"# {ndev} visible GPUs; {jne.version()}", flush=True)

if "c3" in args:
    eng = jne.Engine(None)
    qs = [0.90, 0.95, 0.99]
    eng.simulate_percentiles_multi(range(5), 12, 10000, 4096 * ndev, qs)     # warm-up: tables, allocations
    print(f"# c3: default sweep dim 1..12, T 10 000, {runs} runs per (dim, model), models 0-4, {ndev} GPUs in one context;")
    print("#     95 % quantile of the trace (relative difference to MacKinnon-Haug-Michelis) | of max-eig; seconds per dim incl. the quantiles")
    t_all = time.perf_counter()
    worst = 0.0
    for dim in range(1, 13):
        t0 = time.perf_counter()
        res = eng.simulate_percentiles_multi(range(5), dim, 10000, runs, qs)
        dt = time.perf_counter() - t0
        cells = []
        for m in range(5):
            tr, mx = res[m][0][1], res[m][1][1]
            rt, rm = tr / MHM95_TRACE[m][dim - 1] - 1, mx / MHM95_MAX[m][dim - 1] - 1
            worst = max(worst, abs(rt), abs(rm))
            cells.append(f"{tr:9.4f} ({100 * rt:+.3f} %) | {mx:8.4f} ({100 * rm:+.3f} %)")
        print(f"dim {dim:2d}  {dt:6.3f} s  " + "  ".join(cells), flush=True)
    total = time.perf_counter() - t_all
    print(f"# c3 total wall {total:.2f} s for {12 * 5 * runs / 1e6:.0f} M runs = {12 * 5 * runs / total / 1e6:.1f} M runs/s incl. quantiles; worst |rel| {100 * worst:.3f} %")
    eng.close()

if "c4" in args:
    print(f"# c4: model 4, dim 12, T 10 000, {runs} runs, strong scaling (jne_eigs_batch, host buffers, one context over N devices)")
    seeds = np.arange(1, runs + 1, dtype=np.uint32)
    base = None
    ref = None
    n_list = [n for n in (1, 2, 4, 8) if n <= ndev]
    for n in n_list:
        eng = jne.Engine(list(range(n)))
        eng.eigs_batch(4, 12, 10000, seeds[: 8192 * n])
        best = 1e30
        for rep in range(2):
            t0 = time.perf_counter()
            out = eng.eigs_batch(4, 12, 10000, seeds)
            best = min(best, time.perf_counter() - t0)
        if ref is None:
            ref, base = out[:100000].copy(), best
        same = np.array_equal(out[:100000], ref)
        print(f"N {n}: {best:7.3f} s  {runs / best / 1e6:7.3f} M runs/s  speed-up {base / best:5.2f}  efficiency {base / best / n:5.3f}  "
              f"bit-identical to N=1: {same}", flush=True)
        eng.close()
        del out

if "dat" in args:
    eng = jne.Engine(None)
    base_dir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    print(f"# dat: fused five-model job incl. five EIGENVALS_V6 files, dim 12, T 10 000, {ndev} GPUs; tmpfs free "
          f"{shutil.disk_usage(base_dir or '/tmp').free / 1e9:.0f} GB")
    for n in (2_000_000, 8_000_000):
        d = tempfile.mkdtemp(prefix="jne_job_", dir=base_dir)
        try:
            names = {m: os.path.join(d, f"eigenvalues_model{m}_dim12_steps10000.dat") for m in range(5)}
            t0 = time.perf_counter()
            dat.run_models_simulation(range(5), 12, 10000, n, names, quiet=True, engine=eng)
            dt = time.perf_counter() - t0
            size = sum(os.path.getsize(f) for f in names.values())
            t1 = time.perf_counter()
            st2 = dat.run_models_simulation(range(5), 12, 10000, n, names, quiet=True, engine=eng)   # resume scan only
            dt2 = time.perf_counter() - t1
            info = dat.file_info(names[3])
            print(f"n {n}: {5 * n / dt / 1e6:.2f} M runs/s incl. files ({dt:.3f} s, {size / 1e6:.0f} MB, {size / dt / 1e9:.2f} GB/s); "
                  f"rescan of the complete files {dt2 * 1e3:.0f} ms (computed {st2[0]['computed']}); model 3: {info['records']} records, trailer {info['has_trailer']}", flush=True)
            if n == 2_000_000:      # byte-identity with the serial single-threaded writer on a sample of the same job
                one = jne.Engine([0])
                m0 = os.path.join(d, "serial_model1.dat")
                dat.run_model_simulation(1, 12, 10000, 50_000, m0, quiet=True, devices=[0])
                s1, e1, *_ = dat.read_append_file(m0)
                s2, e2, *_ = dat.read_append_file(names[1])
                order = np.argsort(s2)
                print(f"   records of seeds 1..50000 equal to the single-GPU serial job: {np.array_equal(e2[order][:50000], e1[np.argsort(s1)])}")
                one.close()
        finally:
            shutil.rmtree(d, ignore_errors=True)
    eng.close()

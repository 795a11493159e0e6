"""End-to-end rate of the fused five-model job incl. the five .dat files (run_models_simulation) for several job sizes."""
import os, shutil, sys, tempfile, time
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
from johansen_null_eigenspectra_b200 import dat
import torch
eng = jne.Engine(list(range(torch.cuda.device_count())))
print("devices:", torch.cuda.device_count())
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
print("tmpfs free GB:", shutil.disk_usage(base or "/tmp").free / 1e9)
for n in (666000, 2000000, 8000000):
    d = tempfile.mkdtemp(prefix="jne_job_", dir=base)
    try:
        names = {m: os.path.join(d, f"eigenvalues_model{m}_dim12_steps10000.dat") for m in range(5)}
        t0 = time.perf_counter()
        st = dat.run_models_simulation(range(5), 12, 10000, n, names, quiet=True, engine=eng)
        dt = time.perf_counter() - t0
        size = sum(os.path.getsize(f) for f in names.values())
        t1 = time.perf_counter()
        st2 = dat.run_models_simulation(range(5), 12, 10000, n, names, quiet=True, engine=eng)   # resume scan only
        dt2 = time.perf_counter() - t1
        print(f"n {n}: {5*n/dt/1e6:.2f}M runs/s ({dt:.3f} s, {size/1e6:.0f} MB); rescan of the complete files {dt2*1e3:.1f} ms, computed {st2[0]['computed']}", flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)

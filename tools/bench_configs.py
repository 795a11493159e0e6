"""Throughput of every BASELINE.json config on one B200 (device-resident, CUDA-event timed).  Not the driver's
bench.py contract -- evidence for profiles/: one JSON line per config."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne

eng = jne.Engine([0])
st = torch.cuda.current_stream()
peak = max(eng.fp64_peak_tflops(0, 200.0), eng.fp64_peak_tflops(1, 200.0))


def time_multi(models, dim, T, n, reps=3):
    seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
    width = sum(jne.num_eigs(m, dim) for m in models)
    out = torch.empty((n, width), dtype=torch.float64, device="cuda")
    best = 1e30
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        eng.eigs_batch_multi_device(models, dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    eng.check_async()
    return best


def report(name, models, dim, T, n, note=""):
    ms = time_multi(models, dim, T, n)
    runs = n * len(models)
    flops = sum(jne.flops_per_run(m, dim, T) for m in models) * n
    print(json.dumps({"config": name, "models": list(models), "dim": dim, "T": T, "seeds_per_launch": n,
                      "runs_per_s": runs / (ms * 1e-3), "ms": ms, "model_tflops": flops / (ms * 1e-3) / 1e12,
                      "fp64_peak_tflops": peak, "note": note}), flush=True)


# dim <= 6 runs one THREAD per run (lane family): a launch needs ~10^6 seeds to fill several waves of the GPU
report("c1 model 0 dim 2 T 1000 (100 000 runs)", [0], 2, 1000, 100000)
report("c2 dim 5 T 5000 models 0-4 (fused pass)", range(5), 5, 5000, 1 << 20)
for m in range(5):
    report(f"c2 dim 5 T 5000 model {m} alone", [m], 5, 5000, 1 << 20)
for d in range(1, 13):
    report(f"c3 sweep dim {d} T 10000 models 0-4 (fused pass)", range(5), d, 10000, 1 << 20 if d <= 6 else 1 << 17)
for m in range(5):
    report(f"c4 dim 12 T 10000 model {m} alone", [m], 12, 10000, 1 << 17)
report("c4 dim 12 T 10000 models 0-4 (fused pass)", range(5), 12, 10000, 1 << 17)
report("c5 dim 12 T 100000 models 3,4 (fused pass)", [3, 4], 12, 100000, 1 << 14)
for m in (3, 4):
    report(f"c5 dim 12 T 100000 model {m} alone", [m], 12, 100000, 1 << 14)

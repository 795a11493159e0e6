"""Wave latency of the fused kernel: 1, 2, 3, 6, 45 waves (2960 seeds per wave), T = 10000 and T = 32, dim 12."""
import os, sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
st = torch.cuda.current_stream()
eng = jne.Engine([0])
for models in ([0], [0, 1, 2, 3, 4]):
  for T in (10000, 32):
    for waves in (1, 2, 3, 6, 45):
        n = 2960 * waves
        seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
        out = torch.empty((n, 62), dtype=torch.float64, device="cuda")
        best = 1e9
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.eigs_batch_multi_device(models, 12, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
            torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        print(f"models {models} T {T} waves {waves}: {best:.3f} ms, {best/waves*1e-3*1.965e9:.0f} cycles/wave, {n/best/1e3:.3f}M seeds/s", flush=True)

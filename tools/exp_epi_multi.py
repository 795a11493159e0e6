"""Epilogue cost of the fused five-model kernel: time it at tiny T and at T = 10000 (dim 12)."""
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); n = 133200
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
for dim in (12,):
    for models in ([0], [1], [2], [3], [4], [0, 1, 2, 3, 4]):
        for T in (32, 10000):
            out = torch.empty((n, 62), dtype=torch.float64, device="cuda")
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.eigs_batch_multi_device(models, dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
                torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
            print(f"dim {dim} models {models} T {T}: {n/best/1e3:.3f}M seeds/s, {best*1e-3*1.965e9*592/n:.0f} SMSP-cycles/seed")
eng.check_async()

"""Cost of the per-run epilogue (assembly + Cholesky + Jacobi): time the kernel at tiny T."""
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); n = 1 << 19
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
for dim in (12, 5):
    for model in (0, 3, 4):
        for T in (32, 64, 10000):
            nn = n if T < 1000 else n // 4
            out = torch.empty((nn, 16), dtype=torch.float64, device="cuda")
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.eigs_batch_device(model, dim, T, seeds.data_ptr(), nn, out.data_ptr(), st.cuda_stream); e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"dim {dim} model {model} T {T}: {nn/ms/1e3:.3f}M runs/s, {ms*1e-3*1.965e9*592/nn:.0f} SMSP-cycles/run")
eng.check_async()

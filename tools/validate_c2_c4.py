"""BASELINE configs c2 (dim 5, T 5000, 10^6 runs, models 0-4, fused pass) and c4 (model 4 alone, dim 12, T 10000, 10^7 runs)
at full size on one GPU, with the 95 % trace quantile next to the published asymptotic value."""
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); st = torch.cuda.current_stream()
pub5 = {0: 60.06141, 1: 76.97277, 2: 69.81889, 3: 88.80380, 4: 79.34145}
n = 1_000_000
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
widths = [jne.num_eigs(m, 5) for m in range(5)]
out = torch.empty((n, sum(widths)), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st); eng.eigs_batch_multi_device(range(5), 5, 5000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record(st)
torch.cuda.synchronize(); eng.check_async(); s = e0.elapsed_time(e1) * 1e-3
print(f"c2: dim 5, T 5000, {n} seeds x models 0-4: {s * 1e3:.1f} ms kernel time, {5 * n / s / 1e6:.1f} M runs/s")
off = 0
for m in range(5):
    q = float(torch.quantile(out[:, off:off + widths[m]].sum(dim=1), 0.95)); off += widths[m]
    print(f"  model {m}: trace 95 % {q:.4f} (published {pub5[m]}, {100 * (q / pub5[m] - 1):+.3f} %)")
del out
n = 10_000_000
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
out = torch.empty((n, 12), dtype=torch.float64, device="cuda")
e0.record(st); eng.eigs_batch_device(4, 12, 10000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record(st)
torch.cuda.synchronize(); eng.check_async(); s = e0.elapsed_time(e1) * 1e-3
q = float(torch.quantile(out.sum(dim=1), 0.95))
print(f"c4: model 4 alone, dim 12, T 10000, {n} runs: {s:.2f} s kernel time, {n / s / 1e6:.3f} M runs/s = {jne.flops_per_run(4, 12, 10000) * n / s / 1e12:.1f} TFLOP/s algorithmic; trace 95 % {q:.4f} (published 358.7184, {100 * (q / 358.7184 - 1):+.3f} %)")

"""BASELINE config c5 at full size: dim 12, T = 100 000, 10^6 runs, models 3 and 4 (fused pass), with the 95 % quantiles
of the trace and maximum-eigenvalue statistics next to the published asymptotic values (finite-T bias ~ 1/T)."""
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
pub = {3: (374.9076, 80.87025), 4: (358.7184, 79.97193)}
n, T, dim = 1_000_000, 100_000, 12
eng = jne.Engine([0]); st = torch.cuda.current_stream()
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
out = torch.empty((n, 13 + 12), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st); eng.eigs_batch_multi_device([3, 4], dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record(st)
torch.cuda.synchronize(); eng.check_async()
s = e0.elapsed_time(e1) * 1e-3
print(f"c5: dim 12, T 100000, {n} seeds x models 3,4: {s:.2f} s kernel time, {2 * n / s / 1e6:.3f} M runs/s")
off = 0
for m, w in ((3, 13), (4, 12)):
    tr = float(torch.quantile(out[:, off:off + w].sum(dim=1), 0.95)); mx = float(torch.quantile(out[:, off], 0.95)); off += w
    print(f"model {m}: trace 95 % {tr:.4f} (published {pub[m][0]}, {100 * (tr / pub[m][0] - 1):+.3f} %), max-eig 95 % {mx:.4f} (published {pub[m][1]}, {100 * (mx / pub[m][1] - 1):+.3f} %)")

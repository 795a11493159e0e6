"""Short horizons on the tensor family (dims >= 7): M seeds/s and useful lane-steps of the segment rule (seg_len_for)."""
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); st = torch.cuda.current_stream()
def t(models, dim, T, n, reps=3):
    seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
    out = torch.empty((n, sum(jne.num_eigs(m, dim) for m in models)), dtype=torch.float64, device="cuda")
    best = 1e30
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        eng.eigs_batch_multi_device(models, dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return n / best / 1e3
print("# fused five-model pass, device-resident, best of 3: M seeds/s and M seed-steps/s")
for dim in (12, 8):
    for T in (100, 200, 400, 960, 1000, 2049, 5000, 10000):
        n = max(2960 * 4, int(133200 * 10000 / T) // 2960 * 2960)
        r = t(range(5), dim, T, min(n, 2960 * 300))
        print(f"dim {dim:2d}  T {T:6d}  {r:9.3f} M seeds/s  {r * T / 1e3:8.2f} G seed-steps/s", flush=True)

"""The product's purpose, end to end: 95 % quantiles of the trace statistic (sum of eigenvalues) for models 0-4, dim 1..12
at T = 10 000 from the fused GPU pass, next to the published asymptotic critical values of MacKinnon, Haug & Michelis
(1999) as reported by standard econometrics packages (cases I-V <-> models 0-4).  Finite-T bias at T = 10 000 and the
Monte Carlo error at 10^6 runs are both of order 0.1 %."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne

MHM95 = {   # trace test, 5 % level, dim = number of common trends 1..12
    0: [4.129906, 12.32090, 24.27596, 40.17493, 60.06141, 83.93712, 111.7805, 143.6691, 179.5098, 219.4016, 263.2603, 311.1288],
    1: [9.164546, 20.26184, 35.19275, 54.07904, 76.97277, 103.8473, 134.6780, 169.5991, 208.4374, 251.2650, 298.1594, 348.9784],
    2: [3.841466, 15.49471, 29.79707, 47.85613, 69.81889, 95.75366, 125.6154, 159.5297, 197.3709, 239.2354, 285.1425, 334.9837],
    3: [12.51798, 25.87211, 42.91525, 63.87610, 88.80380, 117.7082, 150.5585, 187.4701, 228.2979, 273.1889, 322.0692, 374.9076],
    4: [3.841466, 18.39771, 35.01090, 55.24578, 79.34145, 107.3466, 139.2753, 175.1715, 215.1232, 259.0294, 306.8944, 358.7184],
}
MHM95_MAX = {   # maximum-eigenvalue test, 5 % level
    0: [4.129906, 11.22480, 17.79730, 24.15921, 30.43961, 36.63019, 42.77219, 48.87720, 54.96577, 61.03407, 67.07555, 73.09094],
    1: [9.164546, 15.89210, 22.29962, 28.58808, 34.80587, 40.95680, 47.07897, 53.18784, 59.24000, 65.30016, 71.33542, 77.38180],
    2: [3.841466, 14.26460, 21.13162, 27.58434, 33.87687, 40.07757, 46.23142, 52.36261, 58.43354, 64.50472, 70.53513, 76.57843],
    3: [12.51798, 19.38704, 25.82321, 32.11832, 38.33101, 44.49720, 50.59985, 56.70519, 62.75215, 68.81206, 74.83748, 80.87025],
    4: [3.841466, 17.14769, 24.25202, 30.81507, 37.16359, 43.41977, 49.58633, 55.72819, 61.80550, 67.90393, 73.94036, 79.97193],
}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
eng = jne.Engine([0])
st = torch.cuda.current_stream()
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
print(f"trace statistic, 95 % quantile, {n} runs per cell, T = {T}: GPU value (relative difference to the published value)")
print("dim  " + "  ".join(f"model {m:<19d}" for m in range(5)))
worst = 0.0
worst_max = 0.0
max_rows = []
t_total = 0.0
for dim in range(1, 13):
    widths = [jne.num_eigs(m, dim) for m in range(5)]
    out = torch.empty((n, sum(widths)), dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    eng.eigs_batch_multi_device(range(5), dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream)
    e1.record(st); torch.cuda.synchronize(); eng.check_async()
    t_total += e0.elapsed_time(e1) * 1e-3
    cells, mcells, off = [], [], 0
    for m in range(5):
        mx = out[:, off]                                  # rows are descending: the first entry is the maximum
        qm = float(torch.quantile(mx, 0.95))
        relm = qm / MHM95_MAX[m][dim - 1] - 1.0
        worst_max = max(worst_max, abs(relm))
        mcells.append(f"{qm:9.4f} ({100 * relm:+.2f} %)")
        tr = out[:, off:off + widths[m]].sum(dim=1); off += widths[m]
        q = float(torch.quantile(tr[: min(n, 16_000_000)], 0.95)) if n <= 16_000_000 else float(np.quantile(tr.cpu().numpy(), 0.95))
        rel = q / MHM95[m][dim - 1] - 1.0
        worst = max(worst, abs(rel))
        cells.append(f"{q:9.4f} ({100 * rel:+.2f} %)")
    print(f"{dim:3d}  " + "  ".join(f"{c:25s}" for c in cells), flush=True)
    max_rows.append(f"{dim:3d}  " + "  ".join(f"{c:25s}" for c in mcells))
print(f"worst relative difference {100 * worst:.2f} %; GPU time for the 60 cells {t_total:.2f} s")
print()
print("maximum-eigenvalue statistic, 95 % quantile, same runs")
print("dim  " + "  ".join(f"model {m:<19d}" for m in range(5)))
print("\n".join(max_rows))
print(f"worst relative difference {100 * worst_max:.2f} %")

import os, sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
st = torch.cuda.current_stream()
n = 1 << 18
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
for aux in ("0", "1"):
    os.environ["JNE_AUX"] = aux
    eng = jne.Engine([0])
    res = []
    for dim, T in ((5, 5000), (6, 10000), (5, 10000)):
        for label, models in (("m4", [4]), ("multi", [0, 1, 2, 3, 4])):
            out = torch.empty((n, 40), dtype=torch.float64, device="cuda")
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.eigs_batch_multi_device(models, dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
                torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
            res.append(f"d{dim} T{T} {label} {n*len(models)/best/1e3:.2f}M runs/s")
    eng.check_async(); eng.close()
    print(f"JNE_AUX={aux}: " + " | ".join(res), flush=True)

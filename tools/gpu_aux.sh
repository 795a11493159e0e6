#!/bin/bash
mkdir -p gpurun_out
for aux in 0 1; do
JNE_AUX=$aux python - <<PY 2>&1 | tee -a gpurun_out/exp_aux.txt
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); n = 133200
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
res = []
for label, models in (("m0", [0]), ("m2", [2]), ("m4", [4]), ("multi", [0, 1, 2, 3, 4])):
    out = torch.empty((n, 62), dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eigs_batch_multi_device(models, 12, 10000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    res.append("%s %.3fM paths/s" % (label, n / best / 1e3))
    if label == "multi": torch.save(out[:2000].cpu(), "gpurun_out/multi_aux$aux.pt")
eng.check_async()
print("JNE_AUX=$aux:", " | ".join(res))
PY
done
python - <<'PY' 2>&1 | tee -a gpurun_out/exp_aux.txt
import torch
a = torch.load("gpurun_out/multi_aux0.pt"); b = torch.load("gpurun_out/multi_aux1.pt")
rel = ((a - b).abs() / a.abs().clamp_min(1e-300))
big = a.abs() > 1e-6
print("aux vs scalar sums: max rel diff on eigenvalues > 1e-6:", rel[big].max().item())
PY
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt

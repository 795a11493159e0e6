#!/bin/bash
mkdir -p gpurun_out
python tools/exp_waves.py 2>&1 | tee gpurun_out/exp_waves.txt
echo "--- NOEPI variant"
JNE_LIBRARY=$PWD/johansen_null_eigenspectra_b200/libjne_exp_NOEPI.so python tools/exp_waves.py 2>&1 | tee gpurun_out/exp_waves_noepi.txt
echo "--- huge skew sanity"
python - <<'PY' 2>&1 | tee gpurun_out/exp_skew_sanity.txt
import os, sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
st = torch.cuda.current_stream()
n = 133200
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
for skew in (0, 4000000):
    os.environ["JNE_SKEW_CYCLES"] = str(skew)
    eng = jne.Engine([0])
    out = torch.empty((n, 62), dtype=torch.float64, device="cuda")
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eigs_batch_multi_device([0,1,2,3,4], 12, 10000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
        torch.cuda.synchronize()
    print("skew", skew, "ms", e0.elapsed_time(e1))
    eng.close()
PY

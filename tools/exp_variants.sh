#!/bin/bash
# builds experiment variants of the library (HERE, before gpurun) or times them (on the GPU box)
if [ "$1" == "build" ]; then
  for v in NORNG NOMMA NOBM; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Iinclude -DJNE_EXP_$v \
      -o johansen_null_eigenspectra_b200/libjne_exp_$v.so johansen_null_eigenspectra_b200/csrc/jne_api.cu &
  done
  wait; ls -la johansen_null_eigenspectra_b200/*.so
else
  for v in "" _exp_NORNG _exp_NOMMA _exp_NOBM; do
    JNE_LIBRARY=$PWD/johansen_null_eigenspectra_b200/libjne$v.so python - <<PY
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); n = 1 << 17
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
for model in (0, 4):
    out = torch.empty((n, 12), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eigs_batch_device(model, 12, 10000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("variant '$v' model", model, "runs/s %.3fM" % (n / ms / 1e3), "cycles/iter/SMSP %.0f" % (ms * 1e-3 * 1.965e9 * 592 / (n * 625.0)))
PY
  done
fi

#!/bin/bash
for v in "$@"; do
  [ "$v" == "base" ] && lib=libjne.so || lib=libjne_exp_$v.so
  JNE_KERNEL=ws JNE_LIBRARY=$PWD/johansen_null_eigenspectra_b200/$lib timeout 300 python - <<PY
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); n = 133200
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
res = []
for label, models in (("m0", [0]), ("m4", [4]), ("multi", [0, 1, 2, 3, 4])):
    out = torch.empty((n, 62), dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eigs_batch_multi_device(models, 12, 10000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    res.append("%s %.3fM seeds/s" % (label, n / best / 1e3))
pass
print("ws variant $v:", " | ".join(res))
PY
done

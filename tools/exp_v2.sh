#!/bin/bash
# v2 moments-kernel geometry sweep: build variants HERE (./tools/exp_v2.sh build), time them on the GPU box (run)
VARS="4_2 2_5 4_3 2_6 1_10 3_3 2_7"
if [ "$1" == "build" ]; then
  for v in $VARS; do
    w=${v%_*}; b=${v#*_}
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC,-pthread -Iinclude -DJNE_V2_WARPS=$w -DJNE_V2_MINB=$b \
      -o johansen_null_eigenspectra_b200/libjne_exp_$v.so johansen_null_eigenspectra_b200/csrc/jne_api.cu johansen_null_eigenspectra_b200/csrc/jne_dat.cpp johansen_null_eigenspectra_b200/csrc/jne_host.cpp &
  done
  wait
  for v in $VARS; do cuobjdump -res-usage johansen_null_eigenspectra_b200/libjne_exp_$v.so | grep -A1 "jne_moments12_kernelILi[02]ELb1" | grep REG | sed -E "s/.*REG:([0-9]+) STACK:([0-9]+).*/$v REG=\1 STACK=\2/"; done
else
  for v in $VARS; do
    JNE_LIBRARY=$PWD/johansen_null_eigenspectra_b200/libjne_exp_$v.so python - <<PY
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); n = 1 << 17
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
res = []
for label, models in (("m0", [0]), ("m4", [4]), ("multi", [0, 1, 2, 3, 4])):
    out = torch.empty((n, 62), dtype=torch.float64, device="cuda")
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eigs_batch_multi_device(models, 12, 10000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res.append("%s %.3fM paths/s" % (label, n / ms / 1e3))
print("variant $v:", " | ".join(res))
PY
  done
fi

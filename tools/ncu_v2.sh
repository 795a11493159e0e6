#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/ncu_v2.py <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); seeds = np.arange(1, 37889, dtype=np.uint32)   # 148 SMs x 2 CTAs x 32 runs x 4 waves
eng.eigs_batch(0, 12, 10000, seeds[:592])
out = eng.eigs_batch(0, 12, 10000, seeds); print(out.shape)
out = eng.eigs_batch_multi(range(5), 12, 10000, seeds); print(out[0].shape)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"jne_moments12|jne_solve" -s 2 -c 4 -f -o gpurun_out/prof_v2 python /tmp/ncu_v2.py > gpurun_out/ncu_v2.log 2>&1
tail -2 gpurun_out/ncu_v2.log

#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/ncu_multi.py <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); seeds = np.arange(1, 29601, dtype=np.uint32)
eng.eigs_batch_multi(range(5), 12, 10000, seeds[:592])
out = eng.eigs_batch_multi(range(5), 12, 10000, seeds); print(out[0].shape)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jne_run_kernel -s 1 -c 1 -f -o gpurun_out/prof_multi_aux python /tmp/ncu_multi.py > gpurun_out/ncu_multi_aux.log 2>&1
tail -2 gpurun_out/ncu_multi_aux.log
ls -la gpurun_out/*.ncu-rep

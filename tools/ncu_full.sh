#!/bin/bash
# ncu --set full capture of the fused kernel (models 0 and 4) at dim 12, T 10 000.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jne_run_kernel -s 2 -c 2 -f -o gpurun_out/prof_run_kernel \
    python tools/ncu_target.py 23680 0,4 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out

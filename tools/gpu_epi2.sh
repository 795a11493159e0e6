#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jne_run_kernel -s 1 -c 1 -f -o gpurun_out/prof_epi_multi2 python tools/ncu_target_epi_multi.py > gpurun_out/ncu_epi_multi2.log 2>&1
tail -2 gpurun_out/ncu_epi_multi2.log

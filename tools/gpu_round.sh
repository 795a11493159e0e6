#!/bin/bash
# One GPU visit: tests, bench, ncu launch list, ncu full capture of the fused kernel.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --runs 32768 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jne_run_kernel -s 5 -c 2 -f -o gpurun_out/prof_run_kernel \
    python tools/ncu_target.py 23680 0,4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

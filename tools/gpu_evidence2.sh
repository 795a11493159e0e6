#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_target.py 2>&1 | tail -6 > gpurun_out/sanitizer_$tool.txt
  tail -3 gpurun_out/sanitizer_$tool.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dat --runs 32560 > gpurun_out/bench_under_ncu.log 2>&1
tail -4 gpurun_out/launches.csv | cut -c1-220

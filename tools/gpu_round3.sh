#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/bench.txt | cut -c1-200
timeout 900 python tools/bench_configs.py 2>&1 | tee gpurun_out/bench_configs.txt | cut -c1-220

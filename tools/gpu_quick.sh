#!/bin/bash
# quick GPU visit: parity tests + bench (no CPU baseline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tee gpurun_out/bench.txt

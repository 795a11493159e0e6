#!/bin/bash
# One GPU visit: tests, bench (both arms), sanitizers, ncu launch list, ncu full capture of the fused kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench.txt | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref.txt | cut -c1-400
./tools/gpu_evidence.sh
python tools/exp_epi_multi.py 2>&1 | tee gpurun_out/exp_epi_multi_new.txt
ls -la gpurun_out | tail -12

#!/bin/bash
# Round-end visit: build check, smoke, tests, both bench arms.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep "^{" | tee gpurun_out/bench_ref.txt | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/bench.txt | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_target.py 2>&1 | tail -6 > gpurun_out/sanitizer_$tool.txt
  tail -3 gpurun_out/sanitizer_$tool.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --runs 32768 > gpurun_out/bench_under_ncu.log 2>&1
cat > /tmp/ncu_multi.py <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); seeds = np.arange(1, 23681, dtype=np.uint32)
eng.eigs_batch_multi(range(5), 12, 10000, seeds[:592])
out = eng.eigs_batch_multi(range(5), 12, 10000, seeds); print(out[0].shape)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jne_run_kernel -s 1 -c 1 -f -o gpurun_out/prof_multi python /tmp/ncu_multi.py > gpurun_out/ncu_multi.log 2>&1
tail -2 gpurun_out/ncu_multi.log

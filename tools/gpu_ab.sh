#!/bin/bash
# A/B of the two kernel families + parity tests on the default one
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
for kf in v2 v1; do
  JNE_KERNEL=$kf timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/bench_$kf.txt
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$kf.txt").read())
print("$kf: fused %.3fM runs/s  e2e %.3fM  per-model %.3fM (%.3f of peak) %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["per_model_path"]["value"]/1e6, d["per_model_path"]["frac_of_fp64_peak"], d["per_model_path"]["achieved_tflops_per_model"]))
PY
done

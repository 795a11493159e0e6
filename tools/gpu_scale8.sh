#!/bin/bash
# 8-GPU visit: the scaling bench (one process per GPU under torchrun) and the in-process multi-device tests
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/scale8_gpus.txt
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/bench_${N}gpu.txt | cut -c1-300
timeout 900 python -m pytest tests -x -q -m gpu -k "multi or device" 2>&1 | tail -4 | tee gpurun_out/pytest_${N}gpu.txt
timeout 300 python - <<PY 2>&1 | tee gpurun_out/inproc_${N}gpu.txt
import sys, time, numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
n_dev = jne.device_count() if hasattr(jne, "device_count") else $N
seeds = np.arange(1, 1 + 133200 * n_dev, dtype=np.uint32)
one = jne.Engine([0]); ref = one.eigs_batch_multi(range(5), 12, 10000, seeds[:8192])
eng = jne.Engine(list(range(n_dev)))
eng.eigs_batch_multi(range(5), 12, 10000, seeds[: 4096 * n_dev])
t0 = time.time(); out = eng.eigs_batch_multi(range(5), 12, 10000, seeds); dt = time.time() - t0
same = all(np.array_equal(out[m][:8192], ref[m]) for m in range(5))
print(f"in-process context over {n_dev} devices: {5 * seeds.size / dt / 1e6:.2f} M runs/s end to end (host buffers), first 8192 seeds bit-identical to the 1-device context: {same}")
PY

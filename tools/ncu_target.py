"""Short single-GPU target for ncu captures: one warm-up and one measured launch of the hot path through the
device-pointer entry (one kernel launch per call, no chunking), at any configuration.
  python tools/ncu_target.py --dim 12 --T 10000 --n 133200 --models 0,1,2,3,4
`ncu -s 1 -c 1 -k regex:<kernel>` then captures the measured launch of that kernel."""
import argparse
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=12)
ap.add_argument("--T", type=int, default=10000)
ap.add_argument("--n", type=int, default=133200)
ap.add_argument("--models", default="0,1,2,3,4")
a = ap.parse_args()
models = [int(x) for x in a.models.split(",")]
eng = jne.Engine([0])
st = torch.cuda.current_stream()
seeds = torch.arange(1, a.n + 1, dtype=torch.int32, device="cuda")
out = torch.empty((a.n, sum(jne.num_eigs(m, a.dim) for m in models)), dtype=torch.float64, device="cuda")
for rep in range(2):                 # launch 0 = warm-up (skip with ncu -s 1), launch 1 = measured
    if len(models) == 1:
        eng.eigs_batch_device(models[0], a.dim, a.T, seeds.data_ptr(), a.n, out.data_ptr(), st.cuda_stream)
    else:
        eng.eigs_batch_multi_device(models, a.dim, a.T, seeds.data_ptr(), a.n, out.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
eng.check_async()
print("dim", a.dim, "T", a.T, "n", a.n, "models", models, "mean trace of the first block", float(out[:, :jne.num_eigs(models[0], a.dim)].sum(1).mean()))

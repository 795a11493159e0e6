#!/bin/bash
# warp-specialised kernel family: bitwise regression against the default family, then timing of both
mkdir -p gpurun_out
timeout 300 python tools/dump_eigs.py /tmp/v1.npz 2>&1 | tail -1
JNE_KERNEL=ws timeout 300 python tools/dump_eigs.py /tmp/ws.npz 2>&1 | tail -1
python tools/cmp_dumps.py /tmp/v1.npz /tmp/ws.npz 2>&1 | tail -5
for kf in v1 ws; do
JNE_KERNEL=$kf timeout 300 python - <<PY
import sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); n = 133200
seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
res = []
for label, models, dim, T in (("m0", [0], 12, 10000), ("m4", [4], 12, 10000), ("multi", [0, 1, 2, 3, 4], 12, 10000), ("multi d5 T5000", [0, 1, 2, 3, 4], 5, 5000), ("multi d8", [0, 1, 2, 3, 4], 8, 10000)):
    out = torch.empty((n, 65), dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eigs_batch_multi_device(models, dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    res.append("%s %.3fM seeds/s" % (label, n / best / 1e3))
eng.check_async()
print("family $kf:", " | ".join(res))
PY
done

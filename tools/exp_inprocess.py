"""One in-process context over all visible GPUs (what the Rust shim's Gpu::new() creates): end-to-end rate of
jne_eigs_batch_multi with host buffers, fresh vs reused output array, and bit-identity with the 1-device context."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
import torch
n_dev = torch.cuda.device_count()
per = 133200 * 4
seeds = np.arange(1, 1 + per * n_dev, dtype=np.uint32)
one = jne.Engine([0]); ref = one.eigs_batch_multi(range(5), 12, 10000, seeds[:8192]); one.close()
eng = jne.Engine(list(range(n_dev)))
eng.eigs_batch_multi(range(5), 12, 10000, seeds[: 8192 * n_dev])
t0 = time.time(); out = eng.eigs_batch_multi(range(5), 12, 10000, seeds); dt = time.time() - t0
same = all(np.array_equal(out[m][:8192], ref[m]) for m in range(5))
print(f"in-process context over {n_dev} devices, {seeds.size} seeds x 5 models: {5 * seeds.size / dt / 1e6:.2f} M runs/s end to end into a fresh array (first-touch page faults inside), bit-identical to the 1-device context: {same}")
buf = np.zeros((seeds.size, 62))
best = 1e9
for rep in range(3):
    t0 = time.time(); eng.eigs_batch_multi(range(5), 12, 10000, seeds, out=buf); best = min(best, time.time() - t0)
print(f"same call into a reused (already touched) array: {5 * seeds.size / best / 1e6:.2f} M runs/s")

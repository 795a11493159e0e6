"""Phase-stagger experiment: fused five-model kernel (and models 0 / 4 alone) at dim 12, T 10000 for several
JNE_SKEW_CYCLES (start delay per resident-CTA slot of the first wave).  Also checks that records do not change."""
import os, sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
st = torch.cuda.current_stream()
ref = None
for n in (133200, 532800):
    seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
    for skew in (0, 40000, 80000, 120000, 160000, 240000, 390000):
        os.environ["JNE_SKEW_CYCLES"] = str(skew)
        eng = jne.Engine([0])
        res = []
        for label, models in (("m0", [0]), ("m4", [4]), ("multi", [0, 1, 2, 3, 4])):
            if n > 133200 and label != "multi":
                continue
            out = torch.empty((n, 62), dtype=torch.float64, device="cuda")
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.eigs_batch_multi_device(models, 12, 10000, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream); e1.record()
                torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
            res.append("%s %.3fM seeds/s" % (label, n / best / 1e3))
            if label == "multi" and n == 133200:
                if ref is None: ref = out.clone()
                else: assert torch.equal(ref, out), "records changed under stagger"
        eng.check_async(); eng.close()
        print(f"n {n} skew {skew}: " + " | ".join(res), flush=True)

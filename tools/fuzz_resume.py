"""Randomised resume sweep for run_models_simulation: random model masks, random pre-existing files (random subsets of
seeds in random order, finished or unfinished, sometimes with a foreign header), then the fused job; every selected file
must end with each seed of 1..num_runs exactly once and the records jne_eigs_batch gives."""
import os, shutil, sys, tempfile
import numpy as np
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
from johansen_null_eigenspectra_b200 import dat
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 40
eng = jne.Engine([0])
bad = 0
for t in range(trials):
    dim, T, n = int(rng.integers(1, 6)), int(rng.integers(40, 90)), int(rng.integers(1, 3000))
    models = [m for m in range(5) if rng.random() < 0.7] or [2]
    d = tempfile.mkdtemp(prefix="jne_fuzz_")
    try:
        names = {m: os.path.join(d, f"m{m}.dat") for m in range(5)}
        truth = {m: eng.eigs_batch(m, dim, T, np.arange(1, n + 1, dtype=np.uint32)) for m in range(5)}
        expect_before = {}
        open_complete = set()      # complete but never finished: left as it is (the reference returns early as well)
        untouched = {}
        for m in range(5):
            kind = rng.choice(["absent", "partial", "partial_open", "complete", "foreign"])
            if kind == "absent":
                expect_before[m] = 0; continue
            if kind == "foreign":
                w = dat.AppendOnlyWriter(names[m], m, dim, T + 1); w.append_batch([1], truth[m][:1]); w.finish()
                expect_before[m] = 0; continue
            k = n if kind == "complete" else int(rng.integers(0, n + 1))
            have = rng.permutation(np.arange(1, n + 1))[:k].astype(np.uint32)
            w = dat.AppendOnlyWriter(names[m], m, dim, T)
            if k: w.append_batch(have, truth[m][have - 1])
            (w.finish if kind != "partial_open" else w.abandon)()
            expect_before[m] = k
            if kind == "partial_open" and k == n: open_complete.add(m)
        for m in range(5):
            if m not in models and os.path.exists(names[m]): untouched[m] = open(names[m], "rb").read()
        st = dat.run_models_simulation(models, dim, T, n, names, devices=None, engine=eng)
        for m in models:
            seeds, eigs, mm, dd, tt = dat.read_append_file(names[m])
            ok = (mm, dd, tt) == (m, dim, T) and len(seeds) == n and np.array_equal(np.sort(seeds), np.arange(1, n + 1))
            ok = ok and np.array_equal(eigs[np.argsort(seeds)], truth[m]) and dat.file_info(names[m])["has_trailer"] == ((st[m]["computed"] > 0 or expect_before[m] == n) and m not in open_complete)
            ok = ok and st[m]["completed_before"] == expect_before[m] and st[m]["computed"] == n - expect_before[m] and st[m]["total_in_file"] == n
            if not ok:
                bad += 1; print("FAIL trial", t, "model", m, dim, T, n, st[m], expect_before[m], flush=True)
        for m, b in untouched.items():
            if open(names[m], "rb").read() != b:
                bad += 1; print("FAIL untouched file changed", t, m, flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
print(f"{trials} random resume scenarios (dim 1..5, 1..3000 runs, random masks, five file states): failures {bad}")

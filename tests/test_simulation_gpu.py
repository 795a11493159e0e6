"""End-to-end orchestration on the GPU: run_model_simulation (C++ csrc/jne_host.cpp) -> batched EIGENVALS_V6 file,
mirroring the reference's integration tests.  (-m gpu)"""
import numpy as np
import pytest

from johansen_null_eigenspectra_b200 import dat

pytestmark = pytest.mark.gpu


def test_resumable_simulation(tmp_path, engine):
    """src/tests/data_storage/integration/resumable.rs:4-76: model 0, dim 2, T 103, 5 runs; drop seeds 2 and 4 from
    the file; re-run; every seed 1..5 present and each record identical (the reference asks 1e-10 on the sum)."""
    path = tmp_path / "eigenvalues_model0_dim2_steps103.dat"
    st = dat.run_model_simulation(0, 2, 103, 5, str(path), devices=[0])
    assert st == {"completed_before": 0, "computed": 5, "total_in_file": 5}
    seeds, eigs, m, d, t = dat.read_append_file(path)
    assert (m, d, t) == (0, 2, 103) and sorted(seeds) == [1, 2, 3, 4, 5] and eigs.shape == (5, 2)   # basic_api.rs:35
    first = {int(s): e.copy() for s, e in zip(seeds, eigs)}
    # integration/helpers.rs:25-68: rewrite the file without seeds 2 and 4
    keep = np.isin(seeds, [1, 3, 5])
    path.unlink()
    w = dat.AppendOnlyWriter(path, 0, 2, 103)
    w.append_batch(seeds[keep], eigs[keep])
    w.finish()
    st = dat.run_model_simulation(0, 2, 103, 5, str(path), devices=[0])
    assert st == {"completed_before": 3, "computed": 2, "total_in_file": 5}
    seeds2, eigs2, *_ = dat.read_append_file(path)
    assert sorted(seeds2) == [1, 2, 3, 4, 5]
    for s, e in zip(seeds2, eigs2):
        assert np.array_equal(e, first[int(s)])                 # bit-identical, stronger than 1e-10
        assert np.array_equal(e, engine.eigs_batch(0, 2, 103, [int(s)])[0])
    # a third run has nothing to do (parallel_compute.rs:182-198)
    assert dat.run_model_simulation(0, 2, 103, 5, str(path), devices=[0])["computed"] == 0


def test_all_models_files_and_mismatch_restart(tmp_path, engine):
    """integration/multiple_models.rs:4-45 (all five models, finite values, seeds in 1..=num_runs) and the
    delete-and-restart policy on a parameter mismatch (parallel_compute.rs:159-175)."""
    for model in range(5):
        path = tmp_path / f"eigenvalues_model{model}_dim2_steps211.dat"
        st = dat.run_model_simulation(model, 2, 211, 300, str(path), devices=[0])
        assert st["total_in_file"] == 300
        seeds, eigs, m, d, t = dat.read_append_file(path)
        assert m == model and np.all(np.isfinite(eigs)) and eigs.shape[1] == (3 if model in (1, 3) else 2)
        assert sorted(seeds) == list(range(1, 301))
        assert np.array_equal(eigs[np.argsort(seeds)], engine.eigs_batch(model, 2, 211, np.arange(1, 301)))
    # same file name, different steps -> incompatible header -> removed and recomputed
    path = tmp_path / "eigenvalues_model0_dim2_steps211.dat"
    st = dat.run_model_simulation(0, 2, 212, 10, str(path), devices=[0])
    assert st == {"completed_before": 0, "computed": 10, "total_in_file": 10}
    assert dat.file_info(path)["steps"] == 212


def test_fused_models_simulation(tmp_path, engine):
    """run_models_simulation (the CLI's model loop of src/main.rs:109-114 as one fused pass): every file is
    byte-identical to the one run_model_simulation writes for that model; files with different progress are resumed
    independently (seeds grouped by the set of models that lack them); complete files are left alone; a header
    mismatch restarts that file only."""
    dim, T, n = 3, 150, 700
    ref_dir, out_dir = tmp_path / "ref", tmp_path / "out"
    ref_dir.mkdir(); out_dir.mkdir()
    names = {m: str(out_dir / f"eigenvalues_model{m}_dim{dim}_steps{T}.dat") for m in range(5)}
    for m in range(5):
        dat.run_model_simulation(m, dim, T, n, str(ref_dir / f"m{m}.dat"), devices=[0])
    # fresh job, all five models
    st = dat.run_models_simulation(range(5), dim, T, n, names, devices=[0])
    for m in range(5):
        assert st[m] == {"completed_before": 0, "computed": n, "total_in_file": n}
        assert open(names[m], "rb").read() == open(ref_dir / f"m{m}.dat", "rb").read()
    # nothing left to do
    st = dat.run_models_simulation(range(5), dim, T, n, names, devices=[0])
    assert all(st[m]["computed"] == 0 and st[m]["total_in_file"] == n for m in range(5))
    # uneven progress: model 2 has seeds 1..300 already, model 4 has an incompatible header, model 1 is complete
    import os
    os.remove(names[0]); os.remove(names[2]); os.remove(names[3]); os.remove(names[4])
    dat.run_model_simulation(2, dim, T, 300, names[2], devices=[0])
    dat.run_model_simulation(4, dim, T + 1, 50, names[4], devices=[0])
    before1 = open(names[1], "rb").read()
    st = dat.run_models_simulation([0, 1, 2, 4], dim, T, n, names, devices=[0])
    assert st[0]["computed"] == n and st[1]["computed"] == 0 and st[4] == {"completed_before": 0, "computed": n, "total_in_file": n}
    assert st[2] == {"completed_before": 300, "computed": n - 300, "total_in_file": n}
    assert open(names[1], "rb").read() == before1 and not os.path.exists(names[3])
    for m in (0, 2, 4):
        seeds, eigs, mm, d, t = dat.read_append_file(names[m])
        assert (mm, d, t) == (m, dim, T) and sorted(seeds) == list(range(1, n + 1))
        assert np.array_equal(eigs[np.argsort(seeds)], engine.eigs_batch(m, dim, T, np.arange(1, n + 1)))
        assert dat.file_info(names[m])["has_trailer"]


def test_streaming_percentiles(engine):
    """Row f3: on-device trace / max-eig percentiles == the reference's analyser (src/simulation_analyzers.rs:4-65,
    restated in oracle.percentiles) applied to the same records on the host."""
    import torch
    from oracle import johansen_oracle as orc
    qs = [0.0, 0.5, 0.9, 0.95, 0.975, 0.99, 1.0]
    for model, dim, T, n in [(0, 2, 103, 1), (1, 3, 200, 10001), (4, 12, 400, 30000)]:
        ev = engine.eigs_batch(model, dim, T, np.arange(1, n + 1, dtype=np.uint32))
        tr, mx = engine.simulate_percentiles(model, dim, T, n, qs)
        trace = ev[:, 0].copy()
        for k in range(1, ev.shape[1]):          # eigenvalues.iter().sum(): left to right in stored order
            trace += ev[:, k]
        ref_tr = orc.percentiles(trace, qs)
        assert np.array_equal(tr, ref_tr), (tr - ref_tr)
        assert np.array_equal(mx, orc.percentiles(ev[:, 0], qs))
        d = torch.from_numpy(ev).cuda()
        tr2, mx2 = engine.percentiles_device(d.data_ptr(), n, ev.shape[1], ev.shape[1], qs, torch.cuda.current_stream().cuda_stream)
        assert np.array_equal(tr2, tr) and np.array_equal(mx2, mx)
    # every model of a fused pass: identical to the per-model calls (rows f2 x f3)
    qs2 = [0.5, 0.9, 0.95, 0.99]
    both = engine.simulate_percentiles_multi(range(5), 5, 120, 20011, qs2)
    for m in range(5):
        tr1, mx1 = engine.simulate_percentiles(m, 5, 120, 20011, qs2)
        assert np.array_equal(both[m][0], tr1) and np.array_equal(both[m][1], mx1)
    part = engine.simulate_percentiles_multi([1, 4], 12, 64, 3000, qs2, first_seed=7)
    for m in (1, 4):
        tr1, mx1 = engine.simulate_percentiles(m, 12, 64, 3000, qs2, first_seed=7)
        assert np.array_equal(part[m][0], tr1) and np.array_equal(part[m][1], mx1)
    # a different first seed selects a different sample (seeds 101..200 == the tail of 1..200)
    ev = engine.eigs_batch(2, 4, 150, np.arange(101, 201, dtype=np.uint32))
    tr, mx = engine.simulate_percentiles(2, 4, 150, 100, [0.5], first_seed=101)
    trace = ev[:, 0].copy()
    for k in range(1, ev.shape[1]):
        trace += ev[:, k]
    assert np.array_equal(tr, orc.percentiles(trace, [0.5]))


def test_streaming_sink_delivers_the_batch(engine):
    """jne_eigs_batch_multi_stream: the reference's channel model (rows SENT as they finish,
    src/data_storage/parallel_compute.rs:14-41).  Every seed's row arrives exactly once, bit-identical to the batch
    call, whatever the chunking; a failing sink aborts the call with an error and leaves the context usable."""
    import johansen_null_eigenspectra_b200 as jne
    for dim, T, n in [(12, 64, 20011), (5, 50, 300001), (9, 40, 12345), (2, 30, 7)]:
        seeds = np.arange(5, 5 + n, dtype=np.uint32)
        ref = engine.eigs_batch_multi(range(5), dim, T, seeds)
        width = sum(v.shape[1] for v in ref.values())
        got = np.full((n, width), np.nan)
        hits = np.zeros(n, dtype=np.int32)

        def sink(first, rows):
            got[first:first + rows.shape[0]] = rows
            hits[first:first + rows.shape[0]] += 1

        engine.eigs_batch_multi_stream(range(5), dim, T, seeds, sink)
        assert np.all(hits == 1)
        off = 0
        for m in range(5):
            assert np.array_equal(got[:, off:off + ref[m].shape[1]], ref[m]), (dim, m)
            off += ref[m].shape[1]
    calls = []
    with pytest.raises(jne.JneError) as e:
        engine.eigs_batch_multi_stream([0, 3], 12, 64, np.arange(1, 50001, dtype=np.uint32), lambda first, rows: calls.append(first) or True)
    assert e.value.status == -5 and len(calls) >= 1
    with pytest.raises(ZeroDivisionError):
        engine.eigs_batch_multi_stream([0], 3, 20, np.arange(1, 100, dtype=np.uint32), lambda first, rows: 1 / 0)
    assert engine.eigs_batch(0, 3, 20, np.arange(1, 100, dtype=np.uint32)).shape == (99, 3)

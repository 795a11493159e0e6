"""The in-process multi-device context (-m gpu; needs >= 2 GPUs, skipped otherwise): what the reference-side FFI shim
creates with jne_init(NULL, 0) (ffi/rust/src/gpu_ffi.rs `Gpu::new`) and what replaces the rayon pool of
src/data_storage/parallel_compute.rs:14-41.  Seeds are sharded contiguously over the context's devices with no
collective; a record must not depend on the device that computed it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    import johansen_null_eigenspectra_b200 as jne
    if jne.lib.jne_device_count() < 2:
        pytest.skip("needs at least two GPUs")
    one, all_ = jne.Engine([0]), jne.Engine(None)
    yield one, all_
    one.close(); all_.close()


def test_all_devices_bit_identical_to_one(engines):
    import johansen_null_eigenspectra_b200 as jne
    one, all_ = engines
    nd = jne.lib.jne_ctx_device_count(all_._ctx)
    assert nd == jne.lib.jne_device_count() >= 2
    for dim, T, n in [(12, 400, 20011), (5, 300, 150001), (2, 64, 7), (9, 103, 1)]:      # ragged shares, fewer seeds than devices
        seeds = np.arange(3, 3 + n, dtype=np.uint32)
        a = one.eigs_batch_multi(range(5), dim, T, seeds)
        b = all_.eigs_batch_multi(range(5), dim, T, seeds)
        for m in range(5):
            assert np.array_equal(a[m], b[m]), (dim, T, n, m)
        assert np.array_equal(one.eigs_batch(3, dim, T, seeds), all_.eigs_batch(3, dim, T, seeds))
    out = np.empty((5000, 13))
    all_.wait(all_.submit(1, 12, 100, np.arange(1, 5001, dtype=np.uint32), out))           # async pair on every device
    assert np.array_equal(out, one.eigs_batch(1, 12, 100, np.arange(1, 5001, dtype=np.uint32)))


def test_callers_current_device_is_left_alone():
    """An FFI library must hand the calling thread back with the CUDA device it came with: a context over several
    devices switches devices internally (jne_init, every batched call)."""
    import torch
    import johansen_null_eigenspectra_b200 as jne
    if jne.lib.jne_device_count() < 2:
        pytest.skip("needs at least two GPUs")
    for cur in (0, 1):
        torch.cuda.set_device(cur)
        eng = jne.Engine(None)
        assert torch.cuda.current_device() == cur
        eng.eigs_batch_multi(range(5), 3, 50, np.arange(1, 4001, dtype=np.uint32))
        eng.simulate_percentiles(2, 3, 50, 4000, [0.5])
        assert torch.cuda.current_device() == cur
        assert torch.zeros(1, device="cuda").device.index == cur
        eng.close()
        assert torch.cuda.current_device() == cur
    torch.cuda.set_device(0)


def test_streaming_sink_over_all_devices(engines):
    """The row sink is called from one host thread per device, concurrently, for disjoint ranges: every row once,
    bit-identical to the one-device batch; and the fused .dat job on all devices writes the files of the one-device job."""
    import os, tempfile
    from johansen_null_eigenspectra_b200 import dat
    one, all_ = engines
    n, dim, T = 100003, 12, 48
    seeds = np.arange(1, n + 1, dtype=np.uint32)
    ref = one.eigs_batch_multi(range(5), dim, T, seeds)
    got = np.full((n, 62), np.nan)
    hits = np.zeros(n, dtype=np.int32)

    def sink(first, rows):
        got[first:first + rows.shape[0]] = rows
        hits[first:first + rows.shape[0]] += 1

    all_.eigs_batch_multi_stream(range(5), dim, T, seeds, sink)
    assert np.all(hits == 1)
    assert np.array_equal(got, np.concatenate([ref[m] for m in range(5)], axis=1))
    with tempfile.TemporaryDirectory() as d:
        a = {m: os.path.join(d, f"a{m}.dat") for m in range(5)}
        b = {m: os.path.join(d, f"b{m}.dat") for m in range(5)}
        dat.run_models_simulation(range(5), 5, 40, 200003, a, engine=one)
        dat.run_models_simulation(range(5), 5, 40, 200003, b, engine=all_)
        for m in range(5):
            assert open(a[m], "rb").read() == open(b[m], "rb").read(), m


def test_error_from_any_device_surfaces(engines):
    import johansen_null_eigenspectra_b200 as jne
    _, all_ = engines
    seeds = np.arange(1, 4001, dtype=np.uint32)
    with pytest.raises(jne.JneError) as e:          # T < p: singular S2 on every device -> NaN (the reference panics at :45)
        all_.eigs_batch(0, 8, 5, seeds)
    assert e.value.status == -3 and "non-finite" in str(e.value)
    assert all_.eigs_batch(0, 8, 100, seeds).shape == (4000, 8)      # the context stays usable, no stale rows
    assert np.array_equal(all_.eigs_batch(0, 8, 100, seeds[:10]), all_.eigs_batch(0, 8, 100, seeds)[:10])


def test_percentiles_over_all_devices(engines):
    """Row f3 on the multi-device context: per-device (trace, max) arrays, exact distributed selection (histograms are
    the only data that leaves a device): bit-equal to the single-device result and to the host analyser."""
    from oracle import johansen_oracle as orc
    one, all_ = engines
    qs = [0.0, 0.5, 0.9, 0.95, 0.99, 1.0]
    for dim, T, n in [(5, 120, 20011), (12, 64, 3001), (3, 50, 3)]:
        a = one.simulate_percentiles_multi(range(5), dim, T, n, qs, first_seed=11)
        b = all_.simulate_percentiles_multi(range(5), dim, T, n, qs, first_seed=11)
        for m in range(5):
            assert np.array_equal(a[m][0], b[m][0]) and np.array_equal(a[m][1], b[m][1]), (dim, T, n, m)
    ev = one.eigs_batch(2, 5, 120, np.arange(11, 11 + 20011, dtype=np.uint32))
    trace = ev[:, 0].copy()
    for k in range(1, ev.shape[1]):
        trace += ev[:, k]
    tr, mx = all_.simulate_percentiles(2, 5, 120, 20011, qs, first_seed=11)
    assert np.array_equal(tr, orc.percentiles(trace, qs)) and np.array_equal(mx, orc.percentiles(ev[:, 0], qs))

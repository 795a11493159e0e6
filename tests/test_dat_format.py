"""CPU tests of the batched EIGENVALS_V6 writer / reader / resume scan
against the byte vectors of the reference's DATA_FORMAT.md and its own storage tests.  (-m "not gpu")"""
import os
import struct

import numpy as np
import pytest

from johansen_null_eigenspectra_b200 import JneError, dat


# src/tests/data_storage/uleb128_unit_test.rs:4-30, uleb128_test.rs
def test_uleb128_vectors():
    enc = {0: [0x00], 1: [0x01], 127: [0x7F], 128: [0x80, 0x01], 255: [0xFF, 0x01], 256: [0x80, 0x02],
           300: [0xAC, 0x02], 16383: [0xFF, 0x7F], 16384: [0x80, 0x80, 0x01]}
    for v, b in enc.items():
        assert dat.uleb128_encode(v) == bytes(b)
        assert dat.uleb128_decode(bytes(b)) == (v, len(b))
    sizes = {0: 1, 127: 1, 128: 2, 16383: 2, 16384: 3, 2097151: 3, 2097152: 4, 268435455: 4, 268435456: 5, 2**32 - 1: 5}
    for v, n in sizes.items():
        assert dat.uleb128_encoded_size(v) == n == len(dat.uleb128_encode(v))
        assert dat.uleb128_decode(dat.uleb128_encode(v)) == (v, n)


def test_uleb128_errors():   # uleb128_unit_test.rs:67-100
    with pytest.raises(ValueError, match="Incomplete"):
        dat.uleb128_decode(bytes([0x80]))
    with pytest.raises(ValueError, match="Incomplete"):
        dat.uleb128_decode(b"")
    assert dat.uleb128_decode(bytes([0xFF, 0xFF, 0xFF, 0xFF, 0x0F])) == (2**32 - 1, 5)
    with pytest.raises(ValueError, match="too large"):
        dat.uleb128_decode(bytes([0xFF, 0xFF, 0xFF, 0xFF, 0x1F]))
    with pytest.raises(ValueError, match="too long"):
        dat.uleb128_decode(bytes([0x80, 0x80, 0x80, 0x80, 0x80, 0x01]))


def test_bytes_match_data_format_md(tmp_path):
    """DATA_FORMAT.md:98-150: header, two records, trailer, byte for byte."""
    path = tmp_path / "eigenvalues_model0_dim1_steps10.dat"
    w = dat.AppendOnlyWriter(path, 0, 1, 10)
    w.append_eigenvalues(1, [1.0])
    w.append_eigenvalues(300, [2.0])
    w.finish()
    raw = path.read_bytes()
    header = bytes.fromhex("45 49 47 45 4E 56 41 4C 53 5F 56 36 00 01 0A 00 00 00")
    rec1 = bytes.fromhex("01 01 3F F0 00 00 00 00 00 00")[:2] + struct.pack("<d", 1.0)
    rec2 = bytes.fromhex("AC 02 01") + struct.pack("<d", 2.0)
    trailer = b"EOF_MARK" + struct.pack("<Q", 2) + bytes([1])
    assert raw == header + rec1 + rec2 + trailer
    assert len(raw) == 18 + 10 + 11 + 17


def test_expected_file_size_matches_written_file(tmp_path):
    n, p = 20000, 3
    path = tmp_path / "f.dat"
    w = dat.AppendOnlyWriter(path, 2, 3, 77)
    w.append_batch(np.arange(1, n + 1), np.random.default_rng(0).random((n, p)))
    w.finish()
    assert os.path.getsize(path) == dat.expected_file_size(n, p)     # file_format.rs:14-27
    # DATA_FORMAT.md:83-96 example (dim 1, 10M records, seeds 0..9 999 999): its 47 886 371 bytes count header,
    # seed, count byte and trailer but leave out the 8-byte eigenvalue of every record (doc drift); with the
    # payload, and seeds 1..=10M as the code numbers them (progress.rs:58), the size differs by 3 bytes
    assert abs(dat.expected_file_size(10_000_000, 1) - (47_886_371 + 80_000_000)) <= 3


def test_round_trip_and_info(tmp_path):   # append_writer_test.rs round-trip / 1000 records
    rng = np.random.default_rng(1)
    seeds = rng.permutation(np.arange(1, 1001)).astype(np.uint32)
    eigs = rng.random((1000, 13))
    path = tmp_path / "rt.dat"
    w = dat.AppendOnlyWriter(path, 3, 12, 10000)
    assert w.existing_records == 0
    w.append_batch(seeds[:400], eigs[:400])
    w.append_batch(seeds[400:], eigs[400:])
    w.finish()
    info = dat.file_info(path)
    assert info == {"model": 3, "dim": 12, "steps": 10000, "records": 1000, "eigenvalues_per_run": 13, "has_trailer": True}
    s, e, m, d, t = dat.read_append_file(path)
    assert (m, d, t) == (3, 12, 10000) and np.array_equal(s, seeds) and np.array_equal(e, eigs)


def test_unfinished_file_is_scan_read_and_torn_record_dropped(tmp_path):   # append_writer_test.rs:66-89, reader.rs:199-210
    path = tmp_path / "unfinished.dat"
    w = dat.AppendOnlyWriter(path, 0, 2, 103)
    w.append_batch([1, 2, 3], np.arange(6.0).reshape(3, 2))
    w.abandon()                                   # no trailer
    info = dat.file_info(path)
    assert info["records"] == 3 and not info["has_trailer"] and info["eigenvalues_per_run"] == 2
    with open(path, "ab") as f:                   # a torn 4th record: seed + count + half an eigenvalue
        f.write(bytes([4, 2]) + b"\x00\x00\x00\x00")
    s, e, *_ = dat.read_append_file(path)
    assert list(s) == [1, 2, 3] and np.array_equal(e, np.arange(6.0).reshape(3, 2))
    # re-opening truncates the torn tail and appends cleanly
    w = dat.AppendOnlyWriter(path, 0, 2, 103)
    assert w.existing_records == 3
    w.append_batch([4], [[7.0, 8.0]])
    w.finish()
    s, e, *_ = dat.read_append_file(path)
    assert list(s) == [1, 2, 3, 4] and list(e[3]) == [7.0, 8.0] and dat.file_info(path)["has_trailer"]


def test_resume_removes_trailer_and_counts(tmp_path):   # writer.rs:79-115,181-203
    path = tmp_path / "resume.dat"
    w = dat.AppendOnlyWriter(path, 1, 2, 50)
    w.append_batch([1, 2], np.ones((2, 3)))
    w.finish()
    w = dat.AppendOnlyWriter(path, 1, 2, 50)
    assert w.existing_records == 2
    w.append_batch([5], np.full((1, 3), 2.0))
    w.finish()
    info = dat.file_info(path)
    assert info["records"] == 3 and info["has_trailer"]
    assert os.path.getsize(path) == 18 + 3 * (1 + 1 + 24) + 17


def test_header_mismatch_and_count_errors(tmp_path):   # append_writer_test.rs:143-227
    path = tmp_path / "mm.dat"
    w = dat.AppendOnlyWriter(path, 0, 2, 50)
    w.append_batch([1], [[1.0, 2.0]])
    with pytest.raises(JneError, match="Eigenvalue count mismatch: expected 2, actual 3"):
        w.append_batch([2], [[1.0, 2.0, 3.0]])
    with pytest.raises(JneError, match="Too many eigenvalues: 256 exceeds maximum of 255"):
        w.append_batch([3], np.zeros((1, 256)))
    w.finish()
    with pytest.raises(JneError, match="Model mismatch: file has model 0, expected 1"):
        dat.AppendOnlyWriter(path, 1, 2, 50)
    with pytest.raises(JneError, match="Dimension mismatch: file has dim 2, expected 3"):
        dat.AppendOnlyWriter(path, 0, 3, 50)
    with pytest.raises(JneError, match="Steps mismatch: file has steps 50, expected 51"):
        dat.AppendOnlyWriter(path, 0, 2, 51)
    bad = tmp_path / "bad.dat"
    bad.write_bytes(b"NOT_A_DAT_FILE_AT_ALL" * 3)
    with pytest.raises(JneError, match="magic header mismatch"):
        dat.file_info(bad)
    w = dat.AppendOnlyWriter(bad, 0, 2, 50)          # foreign magic: recreated (writer.rs:116-150)
    assert w.existing_records == 0
    w.finish()
    assert dat.file_info(bad)["records"] == 0


def test_resume_scan_bitmap(tmp_path):   # progress.rs:11-61, integration/helpers.rs:25-68 (file rewritten without some seeds)
    path = tmp_path / "p.dat"
    assert dat.check_append_progress(path, 0, 2, 103, 5)[0] == 0             # missing file
    assert list(dat.check_append_progress(path, 0, 2, 103, 5)[1]) == [1, 2, 3, 4, 5]
    w = dat.AppendOnlyWriter(path, 0, 2, 103)
    w.append_batch([1, 3, 5, 9], np.zeros((4, 2)))                           # 9 > num_runs: counted, not mapped
    w.finish()
    done, rem = dat.check_append_progress(path, 0, 2, 103, 5)
    assert done == 4 and list(rem) == [2, 4]
    with pytest.raises(JneError, match="Steps mismatch"):
        dat.check_append_progress(path, 0, 2, 104, 5)
    n = 100_000
    big = tmp_path / "big.dat"
    w = dat.AppendOnlyWriter(big, 4, 1, 10)
    keep = np.setdiff1d(np.arange(1, n + 1), np.arange(7, n + 1, 1000))
    w.append_batch(keep, np.zeros((keep.size, 1)))
    w.abandon()
    done, rem = dat.check_append_progress(big, 4, 1, 10, n)
    assert done == keep.size and np.array_equal(rem, np.arange(7, n + 1, 1000))


def test_filename():   # simulation_test.rs:68
    assert dat.get_filename(0, 5, 999) == "data/eigenvalues_model0_dim5_steps999.dat"


def test_strided_append_matches_contiguous(tmp_path):
    """jne_dat_append_batch_strided writes one model's block of fused multi-model rows: same bytes as appending the
    de-interleaved copy; a stride below the eigenvalue count is refused."""
    rng = np.random.default_rng(3)
    n, width = 1000, 12 + 13 + 12
    rows = rng.standard_normal((n, width))
    seeds = rng.permutation(np.arange(1, n + 1)).astype(np.uint32)
    for off, p, model in ((0, 12, 0), (12, 13, 1), (25, 12, 2)):
        a, b = tmp_path / f"a{model}.dat", tmp_path / f"b{model}.dat"
        w = dat.AppendOnlyWriter(a, model, 12, 50); w.append_batch_strided(seeds, rows, off, p); w.finish()
        w = dat.AppendOnlyWriter(b, model, 12, 50); w.append_batch(seeds, rows[:, off:off + p].copy()); w.finish()
        assert a.read_bytes() == b.read_bytes()
    # encoded by several threads: same bytes
    big_n = 50_000
    big = rng.standard_normal((big_n, width)); big_seeds = rng.permutation(np.arange(1, big_n + 1)).astype(np.uint32)
    for threads in (1, 3, 8):
        f = tmp_path / f"mt{threads}.dat"
        w = dat.AppendOnlyWriter(f, 1, 12, 50); w.append_batch_strided(big_seeds, big, 12, 13, threads=threads); w.finish()
    assert (tmp_path / "mt1.dat").read_bytes() == (tmp_path / "mt3.dat").read_bytes() == (tmp_path / "mt8.dat").read_bytes()
    s2, e2, *_ = dat.read_append_file(tmp_path / "mt8.dat")
    assert np.array_equal(s2, big_seeds) and np.array_equal(e2, big[:, 12:25])
    w = dat.AppendOnlyWriter(tmp_path / "c.dat", 0, 2, 5)
    rc = dat.lib.jne_dat_append_batch_strided(w._w, seeds.ctypes.data, rows.ctypes.data, 10, 3, 2)
    assert rc < 0 and b"stride" in dat.lib.jne_dat_last_error()
    w.finish()


def test_damaged_finished_file(tmp_path):
    """A FINISHED file (trailer present) whose records do not add up to the trailer.  The reference's fast read fails on
    it (reader.rs:107-160: InvalidData / UnexpectedEof), its progress check then reports no progress (progress.rs:51),
    and its writer appends behind the damaged bytes (writer.rs:150-158).  Here: the read is an error, the progress check
    says restart, and the writer sets the damaged file aside as <name>.damaged instead of truncating or appending."""
    path = tmp_path / "dmg.dat"
    w = dat.AppendOnlyWriter(path, 0, 2, 50)
    w.append_batch(np.arange(1, 11), np.arange(20, dtype=np.float64).reshape(10, 2))
    w.finish()
    raw = path.read_bytes()
    rec = 1 + 1 + 16
    assert len(raw) == 18 + 10 * rec + 17
    # (a) the count byte of record 4 zeroed: "Invalid eigenvalue count: cannot be zero" in the reference
    bad = bytearray(raw); bad[18 + 3 * rec + 1] = 0
    path.write_bytes(bytes(bad))
    with pytest.raises(JneError, match="Invalid eigenvalue count"):
        dat.read_append_file(path)
    done, rem = dat.check_append_progress(path, 0, 2, 50, 10)
    assert done == 0 and list(rem) == list(range(1, 11))               # restart, as check_append_progress does
    # (b) three records cut out of the middle, trailer kept: the trailer promises 10, 7 are there
    path.write_bytes(raw[:18 + 4 * rec] + raw[18 + 7 * rec:])
    with pytest.raises(JneError, match="unexpected end of file|Incomplete ULEB128"):
        dat.read_append_file(path)
    assert dat.check_append_progress(path, 0, 2, 50, 10)[0] == 0
    w = dat.AppendOnlyWriter(path, 0, 2, 50)                            # nothing is dropped silently ...
    assert w.existing_records == 0
    w.append_batch([1], [[1.0, 2.0]])
    w.finish()
    aside = tmp_path / "dmg.dat.damaged"
    assert aside.exists() and aside.read_bytes() == raw[:18 + 4 * rec] + raw[18 + 7 * rec:]   # ... the damaged bytes are kept
    seeds, eigs, *_ = dat.read_append_file(path)
    assert list(seeds) == [1] and eigs.tolist() == [[1.0, 2.0]]
    # an INTERRUPTED file (no trailer) with a torn tail is still resumed in place (reader.rs:157-216)
    torn = tmp_path / "torn.dat"
    torn.write_bytes(raw[:18 + 6 * rec + 5])
    done, rem = dat.check_append_progress(torn, 0, 2, 50, 10)
    assert done == 6 and list(rem) == [7, 8, 9, 10]
    w = dat.AppendOnlyWriter(torn, 0, 2, 50)
    assert w.existing_records == 6
    w.finish()
    assert os.path.getsize(torn) == 18 + 6 * rec + 17 and not (tmp_path / "torn.dat.damaged").exists()


def test_random_access_batch(tmp_path):
    """jne_dat_batch_*: ranges filled out of order by several producers give the bytes of the serial writer; an
    uncommitted batch is invisible (poisoned head) and is truncated away; a short commit is refused."""
    rng = np.random.default_rng(5)
    n, p, width, off = 30000, 13, 62, 12
    seeds = np.concatenate([np.arange(100, 100 + n // 2), np.arange(5_000_000, 5_000_000 + n - n // 2)]).astype(np.uint32)   # 1-, 2-, 3- and 4-byte ULEB128s
    seeds[:200] = np.arange(1, 201)
    rows = rng.standard_normal((n, width))
    ref = tmp_path / "serial.dat"
    w = dat.AppendOnlyWriter(ref, 3, 12, 77); w.append_batch([7], np.ones((1, p))); w.append_batch_strided(seeds, rows, off, p); w.finish()
    f = tmp_path / "ra.dat"
    w = dat.AppendOnlyWriter(f, 3, 12, 77); w.append_batch([7], np.ones((1, p)))
    b = w.batch(seeds, p)
    cuts = [0, 1, 999, 1024, 1025, 7000, 20000, n]
    order = [3, 0, 6, 2, 5, 1, 4]                      # any order, record 0's range not first
    import threading
    th = [threading.Thread(target=lambda k=k: b.fill(cuts[k], rows[cuts[k]:cuts[k + 1]], off, threads=3)) for k in order]
    # mid-way the file is longer but a scan stops at the poisoned head of the batch: only the record before it is visible
    for t in th[:3]: t.start()
    for t in th[:3]: t.join()
    assert dat.file_info(f)["records"] == 1 and not dat.file_info(f)["has_trailer"]
    for t in th[3:]: t.start()
    for t in th[3:]: t.join()
    b.end(True)
    w.finish()
    assert f.read_bytes() == ref.read_bytes()
    # abort: the file returns to its length before the batch; a batch with a hole cannot be committed
    g = tmp_path / "abort.dat"
    w = dat.AppendOnlyWriter(g, 3, 12, 77); w.append_batch([7], np.ones((1, p)))
    size0 = 18 + 1 + 1 + 8 * p
    b = w.batch(seeds, p); b.fill(0, rows[:5000], off); b.end(False)
    assert os.path.getsize(g) == size0
    b = w.batch(seeds, p); b.fill(0, rows[:5000], off)
    with pytest.raises(JneError, match="5000 of 30000 records filled"):
        b.end(True)
    assert os.path.getsize(g) == size0
    w.append_batch_strided(seeds, rows, off, p, threads=4)           # the mt append is a batch filled in ranges
    w.finish()
    assert g.read_bytes() == ref.read_bytes()
    with pytest.raises(JneError, match="Eigenvalue count mismatch"):
        w2 = dat.AppendOnlyWriter(g, 3, 12, 77); w2.batch(seeds, p + 1)

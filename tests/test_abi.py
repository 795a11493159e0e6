"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/*.h declares,
argument validation that needs no device, the Python mirror of the reference's interface.
(-m "not gpu")"""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_functions():
    names = []
    for h in sorted((ROOT / "include").glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names += re.findall(r"\b(jne_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import johansen_null_eigenspectra_b200 as jne
    names = declared_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(str(Path(jne.__file__).parent / "libjne.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_is_plain_c():
    """The boundary must be bindgen-trivial: compile the headers as C."""
    import subprocess, tempfile
    for h in sorted((ROOT / "include").glob("*.h")):
        with tempfile.NamedTemporaryFile("w", suffix=".c") as f:
            f.write(f'#include "{h}"\nint main(void) {{ return 0; }}\n')
            f.flush()
            subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f.name], check=True)


def test_deviceless_entry_points():
    import johansen_null_eigenspectra_b200 as jne
    assert "sm_100a" in jne.version()
    # eigenvalues per run: src/data_storage/thread_manager.rs:40-44
    assert [jne.num_eigs(m, 12) for m in range(5)] == [12, 13, 12, 13, 12]
    assert jne.num_eigs(0, 2) == 2          # integration/basic_api.rs:35
    with pytest.raises(jne.JneError):
        jne.num_eigs(5, 3)
    with pytest.raises(jne.JneError):
        jne.num_eigs(0, 0)
    # F_alg = 2 T [p(p+1)/2 + p d]  
    assert jne.flops_per_run(0, 12, 10000) == 4.44e6
    assert jne.flops_per_run(1, 12, 10000) == 4.94e6
    assert jne.flops_per_run(0, 2, 1000) == 14e3


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import johansen_null_eigenspectra_b200 as jne
    if jne.lib.jne_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(jne.JneError) as e:
        jne.Engine()
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(jne.JneError):
        jne.calculate_eigenvalues(2, 100, 1, 0)


def test_product_never_imports_oracle():
    pkg = ROOT / "johansen_null_eigenspectra_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.hpp")):
        text = f.read_text()
        assert not re.search(r"(import\s+oracle|from\s+oracle|libjne_oracle|oracle/)", text), f"{f} reaches into oracle/"


def test_model_enum_mirror():
    """src/tests/johansen_models_test.rs: numbers, predicates, Default = model 2."""
    from johansen_null_eigenspectra_b200 import JohansenModel as M
    assert [m.to_number() for m in M.all_models()] == [0, 1, 2, 3, 4]
    assert M.from_number(3) is M.InterceptTrendUnrestrictedInterceptRestrictedTrend
    assert M.from_number(5) is None
    assert M.default() is M.InterceptNoTrendUnrestrictedIntercept
    assert [m.has_intercept() for m in M] == [False, True, True, True, True]
    assert [m.has_trend() for m in M] == [False, False, False, True, True]
    assert [m.num_eigs(5) for m in M] == [5, 6, 5, 6, 5]


def test_shard_bounds_cover_and_disjoint():
    from johansen_null_eigenspectra_b200.sharding import shard_bounds, weak_scaling_seeds
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    s = np.concatenate([weak_scaling_seeds(5, 4, r) for r in range(4)])
    assert np.array_equal(s, np.arange(1, 21, dtype=np.uint32))


@pytest.mark.parametrize("ne", [2, 4, 6, 8, 10, 12, 14, 16])
def test_jacobi_step_table(ne):
    """Host logic of K3 (make_jacobi_tables): a sweep of ne - 1 round-robin steps must rotate every pair exactly once,
    the pairs of a step must be disjoint, and every word must address the packed upper triangle consistently."""
    import johansen_null_eigenspectra_b200 as jne
    tab = jne.jacobi_table(ne)
    npair = ne // 2
    nblk = npair * (npair + 1) // 2
    assert tab.shape == (ne - 1, npair + nblk)
    tri = {}
    for i in range(ne):
        for j in range(i, ne):
            tri[i * ne - i * (i - 1) // 2 + (j - i)] = (i, j)
    assert sorted(tri) == list(range(ne * (ne + 1) // 2))
    seen = set()
    for step in range(ne - 1):
        pairs = []
        for l in range(npair):
            w = int(tab[step, l])
            (p, p2), (q, q2), (a, b) = tri[w & 255], tri[(w >> 8) & 255], tri[(w >> 16) & 255]
            assert p == p2 and q == q2 and (a, b) == (p, q) and p < q and w >> 24 == 0
            pairs.append((p, q))
        assert sorted(x for pq in pairs for x in pq) == list(range(ne))       # disjoint, everybody plays
        assert not (set(pairs) & seen)
        seen |= set(pairs)
        blk = 0
        touched = []
        for P1 in range(npair):
            for P2 in range(P1, npair):
                w = int(tab[step, npair + blk]); blk += 1
                (p1, q1), (p2, q2) = pairs[P1], pairs[P2]
                want = [tuple(sorted(e)) for e in ((p1, p2), (p1, q2), (q1, p2), (q1, q2))]
                got = [tri[(w >> s) & 255] for s in (0, 8, 16, 24)]
                assert got == want
                touched += got if P1 != P2 else [got[0], got[1], got[3]]
        assert sorted(touched) == sorted(tri.values())      # the blocks of a step tile the whole triangle once
    assert len(seen) == ne * (ne - 1) // 2


def test_jacobi_table_rejects_bad_sizes():
    import johansen_null_eigenspectra_b200 as jne
    for ne in (0, 1, 3, 18):
        with pytest.raises(jne.JneError):
            jne.jacobi_table(ne)


@pytest.mark.parametrize("T", [1, 2, 7, 31, 32, 33, 100, 960, 961, 1000, 2049, 10000])
def test_trend_weight_table(T):
    """Host logic of the AUX kernels (make_aux_table): the table fed to the tensor pipe holds, per local step of each
    of the four time segments, the exact integer tail sums of w1 / w2 and the weights themselves -- and with them
    sum_s weight_s dB_s reproduces the trend moments of the path, sum_i w_i c_i (c = segment-local path before step i),
    which is the identity the kernels rely on (summation by parts)."""
    import johansen_null_eigenspectra_b200 as jne
    tab = jne.trend_weight_table(T)
    by_block, by_epoch = 8 * ((T + 31) // 32), 128 * ((T + 511) // 512)
    seg_len = by_block if 0.95 * by_epoch > by_block else by_epoch    # seg_len_for (jne_api.cu)
    assert tab.shape == (seg_len, 4, 4)
    w1 = [2 * i + 1 - T for i in range(T)]
    w2 = [3 * w * w - (T * T - 1) for w in w1]
    rng = np.random.default_rng(T)
    z = rng.standard_normal(T)
    for k in range(4):
        a, e = min(k * seg_len, T), min((k + 1) * seg_len, T)
        for j in range(seg_len):
            i = a + j
            if i >= e:
                assert not tab[j, :, k].any()
                continue
            want = (sum(w1[i + 1:e]), sum(w2[i + 1:e]), w2[i], w1[i])
            for m in range(4):
                assert tab[j, m, k] == float(want[m]), (k, j, m)
        if e > a:
            c = np.concatenate([[0.0], np.cumsum(z[a:e])[:-1]])       # segment-local path before each step
            n = e - a
            for m, w in ((0, w1), (1, w2)):
                direct = float(np.dot(np.array(w[a:e], dtype=np.float64), c))
                by_parts = float(np.dot(tab[:n, m, k], z[a:e]))
                assert abs(direct - by_parts) <= 1e-11 * max(1.0, np.abs(np.array(w[a:e], dtype=np.float64)).sum())


def test_trend_weight_table_rejects_bad_sizes():
    import johansen_null_eigenspectra_b200 as jne
    for T in (0, (1 << 22) + 1):
        with pytest.raises(jne.JneError):
            jne.trend_weight_table(T)


def test_rust_shim_binds_only_declared_and_exported_symbols():
    """ffi/rust/src/gpu_ffi.rs (source only -- no Rust toolchain in this image): every function of its extern "C" block
    must be declared in include/*.h and exported by libjne.so, with the same number of arguments as the C prototype."""
    import johansen_null_eigenspectra_b200 as jne
    text = (ROOT / "ffi" / "rust" / "src" / "gpu_ffi.rs").read_text()
    block = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    fns = re.findall(r"fn (jne_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]+)?;", block, flags=re.S)
    assert len(fns) >= 10
    header = ""
    for h in sorted((ROOT / "include").glob("*.h")):
        header += re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
    lib = ctypes.CDLL(str(Path(jne.__file__).parent / "libjne.so"))
    for name, args in fns:
        assert hasattr(lib, name), name
        proto = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", header, flags=re.S)
        assert proto, f"{name} is not declared in include/*.h"
        c_args = [a for a in proto.group(1).split(",") if a.strip() and a.strip() != "void"]
        rust_args = [a for a in args.split(",") if a.strip()]
        assert len(c_args) == len(rust_args), (name, c_args, rust_args)

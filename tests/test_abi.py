"""CPU tests of the drop-in boundary: the shared libraries load, export every
symbol include/trmf_b200.h declares, PyMatrix has the reference's 80-byte layout,
and the product fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sps

import trmf
from trmf.rf_util import PyMatrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "trmf_b200.h")
CORELIB = os.path.join(ROOT, "exp-trmf-nips16_b200", "trmf", "corelib")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(c_trmf_train|trmf_b200_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_declares_the_reference_entry_point():
    syms = declared_symbols()
    assert "c_trmf_train" in syms and len(syms) >= 20


@pytest.mark.parametrize("lib", ["trmf_float32.so", "trmf_float64.so"])
def test_library_exports_every_declared_symbol(lib):
    dll = ctypes.CDLL(os.path.join(CORELIB, lib))
    for name in declared_symbols():
        assert hasattr(dll, name), "{} does not export {}".format(lib, name)
    dll.trmf_b200_value_bytes.restype = ctypes.c_int
    assert dll.trmf_b200_value_bytes() == (4 if "32" in lib else 8)


def test_pack_thread_budget_follows_the_ranks_on_the_host(monkeypatch):
    """csrc/trmf_b200.cu:pack_thread_budget -- host threads a host-buffer session packs row indices with: the cores divided by
    LOCAL_WORLD_SIZE (one process per GPU shares the host), TRMF_B200_PACK_THREADS pins it, never below 2 (one packer + the
    coordinating thread) nor above 32; below 4 the session uploads plain indices instead of packing."""
    dll = ctypes.CDLL(os.path.join(CORELIB, "trmf_float32.so"))
    dll.trmf_b200_pack_threads.restype = ctypes.c_int
    for name in ("LOCAL_WORLD_SIZE", "TRMF_B200_PACK_THREADS"):
        monkeypatch.delenv(name, raising=False)
    cores = os.cpu_count() or 1
    assert dll.trmf_b200_pack_threads() == max(2, min(cores, 32))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert dll.trmf_b200_pack_threads() == max(2, min(cores // 8, 32))
    monkeypatch.setenv("TRMF_B200_PACK_THREADS", "5")
    assert dll.trmf_b200_pack_threads() == 5
    monkeypatch.setenv("TRMF_B200_PACK_THREADS", "100")
    assert dll.trmf_b200_pack_threads() == 32


def test_pymatrix_layout_is_the_reference_pod():
    # rf_util.py:41-52 / rf_matrix.h:3407-3415: 80 bytes, fixed offsets
    assert ctypes.sizeof(PyMatrix) == 80
    want = dict(rows=0, cols=8, nnz=16, row_ptr=24, col_ptr=32, row_idx=40, col_idx=48, val=56, val_t=64, type=72)
    for name, off in want.items():
        assert getattr(PyMatrix, name).offset == off


def test_pymatrix_sparse_twin_storage_is_bit_exact():
    rng = np.random.RandomState(0)
    A = sps.random(37, 23, density=0.3, random_state=rng, format="csr")
    A.data[:] = rng.randn(A.nnz)
    pm = PyMatrix(A, np.float32)
    csr, csc = sps.csr_matrix(A), sps.csc_matrix(A)
    csr.sort_indices(); csc.sort_indices()
    b = pm.py_buf
    assert pm.type == PyMatrix.SPARSE and pm.nnz == A.nnz and (pm.rows, pm.cols) == (37, 23)
    assert b["row_ptr"].dtype == np.uint64 and b["col_idx"].dtype == np.uint32
    assert np.array_equal(b["row_ptr"], csr.indptr) and np.array_equal(b["col_idx"], csr.indices)
    assert np.array_equal(b["col_ptr"], csc.indptr) and np.array_equal(b["row_idx"], csc.indices)
    assert np.array_equal(b["val_t"], csr.data.astype(np.float32))
    assert np.array_equal(b["val"], csc.data.astype(np.float32))
    # empty and ragged
    E = PyMatrix(sps.csr_matrix((5, 7)), np.float64)
    assert E.nnz == 0 and np.array_equal(E.py_buf["row_ptr"], np.zeros(6, np.uint64))


def test_pymatrix_dense_majors():
    W = np.zeros((6, 3), order="C")
    L = np.zeros((4, 3), order="F")
    assert PyMatrix(W, np.float32).type == PyMatrix.DENSE_ROWMAJOR
    assert PyMatrix(L, np.float32).type == PyMatrix.DENSE_COLMAJOR
    assert PyMatrix(np.zeros((6, 1)), np.float32, major="row").type == PyMatrix.DENSE_ROWMAJOR
    assert PyMatrix(np.zeros((1, 3), order="F"), np.float32, major="col").type == PyMatrix.DENSE_COLMAJOR


def test_model_initialize_follows_the_reference_stream():
    # trmf.py:224-236: seed, then W ~ rand, H ~ rand, lag_val ~ randn in that order
    Y = np.zeros((30, 11), dtype=np.float64)
    m = trmf.Model.initialize(Y, [5, 1, 2], 4, seed=7)
    np.random.seed(7)
    W = np.random.rand(30, 4); H = np.random.rand(11, 4); L = np.random.randn(3, 4)
    assert np.array_equal(m.W, W) and np.array_equal(m.H, H) and np.array_equal(m.lag_val, L)
    assert m.lag_set.dtype == np.uint32 and list(m.lag_set) == [1, 2, 5]
    assert m.W.flags.c_contiguous and m.lag_val.flags.f_contiguous
    assert (m.m, m.n, m.k) == (30, 11, 4) and m.transform is None


def test_forecast_and_warm_start():
    d = trmf.Model.syn_gen(80, 9, 3, [1, 2, 4], seed=3, dtype=np.float64)
    m = trmf.Model.initialize(d["Y"], d["lag_set"], 3, seed=0)
    m.W[:] = d["W"]; m.H[:] = d["H"]; m.lag_val[:] = d["lag_val"]
    Wn = m.latent_forecast(5)
    lags = d["lag_set"].astype(int)
    for i in range(80, 85):
        assert np.allclose(Wn[i], (Wn[i - lags] * d["lag_val"]).sum(axis=0))
    Yn, Wtail = m.forecast(5, threshold=None)
    assert np.allclose(Yn, Wn[80:] @ d["H"].T) and np.allclose(Wtail, Wn[80:])
    Yc, _ = m.forecast(5, threshold=0)
    assert (Yc >= 0).all()
    Y2 = np.vstack([d["Y"], Yn])
    m2 = trmf.Model.initialize(Y2, d["lag_set"], 3, seed=0, warm_start_model=m)
    assert np.allclose(m2.W, Wn) and np.array_equal(m2.H, m.H) and np.array_equal(m2.lag_val, m.lag_val)


def test_transform_metrics_save_load(tmp_path):
    rng = np.random.RandomState(0)
    Y = rng.rand(40, 6) * 5 + 2
    Y[:, 2] = 1.5  # zero std -> treated as 1 (trmf.py:86)
    tr = trmf.NormalizedTransform(Y)
    Z = tr.preprocess(Y)
    assert np.allclose(Z.mean(axis=0)[[0, 1, 3]], 0) and np.allclose(tr.postprocess(Z), Y)
    m = trmf.Model.initialize(Y, [1, 2], 3, seed=1, transform=True)
    assert isinstance(m.transform, trmf.NormalizedTransform)
    m.save(str(tmp_path / "mdl"))
    m2 = trmf.Model.load(str(tmp_path / "mdl"))
    assert np.array_equal(m2.W, m.W) and np.array_equal(m2.lag_val, m.lag_val) and m2.lag_val.flags.f_contiguous
    assert np.allclose(m2.transform.a, m.transform.a)
    met = trmf.Metrics.generate(Y[-10:], Y[-10:] * 1.1)
    assert abs(met.nd - 0.1) < 1e-12 and abs(met.mape - 0.1) < 1e-12 and "nd=" in str(met)
    assert trmf.Metrics.default().m_nd == 1e10


def test_no_silent_cpu_fallback():
    """Without a visible CUDA device train() must raise, never compute on the host."""
    lib = trmf.trmf._clib.clib_float32
    if lib.trmf_b200_device_count() > 0:
        pytest.skip("a GPU is visible")
    d = trmf.Model.syn_gen(50, 8, 2, [1, 2], seed=0)
    m = trmf.Model.initialize(d["Y"], d["lag_set"], 2, seed=0)
    W0 = m.W.copy()
    with pytest.raises(RuntimeError, match="GPU-only"):
        trmf.train(d["Y"], m, missing=False)
    assert np.array_equal(m.W, W0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "exp-trmf-nips16_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f


def test_pymatrix_single_orientation_leaves_the_other_half_null():
    """twin=False (what trmf.train passes down): only the orientation the caller's matrix already has is
    marshalled; the library derives the other one on the device."""
    import scipy.sparse as sps
    from trmf.rf_util import PyMatrix
    A = sps.random(30, 20, density=0.3, format="csr", random_state=np.random.RandomState(0), dtype=np.float64)
    csr = PyMatrix(A, np.float32, twin=False)
    assert csr.type == PyMatrix.SPARSE and csr.nnz == A.nnz and csr.row_ptr and csr.col_idx and csr.val_t
    assert not csr.col_ptr and not csr.row_idx and not csr.val
    csc = PyMatrix(A.tocsc(), np.float32, twin=False)
    assert csc.nnz == A.nnz and csc.col_ptr and csc.row_idx and csc.val
    assert not csc.row_ptr and not csc.col_idx and not csc.val_t
    both = PyMatrix(A, np.float32)
    assert both.row_ptr and both.col_ptr
    assert np.array_equal(both.py_buf["row_ptr"], csr.py_buf["row_ptr"]) and np.array_equal(both.py_buf["col_idx"], csr.py_buf["col_idx"])
    assert np.array_equal(both.py_buf["col_ptr"], csc.py_buf["col_ptr"]) and np.array_equal(both.py_buf["row_idx"], csc.py_buf["row_idx"])


def test_header_is_plain_c_and_pymatrix_is_80_bytes(tmp_path):
    """include/trmf_b200.h is the C ABI: it must compile as C99 (no C++ or torch types) and lay PyMatrix out like the
    reference's POD (rf_matrix.h:3399-3415); a C caller resolves every declared entry point with dlsym."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    syms = [s for s in declared_symbols()]
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <dlfcn.h>\n#include <stdio.h>\n#include <stddef.h>\n#include "trmf_b200.h"\n'
        "int main(int argc, char **argv) {\n"
        "    if (sizeof(PyMatrix) != 80 || offsetof(PyMatrix, type) != 72 || offsetof(PyMatrix, val_t) != 64) return 2;\n"
        "    void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);\n"
        '    if (!h) { fprintf(stderr, "%s\\n", dlerror()); return 3; }\n'
        "    const char *names[] = {" + ", ".join('"%s"' % s for s in syms) + "};\n"
        "    for (unsigned i = 0; i < sizeof names / sizeof *names; ++i)\n"
        '        if (!dlsym(h, names[i])) { fprintf(stderr, "missing %s\\n", names[i]); return 4; }\n'
        "    (void)argc; return 0;\n}\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-D_GNU_SOURCE", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-ldl"], check=True, capture_output=True)
    for lib in ("trmf_float32.so", "trmf_float64.so"):
        proc = subprocess.run([str(exe), os.path.join(CORELIB, lib)], capture_output=True, text=True)
        assert proc.returncode == 0, (lib, proc.stderr)


def test_pack_bitmap_is_the_per_series_mask():
    """TRMF_SPARSE_BITMAP host side (trmf/rf_util.py:pack_bitmap): bit (i & 31) of word [j][i >> 5] <=> (i, j) stored."""
    import scipy.sparse as sps
    from trmf.rf_util import PyMatrix, pack_bitmap
    rng = np.random.RandomState(0)
    for T, n, d in [(300, 200, 0.7), (33, 4, 0.9), (32, 3, 1.0), (1, 1, 1.0), (64, 64, 0.0), (70001, 3, 0.5)]:
        m = sps.random(T, n, density=d, format="csc", random_state=rng)
        m.sort_indices()
        bm = pack_bitmap(m.indptr, m.indices, T).reshape(n, (T + 31) // 32)
        dense = np.zeros((n, ((T + 31) // 32) * 32), dtype=bool)
        dense[np.repeat(np.arange(n), np.diff(m.indptr)), m.indices] = True
        assert np.array_equal(np.packbits(dense, axis=1, bitorder="little").view(np.uint32), bm)
    pm = PyMatrix(m, np.float32, twin=False, pack=True)
    assert pm.type == PyMatrix.SPARSE_BITMAP == 5 and not pm.row_ptr and not pm.col_idx and not pm.val_t
    assert len(pm.py_buf["row_idx"]) == n * ((T + 31) // 32) and pm.nnz == m.nnz


def test_coo_pymatrix_keeps_duplicate_entries_like_the_reference():
    """reference rf_util.py:100-119: a coo_matrix goes through bincount + argsort(row * ncols + col), so duplicate (i, j)
    entries stay separate observations and nnz = len(data); scipy's tocsr() would sum them."""
    import scipy.sparse as sps
    from trmf.rf_util import PyMatrix
    rng = np.random.RandomState(1)
    T, n, m = 40, 30, 500
    row, col, val = rng.randint(0, T, m), rng.randint(0, n, m), rng.randn(m)
    coo = sps.coo_matrix((val, (row, col)), shape=(T, n))
    assert coo.tocsr().nnz < m                      # there are duplicates
    pm = PyMatrix(coo, np.float64)
    assert pm.nnz == m and pm.type == PyMatrix.SPARSE
    b = pm.py_buf

    def ref(major, minor, nmajor, nminor):          # the reference's coo_to_csr, stable
        indptr = np.cumsum(np.bincount(major + 1, minlength=nmajor + 1)).astype(np.uint64)
        order = np.argsort(major * nminor + minor, kind="stable")
        return indptr, minor[order].astype(np.uint32), val[order]
    for got, want in zip((b["row_ptr"], b["col_idx"], b["val_t"]), ref(row, col, T, n)):
        assert np.array_equal(got, want)
    for got, want in zip((b["col_ptr"], b["row_idx"], b["val"]), ref(col, row, n, T)):
        assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        PyMatrix(coo, np.float64, pack=True)


def test_host_packing_of_row_indices_matches_the_numpy_mask():
    """csrc/trmf_b200.cu:pack_bitmap_slabs (gap method for series that are dense inside their span, word-wise OR otherwise,
    packed slab by slab by worker threads): same words as trmf.rf_util.pack_bitmap; unsorted / duplicate / out-of-range
    indices are declined (the caller then uploads plain indices).  No device involved."""
    import ctypes
    import scipy.sparse as sps
    from trmf.rf_util import pack_bitmap
    lib = ctypes.CDLL(os.path.join(CORELIB, "trmf_float32.so"))
    fn = lib.trmf_b200_pack_bitmap_host
    fn.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    rng = np.random.RandomState(5)

    def run(T, col_ptr, row_idx):
        n = len(col_ptr) - 1
        out = np.full(n * ((T + 31) // 32), 0xdeadbeef, dtype=np.uint32)
        rc = fn(T, n, col_ptr.ctypes.data, row_idx.ctypes.data, out.ctypes.data)
        return rc, out

    for T, n, d in [(300, 200, 0.9), (1000, 64, 0.97), (33, 4, 0.9), (32, 3, 1.0), (1, 1, 1.0), (64, 64, 0.0), (70001, 5, 0.5),
                    (5000, 40, 0.05), (4099, 37, 0.6), (257, 100, 0.995)]:
        m = sps.random(T, n, density=d, format="csc", random_state=rng)
        m.sort_indices()
        col_ptr, row_idx = m.indptr.astype(np.uint64), m.indices.astype(np.uint32)
        buf = row_idx if len(row_idx) else np.zeros(1, dtype=np.uint32)   # (a valid pointer for the empty matrix)
        rc, out = run(T, col_ptr, buf)
        assert rc == 0
        assert np.array_equal(out, pack_bitmap(col_ptr, row_idx, T)), (T, n, d)
    # long gaps that span several words, runs that end exactly on word boundaries, a single entry, first / last row only
    T = 1000
    series = [np.arange(0, 32), np.arange(31, 65), np.r_[np.arange(0, 10), np.arange(500, 520), 999], np.array([999]), np.array([0]),
              np.r_[0, 999], np.arange(0, 1000), np.arange(1, 999, 2), np.r_[np.arange(0, 64), np.arange(96, 128)]]
    col_ptr = np.r_[0, np.cumsum([len(x) for x in series])].astype(np.uint64)
    row_idx = np.concatenate(series).astype(np.uint32)
    rc, out = run(T, col_ptr, row_idx)
    assert rc == 0 and np.array_equal(out, pack_bitmap(col_ptr, row_idx, T))
    # declined: a swapped pair inside a dense series, a duplicate, an index >= T
    base = np.arange(0, 400, dtype=np.uint32)
    for bad in (np.r_[base[:100], base[101], base[100], base[102:]], np.r_[base[:50], base[49:]], np.r_[base[:-1], T + 5],
                np.r_[base[200:], base[:200]]):
        cp = np.array([0, len(bad)], dtype=np.uint64)
        rc, _ = run(T, cp, np.ascontiguousarray(bad, dtype=np.uint32))
        assert rc == 1
        # ... and the same series among eleven good ones: four series are packed in lock step there, with the ascent check folded
        # into the packing pass (pack_series_words_multi)
        for where in (0, 2, 3, 5):
            group = [base.copy() for _ in range(12)]
            group[where] = np.ascontiguousarray(bad, dtype=np.uint32)
            cp = np.r_[0, np.cumsum([len(x) for x in group])].astype(np.uint64)
            rc, _ = run(T, cp, np.concatenate(group).astype(np.uint32))
            assert rc == 1, where
    # an out-of-range index in the MIDDLE of an otherwise ascending series (the last one is in range), empty series inside a group
    group = [base.copy() for _ in range(12)]
    group[1] = np.r_[base[:300], 5000, base[300:]].astype(np.uint32)
    cp = np.r_[0, np.cumsum([len(x) for x in group])].astype(np.uint64)
    rc, _ = run(T, cp, np.concatenate(group).astype(np.uint32))
    assert rc == 1
    group = [base.copy(), np.zeros(0, dtype=np.uint32), np.arange(5, 1000, 7, dtype=np.uint32), np.zeros(0, dtype=np.uint32)] + [base.copy() for _ in range(8)]
    cp = np.r_[0, np.cumsum([len(x) for x in group])].astype(np.uint64)
    ri = np.concatenate(group).astype(np.uint32)
    rc, out = run(T, cp, ri)
    assert rc == 0 and np.array_equal(out, pack_bitmap(cp, ri, T))

"""GPU test: the device CSC -> CSR transpose (csrc/ingest.cuh) is bit-identical to scipy's conversion, i.e. to the
arrays the reference's PyMatrix builds on the host (rf_util.py:88-98), including empty rows / columns, a single
entry, ragged rows, and both value types; and a session created from a CSC-only PyMatrix trains to the same
factors as one created from the full twin storage."""
import numpy as np
import pytest
import scipy.sparse as sps

import cases

pytestmark = pytest.mark.gpu


def _random_csc(T, n, density, seed, dtype):
    rng = np.random.RandomState(seed)
    m = sps.random(T, n, density=density, format="csc", random_state=rng, dtype=np.float64)
    m.data = rng.randn(m.nnz)
    m = m.astype(dtype)
    m.sort_indices()
    return m


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("T,n,density", [(300, 200, 0.7), (1, 1, 1.0), (257, 129, 0.05), (40, 3000, 0.3), (5000, 7, 0.9),
                                         (70000, 3, 0.5), (64, 64, 0.0)])
def test_device_transpose_is_bit_exact(dtype, T, n, density):
    from trmf.session import csr_from_csc
    csc = _random_csc(T, n, density, 11 + T + n, dtype)
    if csc.nnz > 10:   # one empty column and one empty row
        lil = csc.tolil()
        lil[:, n // 2] = 0
        lil[T // 3, :] = 0
        csc = lil.tocsc().astype(dtype)
        csc.eliminate_zeros()
        csc.sort_indices()
    ref = csc.tocsr()
    ref.sort_indices()
    row_ptr, col_idx, val_t = csr_from_csc(csc, dtype)
    assert np.array_equal(row_ptr, ref.indptr.astype(np.uint64))
    assert np.array_equal(col_idx, ref.indices.astype(np.uint32))
    assert np.array_equal(val_t.view(np.uint8), ref.data.astype(dtype).view(np.uint8))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_csc_only_pymatrix_trains_like_twin_storage(dtype, monkeypatch):
    from trmf.rf_util import PyMatrix
    from trmf.session import Session
    p = cases.make_problem(240, 150, 16, [1, 3, 7], 0.6, seed=5)
    Y = p["Ysp"].astype(dtype)
    W0, H0, L0 = (a.astype(dtype) for a in (p["W0"], p["H0"], p["L0"]))
    outs = []
    for mode in ("device", "host", "csc_only", "csr_only", "no_slabs"):
        if mode == "host":
            monkeypatch.setenv("TRMF_B200_HOST_CSR", "1")
        else:
            monkeypatch.delenv("TRMF_B200_HOST_CSR", raising=False)
        if mode == "no_slabs":
            monkeypatch.setenv("TRMF_B200_NO_SLAB_UPLOAD", "1")
        if mode == "csc_only":
            pm = PyMatrix(Y.tocsc(), dtype, twin=False)
            assert not pm.row_ptr and not pm.col_idx and not pm.val_t and pm.col_ptr
        elif mode == "csr_only":
            pm = PyMatrix(Y.tocsr(), dtype, twin=False)
            assert not pm.col_ptr and not pm.row_idx and not pm.val and pm.row_ptr
        else:
            pm = PyMatrix(Y, dtype)
        s = Session(pm, p["lags"], W0, H0, L0, missing=True, dtype=dtype, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
        s.train(max_iter=2, period_W=1, period_H=1, period_Lag=1)
        outs.append(s.download())
        s.close()
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)      # same arrays in HBM either way -> bitwise the same factors


def test_slab_upload_trains_like_a_single_copy(monkeypatch):
    """Above 4M entries a host-buffer session uploads the CSC in nnz-balanced series slabs on a copy stream and
    the first F-update runs slab by slab; the factors must not depend on that."""
    from trmf.rf_util import PyMatrix
    from trmf.session import Session
    dtype = np.float32
    rng = np.random.RandomState(3)
    T, n, k = 3000, 2500, 40
    mask = rng.rand(T, n) < 0.62
    mask[:, 7] = False
    Y = sps.csr_matrix(np.where(mask, rng.randn(T, n) + 3.0, 0.0).astype(dtype))
    assert Y.nnz >= (1 << 22)
    lags = np.array([1, 7, 24], dtype=np.uint32)
    W0, H0, L0 = rng.rand(T, k).astype(dtype), rng.rand(n, k).astype(dtype), rng.randn(3, k).astype(dtype)
    outs = []
    for slabs in (True, False):
        if slabs:
            monkeypatch.delenv("TRMF_B200_NO_SLAB_UPLOAD", raising=False)
        else:
            monkeypatch.setenv("TRMF_B200_NO_SLAB_UPLOAD", "1")
        s = Session(PyMatrix(Y, dtype), lags, W0, H0, L0, missing=True, dtype=dtype, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
        s.train(max_iter=2, period_W=1, period_H=1, period_Lag=1)
        outs.append(s.download())
        s.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def _naive_bitmap(csc, T):
    words = (T + 31) // 32
    out = np.zeros((csc.shape[1], words), dtype=np.uint32)
    for j in range(csc.shape[1]):
        for i in csc.indices[csc.indptr[j]:csc.indptr[j + 1]]:
            out[j, i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    return out.ravel()


@pytest.mark.parametrize("T,n,density", [(300, 200, 0.7), (1, 1, 1.0), (31, 5, 0.5), (32, 9, 1.0), (33, 4, 0.9), (257, 129, 0.05),
                                         (5000, 7, 0.9), (70001, 3, 0.5), (64, 64, 0.0)])
def test_bitmap_ingest_expands_to_the_identical_row_idx(T, n, density):
    """TRMF_SPARSE_BITMAP (include/trmf_b200.h): one bitmap per series travels instead of nnz row indices; the device
    expansion (csrc/ingest.cuh: bitmap_expand_kernel) must give back the canonical CSC row_idx array bit for bit."""
    import ctypes
    from trmf.rf_util import pack_bitmap
    from trmf.session import _lib
    csc = _random_csc(T, n, density, 5 + T + n, np.float32)
    if csc.nnz > 10:
        lil = csc.tolil()
        lil[:, n // 2] = 0
        csc = lil.tocsc().astype(np.float32)
        csc.eliminate_zeros()
        csc.sort_indices()
    col_ptr = csc.indptr.astype(np.uint64)
    row_idx = csc.indices.astype(np.uint32)
    bm = pack_bitmap(col_ptr, row_idx, T)
    if T * n <= 20000:
        assert np.array_equal(bm, _naive_bitmap(csc, T))
    lib = _lib(np.float32)
    out = np.full(max(csc.nnz, 1), 0xdeadbeef, dtype=np.uint32)
    lib.trmf_b200_bitmap_expand.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int32]
    assert lib.trmf_b200_bitmap_expand(T, n, csc.nnz, col_ptr.ctypes.data, bm.ctypes.data, out.ctypes.data, 0) == 0
    assert np.array_equal(out[:csc.nnz], row_idx)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bitmap_packed_pymatrix_trains_like_plain_indices(dtype):
    from trmf.rf_util import PyMatrix
    from trmf.session import Session
    p = cases.make_problem(250, 140, 16, [1, 3, 7], 0.8, seed=8)
    Y = p["Ysp"].astype(dtype)
    W0, H0, L0 = (a.astype(dtype) for a in (p["W0"], p["H0"], p["L0"]))
    outs = []
    for pack in (False, True):
        pm = PyMatrix(Y.tocsc(), dtype, twin=False, pack=pack)
        assert pm.type == (PyMatrix.SPARSE_BITMAP if pack else PyMatrix.SPARSE)
        s = Session(pm, p["lags"], W0, H0, L0, missing=True, dtype=dtype, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
        s.train(max_iter=2, period_W=1, period_H=1, period_Lag=1)
        outs.append(s.download())
        s.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_host_packed_index_upload_trains_identically(monkeypatch):
    """A large mostly-observed CSC given with plain row indices: the library packs them into per-series bitmaps on the
    host cores inside the call (4.1 instead of 8 bytes per entry over PCIe) -- same arrays in HBM, same factors,
    with the packing on (default), off, and for a caller-packed PyMatrix."""
    from trmf.rf_util import PyMatrix
    from trmf.session import Session
    dtype = np.float32
    rng = np.random.RandomState(4)
    T, n, k = 3001, 2300, 24
    mask = rng.rand(T, n) < 0.7
    mask[:, 11] = False
    mask[:, 12] = True
    Y = sps.csc_matrix(np.where(mask, rng.randn(T, n) + 3.0, 0.0).astype(dtype))
    Y.sort_indices()
    assert Y.nnz >= (1 << 22)
    lags = np.array([1, 7, 24], dtype=np.uint32)
    W0, H0, L0 = rng.rand(T, k).astype(dtype), rng.rand(n, k).astype(dtype), rng.randn(3, k).astype(dtype)
    outs = []
    for mode in ("host_pack", "plain", "caller_pack", "one_thread"):
        monkeypatch.delenv("TRMF_B200_NO_HOST_PACK", raising=False)
        monkeypatch.delenv("TRMF_B200_PACK_THREADS", raising=False)
        if mode == "plain":
            monkeypatch.setenv("TRMF_B200_NO_HOST_PACK", "1")
        if mode == "one_thread":
            monkeypatch.setenv("TRMF_B200_PACK_THREADS", "1")
        pm = PyMatrix(Y, dtype, twin=False, pack=(mode == "caller_pack"))
        s = Session(pm, lags, W0, H0, L0, missing=True, dtype=dtype, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
        s.train(max_iter=1, period_W=1, period_H=1, period_Lag=1)
        outs.append(s.download())
        s.close()
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)


def test_slab_upload_of_a_thin_matrix_is_not_packed_and_trains_identically(monkeypatch):
    """Below 1/8 density the per-series bitmaps would not pay (n * ceil(T/32) * 4 bytes > nnz): the slab upload sends plain
    indices.  (Round-2 regression: this path once issued only 5 of its 8 slabs.)  Same factors as a single-copy upload."""
    from trmf.rf_util import PyMatrix
    from trmf.session import Session
    dtype = np.float32
    rng = np.random.RandomState(9)
    T, n, k = 20000, 3000, 16
    Y = sps.random(T, n, density=0.075, format="csc", random_state=rng, dtype=np.float64)
    Y.data = (rng.randn(Y.nnz) + 3.0)
    Y = Y.astype(dtype)
    Y.sort_indices()
    assert Y.nnz >= (1 << 22) and n * ((T + 31) // 32) * 4 > Y.nnz
    lags = np.array([1, 7, 24], dtype=np.uint32)
    W0, H0, L0 = rng.rand(T, k).astype(dtype), rng.rand(n, k).astype(dtype), rng.randn(3, k).astype(dtype)
    outs = []
    for slabs in (True, False):
        if slabs:
            monkeypatch.delenv("TRMF_B200_NO_SLAB_UPLOAD", raising=False)
        else:
            monkeypatch.setenv("TRMF_B200_NO_SLAB_UPLOAD", "1")
        s = Session(PyMatrix(Y, dtype, twin=False), lags, W0, H0, L0, missing=True, dtype=dtype, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
        s.train(max_iter=1, period_W=1, period_H=1, period_Lag=1)
        outs.append(s.download())
        s.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_coo_with_duplicates_trains_like_the_reference_core():
    """Duplicate (i, j) entries of a COO matrix are separate observations in the reference (rf_util.py:100-119); the same
    PyMatrix struct is handed to the compiled reference core and to the CUDA library (float64, one outer iteration)."""
    import ctypes
    from oracle import abi
    from trmf.rf_util import PyMatrix
    from trmf.session import _lib
    if not abi.ref_available(np.float64):
        pytest.skip("oracle/_ref did not travel")
    rng = np.random.RandomState(2)
    T, n, k, m = 120, 80, 8, 6000
    row, col = rng.randint(0, T, m), rng.randint(0, n, m)
    coo = sps.coo_matrix((rng.randn(m) + 2.0, (row, col)), shape=(T, n))
    assert coo.tocsr().nnz < m
    lags = np.array([1, 2, 5], dtype=np.uint32)
    W0, H0, L0 = rng.rand(T, k), rng.rand(n, k), rng.randn(3, k)
    outs = []
    for which in ("reference", "cuda"):
        # (plain CDLL handles without argtypes: the same PyMatrix struct goes to both libraries)
        lib = ctypes.CDLL(abi.ref_lib_path(np.float64)) if which == "reference" else ctypes.CDLL(_lib(np.float64)._name)
        pY = PyMatrix(coo, np.float64)
        pW, pH = PyMatrix(W0.copy(), np.float64, major="row"), PyMatrix(H0.copy(), np.float64, major="row")
        pL = PyMatrix(np.asfortranarray(L0.copy()), np.float64, major="col")
        lib.c_trmf_train.restype = None
        lib.c_trmf_train(ctypes.byref(pY), lags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.c_uint32(3), ctypes.byref(pW),
                         ctypes.byref(pH), ctypes.byref(pL), ctypes.c_int(1), ctypes.c_double(0.5), ctypes.c_double(5.0), ctypes.c_double(0.5),
                         ctypes.c_int32(1), ctypes.c_int32(1), ctypes.c_int32(1), ctypes.c_int32(1), ctypes.c_int32(2), ctypes.c_int32(1),
                         ctypes.c_int32(0))
        outs.append((pW.py_buf["val"].copy(), pH.py_buf["val"].copy(), pL.py_buf["val"].copy()))
    for a, b in zip(*outs):
        assert cases.rel(b, a) < 1e-9


def test_host_packing_declines_unsorted_indices(monkeypatch):
    """The C ABI takes the CSC arrays as they are (the reference's core never looks at the order inside a series).  A bitmap
    cannot carry an unsorted series: the in-call host packing must notice and send plain indices -- same factors as with the
    packing switched off."""
    from trmf.rf_util import PyMatrix
    from trmf.session import Session
    dtype = np.float32
    rng = np.random.RandomState(6)
    T, n, k = 3001, 2300, 16
    Y = sps.csc_matrix(np.where(rng.rand(T, n) < 0.7, rng.randn(T, n) + 3.0, 0.0).astype(dtype))
    Y.sort_indices()
    assert Y.nnz >= (1 << 22)
    pm = PyMatrix(Y, dtype, twin=False)
    ri, va, cp = pm.py_buf["row_idx"], pm.py_buf["val"], pm.py_buf["col_ptr"].astype(np.int64)
    j = 1234                                        # swap two entries of one series (indices and values alike)
    a, b = cp[j] + 3, cp[j] + 9
    ri[[a, b]] = ri[[b, a]]
    va[[a, b]] = va[[b, a]]
    lags = np.array([1, 7, 24], dtype=np.uint32)
    W0, H0, L0 = rng.rand(T, k).astype(dtype), rng.rand(n, k).astype(dtype), rng.randn(3, k).astype(dtype)
    outs = []
    for pack in (True, False):
        if pack:
            monkeypatch.delenv("TRMF_B200_NO_HOST_PACK", raising=False)
        else:
            monkeypatch.setenv("TRMF_B200_NO_HOST_PACK", "1")
        s = Session(pm, lags, W0, H0, L0, missing=True, dtype=dtype, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
        s.train(max_iter=1, period_W=1, period_H=1, period_Lag=1)
        outs.append(s.download())
        s.close()
    for x, y in zip(*outs):
        assert np.array_equal(x, y)

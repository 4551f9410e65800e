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

"""GPU test: the on-device synthetic generator (csrc/synth.cuh) and its host twin
(bench.host_synth) produce the same matrix -- index arrays bit-exact, values bit-exact."""
import ctypes
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("T,n,n_total,col0,p", [(300, 200, 200, 0, 0.9), (257, 129, 1000, 413, 0.1), (64, 33, 33, 0, 1.0)])
def test_device_generator_matches_host_twin(dtype, T, n, n_total, col0, p):
    import bench
    from trmf.session import SynthDesc, _lib
    lib = _lib(dtype)
    sd = SynthDesc()
    assert lib.trmf_b200_synth_generate(ctypes.byref(sd), T, n, n_total, col0, 8, p, 0.01, 777, 0) == 0, \
        lib.trmf_b200_last_error().decode()
    csr, csc = bench.host_synth(T, n, n_total, col0, 8, p, 0.01, 777, dtype)
    assert int(sd.nnz) == csr.nnz

    def fetch(ptr, count, dt):
        a = np.empty(count, dtype=dt)
        assert lib.trmf_b200_copy_to_host(a.ctypes.data, ptr, a.nbytes) == 0
        return a
    assert np.array_equal(fetch(sd.d_row_ptr, T + 1, np.uint64), csr.indptr.astype(np.uint64))
    assert np.array_equal(fetch(sd.d_col_idx, csr.nnz, np.uint32), csr.indices.astype(np.uint32))
    assert np.array_equal(fetch(sd.d_val_t, csr.nnz, dtype), csr.data)
    assert np.array_equal(fetch(sd.d_col_ptr, n + 1, np.uint64), csc.indptr.astype(np.uint64))
    assert np.array_equal(fetch(sd.d_row_idx, csr.nnz, np.uint32), csc.indices.astype(np.uint32))
    assert np.array_equal(fetch(sd.d_val, csr.nnz, dtype), csc.data)
    if p >= 1.0:
        assert csr.nnz == T * n
    lib.trmf_b200_free_synth(ctypes.byref(sd))

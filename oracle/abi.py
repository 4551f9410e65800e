"""TEST INFRASTRUCTURE ONLY -- ctypes driver for a ``c_trmf_train`` shared library.

Used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` to drive the
compiled *reference* core (``oracle/_ref/trmf_float{32,64}.so``, built by
``oracle/Makefile`` from ``/root/reference/python/trmf/corelib/trmf.cpp``) on
exactly the host arrays that are handed to the CUDA library.  The product
package never imports this module.

The struct below restates the 80-byte ``PyMatrix`` POD of the reference
(``rf_matrix.h:3407-3415``; Python twin ``rf_util.py:41-52``) and the
17-argument prototype of ``c_trmf_train`` (``trmf.h:205-209``; ctypes prototype
``trmf.py:23-41``).  It is written independently of the product's own
``trmf/rf_util.py`` on purpose: the two sides of a parity test should not share
marshalling code.
"""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

DENSE_ROWMAJOR, DENSE_COLMAJOR, SPARSE = 1, 2, 3


class CMat(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("nnz", C.c_uint64),
                ("row_ptr", C.c_void_p), ("col_ptr", C.c_void_p),
                ("row_idx", C.c_void_p), ("col_idx", C.c_void_p),
                ("val", C.c_void_p), ("val_t", C.c_void_p), ("type", C.c_int32)]


assert C.sizeof(CMat) == 80


def _addr(a):
    return a.ctypes.data if a is not None else None


class HostMatrix:
    """Owns the NumPy buffers a CMat points into (keeps them alive)."""

    def __init__(self, A, dtype):
        self.bufs = {}
        m = CMat()
        m.rows, m.cols = A.shape
        if sps.issparse(A):
            csr = sps.csr_matrix(A)
            csc = sps.csc_matrix(A)
            csr.sort_indices()
            csc.sort_indices()
            b = self.bufs
            b["row_ptr"] = csr.indptr.astype(np.uint64)
            b["col_idx"] = csr.indices.astype(np.uint32)
            b["val_t"] = csr.data.astype(dtype)
            b["col_ptr"] = csc.indptr.astype(np.uint64)
            b["row_idx"] = csc.indices.astype(np.uint32)
            b["val"] = csc.data.astype(dtype)
            m.nnz = int(csr.indptr[-1])
            m.type = SPARSE
            for k, v in b.items():
                setattr(m, k, _addr(v))
        else:
            A = np.asarray(A)
            if A.flags.f_contiguous and not A.flags.c_contiguous:
                buf = np.asfortranarray(A, dtype=dtype).copy(order="F")
                m.type = DENSE_COLMAJOR
            else:
                buf = np.ascontiguousarray(A, dtype=dtype).copy(order="C")
                m.type = DENSE_ROWMAJOR
            self.bufs["val"] = buf
            m.val = _addr(buf)
            m.nnz = A.shape[0] * A.shape[1]
        self.c = m

    @property
    def array(self):
        return self.bufs["val"]


_ARGTYPES = [C.POINTER(CMat), C.c_void_p, C.c_uint32, C.POINTER(CMat), C.POINTER(CMat),
             C.POINTER(CMat), C.c_int32, C.c_double, C.c_double, C.c_double,
             C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]

_cache = {}


def load(path):
    lib = _cache.get(path)
    if lib is None:
        lib = C.CDLL(path)
        lib.c_trmf_train.restype = None
        lib.c_trmf_train.argtypes = _ARGTYPES
        _cache[path] = lib
    return lib


def ref_lib_path(dtype):
    name = "trmf_float64.so" if np.dtype(dtype) == np.float64 else "trmf_float32.so"
    return os.path.join(REF_DIR, name)


def ref_available(dtype=np.float64):
    return os.path.exists(ref_lib_path(dtype))


BIG = 1 << 30  # a period larger than any max_iter disables that phase (trmf.cpp:654,665,677)


def run_train(lib_path, Y, lag_set, W, H, lag_val, *, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1,
              max_iter=1, period_W=1, period_H=1, period_Lag=1, threads=1, missing=True,
              verbose=0, dtype=np.float64):
    """Call ``c_trmf_train`` of the library at ``lib_path`` on copies of the
    inputs; returns the updated (W, H, lag_val) as new arrays.

    Layouts follow ``trmf.cpp:583-594``: W (T x k) and H (n x k) row-major,
    lag_val (L x k) column-major, lag_set sorted uint32.
    """
    lib = load(lib_path)
    dtype = np.dtype(dtype)
    hY = Y if isinstance(Y, HostMatrix) else HostMatrix(Y, dtype)
    hW = HostMatrix(np.ascontiguousarray(W, dtype=dtype), dtype)
    hH = HostMatrix(np.ascontiguousarray(H, dtype=dtype), dtype)
    hL = HostMatrix(np.asfortranarray(lag_val, dtype=dtype), dtype)
    if hL.c.type != DENSE_COLMAJOR:  # L == 1 or k == 1: both-contiguous; the reference wants COLMAJOR
        hL.c.type = DENSE_COLMAJOR
    lags = np.ascontiguousarray(np.sort(np.asarray(lag_set)), dtype=np.uint32)
    lib.c_trmf_train(C.byref(hY.c), lags.ctypes.data, len(lags), C.byref(hW.c), C.byref(hH.c),
                     C.byref(hL.c), 1, float(lambdaI), float(lambdaAR), float(lambdaLag),
                     int(max_iter), int(period_W), int(period_H), int(period_Lag),
                     int(threads), int(bool(missing)), int(verbose))
    return hW.array, hH.array, hL.array


def run_reference(Y, lag_set, W, H, lag_val, dtype=np.float64, **kw):
    return run_train(ref_lib_path(dtype), Y, lag_set, W, H, lag_val, dtype=dtype, **kw)

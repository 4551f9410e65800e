"""Device-resident session API (additive; not part of the reference surface).

Thin ctypes binding of the ``trmf_b200_*`` symbols of ``include/trmf_b200.h``:
Y, the factors and all work vectors stay in HBM across phase calls.  Used by
``bench.py`` (kernel-resident timing, multi-GPU) and by the per-phase parity
tests; ``trmf.train`` itself goes through the drop-in ``c_trmf_train``.
"""
import ctypes
from ctypes import POINTER, byref, c_double, c_int32, c_uint32, c_uint64, c_void_p

import numpy as np

from .rf_util import PyMatrix
from .trmf import _clib

STAT = dict(cg_iters=0, accepted=1, f=2, fnew=3, gnorm=4, kernel_launches=5,
            f_ms=6, x_ms=7, lag_ms=8, f_kernel_ms=9, prered=10, actred=11, collectives=12, x_gram_ms=13, formulation=14,
            cm_gram_ms=15, cm_product_ms=16, cm_missing=17)


class SynthDesc(ctypes.Structure):
    _fields_ = [("T", c_uint64), ("n", c_uint64), ("nnz", c_uint64),
                ("d_row_ptr", c_void_p), ("d_col_idx", c_void_p), ("d_val_t", c_void_p),
                ("d_col_ptr", c_void_p), ("d_row_idx", c_void_p), ("d_val", c_void_p)]


_proto_done = set()


def _lib(dtype):
    lib = _clib.pick(dtype)
    if id(lib) in _proto_done:
        return lib
    P = POINTER(PyMatrix)
    lib.trmf_b200_create.restype = c_void_p
    lib.trmf_b200_create.argtypes = [P, POINTER(c_uint32), c_uint32, P, P, P, c_int32, c_int32]
    lib.trmf_b200_feed_mode.restype = None
    lib.trmf_b200_feed_mode.argtypes = [c_int32]
    lib.trmf_b200_create_device.restype = c_void_p
    lib.trmf_b200_create_device.argtypes = [c_uint64, c_uint64, c_uint64, c_uint32, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, POINTER(c_uint32), c_uint32,
                                            c_void_p, c_void_p, c_void_p, c_int32]
    lib.trmf_b200_destroy.restype = None
    lib.trmf_b200_destroy.argtypes = [c_void_p]
    lib.trmf_b200_set_params.argtypes = [c_void_p, c_double, c_double, c_double]
    lib.trmf_b200_set_stream.argtypes = [c_void_p, c_void_p]
    for name in ("f_update", "x_update", "lag_update", "sync", "save_factors", "restore_factors"):
        getattr(lib, "trmf_b200_" + name).argtypes = [c_void_p]
    lib.trmf_b200_train.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32]
    lib.trmf_b200_download.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.trmf_b200_upload.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.trmf_b200_stat.restype = c_double
    lib.trmf_b200_stat.argtypes = [c_void_p, c_int32]
    lib.trmf_b200_enable_timing.argtypes = [c_void_p, c_int32]
    lib.trmf_b200_nccl_unique_id.argtypes = [c_void_p]
    lib.trmf_b200_dist_init.argtypes = [c_void_p, c_int32, c_int32, c_void_p]
    lib.trmf_b200_dist_attach.argtypes = [c_void_p, c_void_p]
    lib.trmf_b200_copy_to_host.argtypes = [c_void_p, c_void_p, c_uint64]
    lib.trmf_b200_allgather_H.argtypes = [c_void_p, c_void_p, POINTER(c_uint64)]
    lib.trmf_b200_synth_generate.argtypes = [POINTER(SynthDesc), c_uint64, c_uint64, c_uint64, c_uint64, c_uint32,
                                             c_double, c_double, c_uint64, c_int32]
    lib.trmf_b200_free_synth.restype = None
    lib.trmf_b200_free_synth.argtypes = [POINTER(SynthDesc)]
    lib.trmf_b200_csr_from_csc.argtypes = [c_uint64, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_int32]
    lib.trmf_b200_roll_create.restype = c_void_p
    lib.trmf_b200_roll_create.argtypes = [P, POINTER(c_uint32), c_uint32, c_uint32, c_int32, c_int32]
    lib.trmf_b200_roll_window.argtypes = [c_void_p, c_uint64, c_void_p, c_void_p]
    lib.trmf_b200_upload_W_rows.argtypes = [c_void_p, c_uint64, c_uint64, c_void_p]
    lib.trmf_b200_download_W_rows.argtypes = [c_void_p, c_uint64, c_uint64, c_void_p]
    lib.trmf_b200_roll_nnz.restype = c_uint64
    lib.trmf_b200_roll_nnz.argtypes = [c_void_p]
    lib.trmf_b200_roll_export.argtypes = [c_void_p] * 7
    lib.trmf_b200_version.restype = ctypes.c_char_p
    lib.trmf_b200_value_bytes.restype = c_int32
    _proto_done.add(id(lib))
    return lib


def _check(lib, rc, what):
    if rc != 0:
        raise RuntimeError("trmf (CUDA) {} failed: {}".format(what, lib.trmf_b200_last_error().decode()))


class Session(object):
    """``Session(Y, lag_set, W, H, lag_val, missing=True)`` uploads everything
    once; ``f_update() / x_update() / lag_update() / train(...)`` run on the
    device; ``download()`` returns (W, H, lag_val) as NumPy arrays."""

    def __init__(self, Y, lag_set, W, H, lag_val, missing=True, dtype=None, device=0,
                 lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1):
        dtype = np.dtype(dtype if dtype is not None else W.dtype)
        self.dtype = dtype
        self.lib = lib = _lib(dtype)
        if lib.trmf_b200_device_count() <= 0:
            raise RuntimeError("trmf: no CUDA device visible; this solver is GPU-only (B200, sm_100a)")
        self.lag_set = np.ascontiguousarray(np.sort(np.asarray(lag_set)), dtype=np.uint32)
        self.pyY = Y if isinstance(Y, PyMatrix) else PyMatrix(Y, dtype)
        self.pyW = PyMatrix(np.ascontiguousarray(W), dtype, major="row")
        self.pyH = PyMatrix(np.ascontiguousarray(H), dtype, major="row")
        self.pyL = PyMatrix(np.asfortranarray(lag_val), dtype, major="col")
        self.T, self.k = self.pyW.py_buf["val"].shape
        self.n = self.pyH.py_buf["val"].shape[0]
        self.L = len(self.lag_set)
        # (self.pyY keeps Y's host buffers alive for the life of the session: the slab-wise upload of a large sparse Y may go on
        #  on the library's feeder thread while the first update is already being enqueued)
        lib.trmf_b200_feed_mode(1)
        try:
            self.h = lib.trmf_b200_create(byref(self.pyY), self.lag_set.ctypes.data_as(POINTER(c_uint32)), self.L,
                                          byref(self.pyW), byref(self.pyH), byref(self.pyL), int(bool(missing)), device)
        finally:
            lib.trmf_b200_feed_mode(0)
        if not self.h:
            raise RuntimeError("trmf (CUDA) create failed: " + lib.trmf_b200_last_error().decode())
        self.set_params(lambdaI, lambdaAR, lambdaLag)

    @classmethod
    def from_device(cls, dtype, T, n, nnz, k, row_ptr, col_idx, val_t, col_ptr, row_idx, val, lag_set,
                    d_W, d_H, d_lag_val, device=0, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1):
        """Wrap arrays already resident in HBM (integer device addresses)."""
        self = cls.__new__(cls)
        self.dtype = np.dtype(dtype)
        self.lib = lib = _lib(dtype)
        self.lag_set = np.ascontiguousarray(np.sort(np.asarray(lag_set)), dtype=np.uint32)
        self.T, self.n, self.k, self.L = int(T), int(n), int(k), len(self.lag_set)
        self.h = lib.trmf_b200_create_device(T, n, nnz, k, row_ptr, col_idx, val_t, col_ptr, row_idx, val,
                                             self.lag_set.ctypes.data_as(POINTER(c_uint32)), self.L,
                                             d_W, d_H, d_lag_val, device)
        if not self.h:
            raise RuntimeError("trmf (CUDA) create_device failed: " + lib.trmf_b200_last_error().decode())
        self.set_params(lambdaI, lambdaAR, lambdaLag)
        return self

    def set_params(self, lambdaI, lambdaAR, lambdaLag):
        _check(self.lib, self.lib.trmf_b200_set_params(self.h, lambdaI, lambdaAR, lambdaLag), "set_params")

    def set_stream(self, cuda_stream):
        _check(self.lib, self.lib.trmf_b200_set_stream(self.h, cuda_stream), "set_stream")

    def enable_timing(self, on=True):
        self.lib.trmf_b200_enable_timing(self.h, int(on))

    def f_update(self):
        _check(self.lib, self.lib.trmf_b200_f_update(self.h), "f_update")

    def x_update(self):
        _check(self.lib, self.lib.trmf_b200_x_update(self.h), "x_update")

    def lag_update(self):
        _check(self.lib, self.lib.trmf_b200_lag_update(self.h), "lag_update")

    def train(self, max_iter=10, period_W=1, period_H=1, period_Lag=2, verbose=0):
        _check(self.lib, self.lib.trmf_b200_train(self.h, max_iter, period_W, period_H, period_Lag, verbose), "train")

    def save_factors(self):
        _check(self.lib, self.lib.trmf_b200_save_factors(self.h), "save_factors")

    def restore_factors(self):
        _check(self.lib, self.lib.trmf_b200_restore_factors(self.h), "restore_factors")

    def sync(self):
        _check(self.lib, self.lib.trmf_b200_sync(self.h), "sync")

    def stat(self, name):
        return self.lib.trmf_b200_stat(self.h, STAT[name])

    def download(self):
        W = np.empty((self.T, self.k), dtype=self.dtype, order="C")
        H = np.empty((self.n, self.k), dtype=self.dtype, order="C")
        Lv = np.empty((self.L, self.k), dtype=self.dtype, order="F")
        _check(self.lib, self.lib.trmf_b200_download(self.h, W.ctypes.data, H.ctypes.data, Lv.ctypes.data), "download")
        return W, H, Lv

    def download_into(self, W=None, H=None, lag_val=None):
        """Like ``download`` but into the caller's arrays (C-ordered W / H, F-ordered lag_val of this session's
        dtype and shapes) -- e.g. the buffers of a ``trmf.Model``."""
        def ptr(a, shape, order):
            if a is None:
                return None
            assert a.dtype == self.dtype and a.shape == shape and a.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
            return a.ctypes.data
        _check(self.lib, self.lib.trmf_b200_download(self.h, ptr(W, (self.T, self.k), "C"), ptr(H, (self.n, self.k), "C"),
                                                     ptr(lag_val, (self.L, self.k), "F")), "download")

    def upload_W_rows(self, row0, rows):
        rows = np.ascontiguousarray(rows, dtype=self.dtype)
        assert rows.ndim == 2 and rows.shape[1] == self.k
        _check(self.lib, self.lib.trmf_b200_upload_W_rows(self.h, int(row0), rows.shape[0], rows.ctypes.data), "upload_W_rows")

    def download_W_rows(self, row0, nrows):
        out = np.empty((int(nrows), self.k), dtype=self.dtype, order="C")
        _check(self.lib, self.lib.trmf_b200_download_W_rows(self.h, int(row0), int(nrows), out.ctypes.data), "download_W_rows")
        return out

    def upload(self, W=None, H=None, lag_val=None):
        keep = []

        def ptr(a, order):
            if a is None:
                return None
            a = np.asarray(a, dtype=self.dtype, order=order)
            keep.append(a)
            return a.ctypes.data
        _check(self.lib, self.lib.trmf_b200_upload(self.h, ptr(W, "C"), ptr(H, "C"), ptr(lag_val, "F")), "upload")

    def close(self):
        if getattr(self, "h", None):
            self.lib.trmf_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RollingSession(Session):
    """Y resident in HBM across the windows of ``rolling_validate`` (reference trmf.py:303-329).

    ``RollingSession(Y, lag_set, k, missing)`` uploads Y once -- every time stamp any window trains on; a scipy
    sparse matrix (one orientation crosses PCIe, the other is derived on the device) or a dense array (made
    C-contiguous: a prefix of the time axis must be contiguous).  A dense array with ``missing=True`` is
    sparsified on the device: its non-zero cells are the observations, exactly the ``csr_matrix(Y_trn)`` of the
    reference's loop (trmf.py:320-321) without the host conversion.  ``window(T_w, scale, offset)`` then makes all
    ``Session`` calls operate on ``Y[:T_w]`` -- optionally through the per-series affine map of the reference's
    ``NormalizedTransform.preprocess`` -- without touching the host copy of Y again; factors of the previous
    window stay in place and ``upload_W_rows`` appends the warm-start rows of W."""

    def __init__(self, Y, lag_set, k, missing=True, dtype=None, device=0, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1):
        dtype = np.dtype(dtype if dtype is not None else Y.dtype)
        self.dtype = dtype
        self.lib = lib = _lib(dtype)
        if lib.trmf_b200_device_count() <= 0:
            raise RuntimeError("trmf: no CUDA device visible; this solver is GPU-only (B200, sm_100a)")
        self.lag_set = np.ascontiguousarray(np.sort(np.asarray(lag_set)), dtype=np.uint32)
        if isinstance(Y, PyMatrix):
            self.pyY = Y
        elif isinstance(Y, np.ndarray):
            self.pyY = PyMatrix(np.ascontiguousarray(Y), dtype, major="row", copy=False)
        else:
            self.pyY = PyMatrix(Y, dtype, twin=False)
        self.T_cap = self.T = int(self.pyY.rows)
        self.n = int(self.pyY.cols)
        self.k = int(k)
        self.L = len(self.lag_set)
        self.h = lib.trmf_b200_roll_create(byref(self.pyY), self.lag_set.ctypes.data_as(POINTER(c_uint32)), self.L,
                                           self.k, int(bool(missing)), device)
        if not self.h:
            raise RuntimeError("trmf (CUDA) roll_create failed: " + lib.trmf_b200_last_error().decode())
        self.dense_resident = self.pyY.type in (PyMatrix.DENSE_ROWMAJOR, PyMatrix.DENSE_COLMAJOR) and not missing
        self.pyY = None   # the device holds Y now; let the host marshalling copies go
        self.set_params(lambdaI, lambdaAR, lambdaLag)

    def window(self, T_w, scale=None, offset=None):
        """Train on ``Y[:T_w]`` from now on; ``scale`` / ``offset``: n values each, ``y -> y*scale[j] + offset[j]``."""
        keep = []

        def ptr(a):
            if a is None:
                return None
            a = np.ascontiguousarray(np.asarray(a).reshape(-1), dtype=self.dtype)
            assert a.shape[0] == self.n
            keep.append(a)
            return a.ctypes.data
        _check(self.lib, self.lib.trmf_b200_roll_window(self.h, int(T_w), ptr(scale), ptr(offset)), "roll_window")
        self.T = int(T_w)

    def window_stats(self, T_w):
        """(mean, std) per series over ``Y[:T_w]`` of a dense resident Y, computed on the device in NumPy's axis-0 summation
        order -- bit-identical to ``Y[:T_w].mean(axis=0)``, ``Y[:T_w].std(axis=0)`` (reference trmf.py:84-88).  Returns None
        when the session does not hold a dense copy of Y (sparse-born or sparsified sessions)."""
        if not self.dense_resident:
            return None
        mean = np.empty(self.n, dtype=self.dtype)
        std = np.empty(self.n, dtype=self.dtype)
        self.lib.trmf_b200_roll_stats.argtypes = [c_void_p, c_uint64, c_void_p, c_void_p]
        _check(self.lib, self.lib.trmf_b200_roll_stats(self.h, int(T_w), mean.ctypes.data, std.ctypes.data), "roll_stats")
        return mean, std

    @property
    def nnz(self):
        return int(self.lib.trmf_b200_roll_nnz(self.h))

    def export_window(self):
        """(row_ptr, col_idx, val_t, col_ptr, row_idx, val) of the current window, as the device holds them."""
        nz = self.nnz
        row_ptr = np.empty(self.T + 1, dtype=np.uint64)
        col_ptr = np.empty(self.n + 1, dtype=np.uint64)
        col_idx = np.empty(nz, dtype=np.uint32)
        row_idx = np.empty(nz, dtype=np.uint32)
        val_t = np.empty(nz, dtype=self.dtype)
        val = np.empty(nz, dtype=self.dtype)
        _check(self.lib, self.lib.trmf_b200_roll_export(self.h, row_ptr.ctypes.data, col_idx.ctypes.data, val_t.ctypes.data,
                                                        col_ptr.ctypes.data, row_idx.ctypes.data, val.ctypes.data), "roll_export")
        return row_ptr, col_idx, val_t, col_ptr, row_idx, val


def csr_from_csc(csc, dtype=None, device=0):
    """By-time CSR of a scipy CSC matrix through the device transpose (`trmf_b200_csr_from_csc`):
    returns (row_ptr uint64[T+1], col_idx uint32[nnz], val_t dtype[nnz]) -- the arrays
    PyMatrix would otherwise get from a second scipy conversion (reference rf_util.py:88-98)."""
    dtype = np.dtype(dtype or csc.dtype)
    lib = _lib(dtype)
    T, n = csc.shape
    col_ptr = np.ascontiguousarray(csc.indptr, dtype=np.uint64)
    row_idx = np.ascontiguousarray(csc.indices, dtype=np.uint32)
    val = np.ascontiguousarray(csc.data, dtype=dtype)
    nnz = int(col_ptr[-1])
    row_ptr = np.empty(T + 1, dtype=np.uint64)
    col_idx = np.empty(nnz, dtype=np.uint32)
    val_t = np.empty(nnz, dtype=dtype)
    rc = lib.trmf_b200_csr_from_csc(T, n, nnz, col_ptr.ctypes.data, row_idx.ctypes.data, val.ctypes.data,
                                    row_ptr.ctypes.data, col_idx.ctypes.data, val_t.ctypes.data, device)
    if rc != 0:
        raise RuntimeError(lib.trmf_b200_last_error().decode())
    return row_ptr, col_idx, val_t

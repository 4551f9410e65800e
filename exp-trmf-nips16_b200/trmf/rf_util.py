"""ctypes boundary of the B200 TRMF solver: the ``PyMatrix`` struct and the
library loader.

Mirrors the *interface* of the reference's ``python/trmf/rf_util.py`` (the
80-byte ``PyMatrix`` POD at rf_util.py:35-52 == rf_matrix.h:3399-3415, the
twin CSR+CSC buffers with uint64 pointers / uint32 indices at rf_util.py:88-98,
``load_dynamic_library`` at rf_util.py:19-32) on NumPy >= 2 / SciPy >= 1.14.

The libraries loaded here are CUDA-only.  If they are missing they are built
with nvcc (``make -C corelib lib``); if that fails, or no B200 is visible at
call time, the failure is loud -- there is no CPU implementation to fall back
to.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np
import scipy.sparse as sps

__all__ = ["PyMatrix", "fillprototype", "load_dynamic_library", "pack_bitmap"]


def fillprototype(fn, restype, argtypes):
    fn.restype = restype
    fn.argtypes = argtypes


def load_dynamic_library(dirname, soname, forced_rebuild=False):
    """``glob(dirname/soname*.so)[0]`` -> CDLL, running ``make -C dirname lib``
    first when asked to or when nothing is there (rf_util.py:19-32)."""
    def find():
        hits = sorted(glob.glob(os.path.join(dirname, soname) + "*.so"))
        return hits[0] if hits else None

    so = None if forced_rebuild else find()
    if so is None:
        proc = subprocess.run(["make", "-C", dirname, "lib"], capture_output=True, text=True)
        so = find()
        if so is None:
            raise RuntimeError(
                "{} CUDA library cannot be found and could not be built with nvcc:\n{}".format(
                    soname, proc.stdout[-2000:] + proc.stderr[-2000:]))
    return ctypes.CDLL(so)


def pack_bitmap(col_ptr, row_idx, rows):
    """Row indices of a CSC matrix as one bitmap per column: uint32[cols * ceil(rows / 32)], bit (i & 31) of word
    [j][i >> 5] set iff (i, j) is stored (``TRMF_SPARSE_BITMAP`` of include/trmf_b200.h).  Indices must be sorted
    within every column (they are, in the canonical CSC this module builds)."""
    col_ptr = np.asarray(col_ptr).astype(np.int64)
    row_idx = np.asarray(row_idx)
    cols = len(col_ptr) - 1
    words = (int(rows) + 31) // 32
    out = np.zeros(cols * words, dtype=np.uint32)
    if len(row_idx):
        col_of = np.repeat(np.arange(cols, dtype=np.int64), np.diff(col_ptr))
        widx = col_of * words + (row_idx.astype(np.int64) >> 5)
        bit = (np.uint32(1) << (row_idx.astype(np.uint32) & np.uint32(31))).astype(np.uint64)
        starts = np.concatenate(([0], np.flatnonzero(widx[1:] != widx[:-1]) + 1))
        out[widx[starts]] = np.add.reduceat(bit, starts).astype(np.uint32)   # distinct bits of one word: sum == or
    return out


class PyMatrix(ctypes.Structure):
    """Zero-copy view descriptor handed to ``c_trmf_train``.

    sparse: ``row_ptr/col_idx/val_t`` is the CSR and ``col_ptr/row_idx/val`` the
    CSC of the same matrix; dense: only ``val`` (row- or column-major, see
    ``type``).  The NumPy arrays are kept alive in ``py_buf``; ``py_buf['val']``
    of a dense matrix is the array the solver updates in place.
    """
    DENSE_ROWMAJOR = 1
    DENSE_COLMAJOR = 2
    SPARSE = 3
    EYE = 4
    SPARSE_BITMAP = 5   # additive: CSC half only, row indices as one bitmap per column (see pack_bitmap)

    _fields_ = [
        ("rows", ctypes.c_uint64),
        ("cols", ctypes.c_uint64),
        ("nnz", ctypes.c_uint64),
        ("row_ptr", ctypes.POINTER(ctypes.c_uint64)),
        ("col_ptr", ctypes.POINTER(ctypes.c_uint64)),
        ("row_idx", ctypes.POINTER(ctypes.c_uint32)),
        ("col_idx", ctypes.POINTER(ctypes.c_uint32)),
        ("val", ctypes.c_void_p),
        ("val_t", ctypes.c_void_p),
        ("type", ctypes.c_int32),
    ]

    def __init__(self, A, dtype=np.float32, major=None, twin=True, copy=True, pack=False):
        """``major`` ('row' / 'col', optional, additive to the reference signature)
        settles the type tag of arrays that are both C- and F-contiguous (a
        single row or column, e.g. W with k == 1), which the reference would tag
        COLMAJOR and then reject for W/H.

        ``twin`` (additive): the reference always builds BOTH sparse orientations
        on the host (two scipy conversions, rf_util.py:88-98).  ``twin=False``
        keeps only the orientation ``A`` already has (CSC for a csc_matrix, CSR
        otherwise) and leaves the other three pointers NULL: the CUDA library
        derives the missing half on the device, bit-identically
        (``csrc/ingest.cuh``).  ``trmf.train`` uses this; such a PyMatrix is not
        valid input for the reference's own core.

        ``pack`` (additive, sparse only): keep the CSC half only and carry its row
        indices as one bitmap per column (``type = SPARSE_BITMAP``): 4 bytes of
        values + ~1/8 byte of mask per observed cell of a mostly-observed panel
        cross PCIe instead of 8; the library expands the bitmap to the identical
        ``row_idx`` array on the device.  Only this package's CUDA library
        understands it.

        ``copy`` (additive): the reference always copies a dense array
        (``A.astype(dtype)``, rf_util.py:122).  ``copy=False`` adopts ``A`` itself
        when it already has the dtype and is contiguous -- for read-only inputs
        such as a large Y."""
        super().__init__()
        if A is None:
            return
        dtype = np.dtype(dtype)
        self.dtype = dtype
        self.py_buf = buf = {}
        self.rows, self.cols = int(A.shape[0]), int(A.shape[1])
        if sps.issparse(A):
            if A.format == "coo":
                # explicit COO: duplicate (i, j) entries stay separate observations, as in the reference (rf_util.py:100-119:
                # bincount + argsort(row * ncols + col), nnz = len(data)); scipy's own conversions would sum them
                self._init_from_coo(A, dtype, twin, pack)
                self._bind()
                return
            self.type = PyMatrix.SPARSE
            want_csr = (twin or A.format != "csc") and not pack
            want_csc = twin or A.format == "csc" or pack
            if want_csr:
                csr = sps.csr_matrix(A)
                if not csr.has_sorted_indices:
                    csr = csr.sorted_indices()
                self.nnz = int(csr.indptr[-1])
                buf["row_ptr"] = csr.indptr.astype(np.uint64)
                buf["col_idx"] = csr.indices.astype(np.uint32)
                buf["val_t"] = csr.data.astype(dtype, copy=False)
            if want_csc:
                csc = sps.csc_matrix(A)
                if not csc.has_sorted_indices:
                    csc = csc.sorted_indices()
                self.nnz = int(csc.indptr[-1])
                buf["col_ptr"] = csc.indptr.astype(np.uint64)
                buf["row_idx"] = csc.indices.astype(np.uint32)
                buf["val"] = csc.data.astype(dtype, copy=False)
                if pack:
                    buf["row_idx"] = pack_bitmap(buf["col_ptr"], buf["row_idx"], self.rows)
                    self.type = PyMatrix.SPARSE_BITMAP
        elif isinstance(A, np.ndarray):
            arr = A.astype(dtype, copy=copy)   # order='K': keeps the caller's memory layout
            if not (arr.flags.c_contiguous or arr.flags.f_contiguous):
                arr = np.ascontiguousarray(arr)
            buf["val"] = arr
            # same precedence as the reference (rf_util.py:123-126): F-contiguous wins
            self.type = PyMatrix.DENSE_COLMAJOR if arr.flags.f_contiguous else PyMatrix.DENSE_ROWMAJOR
            if major is not None and arr.flags.c_contiguous and arr.flags.f_contiguous:
                self.type = PyMatrix.DENSE_ROWMAJOR if major == "row" else PyMatrix.DENSE_COLMAJOR
            self.nnz = self.rows * self.cols
        else:
            raise TypeError("PyMatrix expects a numpy.ndarray or a scipy.sparse matrix, got {}".format(type(A)))
        self._bind()

    def _bind(self):
        fields = dict(PyMatrix._fields_)
        for name, arr in self.py_buf.items():
            ctype = fields[name]
            if ctype is ctypes.c_void_p:
                setattr(self, name, arr.ctypes.data)
            else:
                setattr(self, name, arr.ctypes.data_as(ctype))

    def _init_from_coo(self, coo, dtype, twin, pack):
        def compress(major, minor, nmajor, nminor):
            indptr = np.cumsum(np.bincount(major.astype(np.int64) + 1, minlength=nmajor + 1), dtype=np.uint64)
            order = np.argsort(major.astype(np.int64) * nminor + minor.astype(np.int64), kind="stable")
            return indptr, minor[order].astype(np.uint32), coo.data[order].astype(dtype)
        buf = self.py_buf
        self.type = PyMatrix.SPARSE
        self.nnz = int(coo.data.shape[0])
        dup = self.nnz != len(np.unique(coo.row.astype(np.int64) * self.cols + coo.col.astype(np.int64)))
        if pack and dup:
            raise ValueError("pack=True cannot represent duplicate (row, col) entries of a COO matrix")
        if twin or not pack:     # CSR (the orientation a COO matrix is closest to); with twin also the CSC
            buf["row_ptr"], buf["col_idx"], buf["val_t"] = compress(coo.row, coo.col, self.rows, self.cols)
        if twin or pack:
            buf["col_ptr"], buf["row_idx"], buf["val"] = compress(coo.col, coo.row, self.cols, self.rows)
        if pack:
            for name in ("row_ptr", "col_idx", "val_t"):
                buf.pop(name, None)
            buf["row_idx"] = pack_bitmap(buf["col_ptr"], buf["row_idx"], self.rows)
            self.type = PyMatrix.SPARSE_BITMAP

    @classmethod
    def identity(cls, size, dtype=np.float32):
        eye = cls(None)
        eye.rows = eye.cols = eye.nnz = int(size)
        eye.dtype = np.dtype(dtype)
        eye.py_buf = {}
        eye.type = PyMatrix.EYE
        return eye

"""Multi-GPU host logic (additive; the reference is single-process).

One process per GPU.  The series axis of Y (its columns) is cut into contiguous
slabs balanced by the number of observed entries; rank r keeps Y[:, slab_r] in
both orientations plus the matching rows of H (the F factor), while W (the X
factor), lag_val and every CG vector are replicated.  The F-update is then
embarrassingly parallel; each Omega-proportional pass of the X-update produces a
T x k partial that the CUDA library sums with one ncclAllReduce over NVLink
(``csrc/extras.cuh``), after which all ranks take identical CG decisions.

``torch.distributed`` is used only as plumbing (to hand rank 0's ncclUniqueId to
the other ranks); with the ``gloo`` backend the same code path runs on CPUs for
the partition logic (tests/test_dist_cpu.py).
"""
import ctypes

import numpy as np
import scipy.sparse as sps

__all__ = ["slab_bounds", "slice_slab", "DistSession"]


def slab_bounds(col_ptr, world):
    """Contiguous series slabs with (nearly) equal numbers of observed entries.

    ``col_ptr``: CSC pointer array of Y (length n+1).  Returns ``world + 1``
    monotone boundaries ``b`` with ``b[0] = 0`` and ``b[world] = n``; slab r is
    ``[b[r], b[r+1])``.  Boundary r is the first series whose prefix count
    reaches r/world of the total, so no slab exceeds its fair share by more than
    one series' worth of entries.  Slabs may be empty when n < world."""
    col_ptr = np.asarray(col_ptr, dtype=np.int64)
    n = len(col_ptr) - 1
    total = int(col_ptr[-1])
    bounds = [0]
    for r in range(1, world):
        if total == 0:
            b = (n * r) // world
        else:
            b = int(np.searchsorted(col_ptr, (total * r) // world, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return bounds


def slice_slab(Y, lo, hi):
    """Y[:, lo:hi] as CSR (column indices local to the slab)."""
    return sps.csr_matrix(sps.csc_matrix(Y)[:, lo:hi])


class DistSession(object):
    """Rank-local view of a sharded training run.

    ``Y`` is either the full T x n matrix (every rank slices its own slab) or
    already the rank's slab (``bounds`` given).  ``H`` likewise.  After
    ``train`` / the phase calls, ``W`` and ``lag_val`` are identical on all
    ranks and ``gather_H`` returns the full series factor on every rank.
    """

    def __init__(self, Y, lag_set, W, H, lag_val, rank, world, device=None, bounds=None, dtype=None,
                 lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1, process_group=None):
        import torch
        import torch.distributed as dist
        from .session import Session, _lib
        self.rank, self.world = rank, world
        if bounds is None:
            bounds = slab_bounds(sps.csc_matrix(Y).indptr, world)
            Y = slice_slab(Y, bounds[rank], bounds[rank + 1])
            H = np.ascontiguousarray(H[bounds[rank]:bounds[rank + 1]])
        self.bounds = list(bounds)
        device = rank if device is None else device
        self.device = device
        self.session = Session(Y, lag_set, W, H, lag_val, missing=True, dtype=dtype, device=device,
                               lambdaI=lambdaI, lambdaAR=lambdaAR, lambdaLag=lambdaLag)
        lib = self.session.lib
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (ctypes.c_ubyte * 128)()
                if lib.trmf_b200_nccl_unique_id(buf) != 0:
                    raise RuntimeError(lib.trmf_b200_last_error().decode())
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            if dist.get_backend(process_group) == "nccl":
                uid = uid.cuda(device)
            dist.broadcast(uid, 0, group=process_group)
            idb = (ctypes.c_ubyte * 128)(*uid.cpu().tolist())
            if lib.trmf_b200_dist_init(self.session.h, rank, world, idb) != 0:
                raise RuntimeError(lib.trmf_b200_last_error().decode())

    def __getattr__(self, name):   # f_update / x_update / lag_update / train / stat / download ...
        return getattr(self.session, name)

    def gather_H(self):
        """Full n x k series factor, assembled with NCCL broadcasts of the slabs."""
        import torch
        s = self.session
        counts = (ctypes.c_uint64 * self.world)(*[self.bounds[r + 1] - self.bounds[r] for r in range(self.world)])
        n = self.bounds[-1]
        tdt = torch.float64 if s.dtype == np.float64 else torch.float32
        full = torch.empty((n, s.k), dtype=tdt, device=torch.device("cuda", self.device))
        torch.cuda.synchronize(self.device)
        if s.lib.trmf_b200_allgather_H(s.h, full.data_ptr(), counts) != 0:
            raise RuntimeError(s.lib.trmf_b200_last_error().decode())
        s.sync()
        return full.cpu().numpy()

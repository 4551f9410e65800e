"""CPU tests of the multi-GPU host logic (gloo, world_size 2): the series-slab partition
and the claim the sharded solver rests on -- F rows are independent given X, and every
Omega-proportional X-update quantity is a sum over slabs -- checked by running the oracle
sharded over two processes with all_reduce and comparing with the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sps

import cases
from oracle import trmf_numpy as tn
from trmf.dist import slab_bounds, slice_slab


def test_slab_bounds_balance_and_cover():
    rng = np.random.RandomState(0)
    counts = rng.randint(0, 50, size=1000)
    counts[100:140] = 0            # a run of empty series
    counts[500] = 5000             # one very heavy series
    ptr = np.concatenate([[0], np.cumsum(counts)])
    for world in (1, 2, 3, 4, 8):
        b = slab_bounds(ptr, world)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == world + 1
        assert all(b[i] <= b[i + 1] for i in range(world))
        loads = [int(ptr[b[i + 1]] - ptr[b[i]]) for i in range(world)]
        assert sum(loads) == int(ptr[-1])
        assert max(loads) <= ptr[-1] / world + counts.max()
    assert slab_bounds(np.zeros(6, dtype=np.int64), 4) == [0, 1, 2, 3, 5]   # empty Y: split series evenly
    assert slab_bounds(np.array([0, 3]), 4)[-1] == 1                         # fewer series than ranks


def test_slice_slab_keeps_entries():
    p = cases.make_problem(40, 30, 3, [1, 2], 0.5, seed=2)
    b = slab_bounds(sps.csc_matrix(p["Ysp"]).indptr, 3)
    parts = [slice_slab(p["Ysp"], b[r], b[r + 1]) for r in range(3)]
    assert sum(x.nnz for x in parts) == p["Ysp"].nnz
    assert np.array_equal(sps.hstack(parts).toarray(), p["Ysp"].toarray())


class ShardedLoss:
    """SparseLoss over this rank's slab; partial sums combined with all_reduce (what
    csrc/extras.cuh does with ncclAllReduce on the T x k partials)."""

    def __init__(self, Yslab, Hslab, dist, torch):
        self.local = tn.SparseLoss(Yslab, Hslab)
        self.dist, self.torch = dist, torch

    def _sum(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        self.dist.all_reduce(t)
        return t.numpy()

    def fun(self, W):
        return float(self._sum(np.array([self.local.fun(W)]))[0])

    def grad(self, W):
        return self._sum(self.local.grad(W))

    def Hv(self, S):
        return self._sum(self.local.Hv(S))


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = cases.make_problem(90, 70, 6, [1, 3, 8], 0.55, seed=31)
    lam = (0.5, 8.0, 0.5)
    b = slab_bounds(sps.csc_matrix(p["Ysp"]).indptr, world)
    Ys = slice_slab(p["Ysp"], b[rank], b[rank + 1])
    W, L = p["W0"].copy(), p["L0"].copy()
    H = p["H0"][b[rank]:b[rank + 1]].copy()
    lags = p["lags"].astype(np.int64)
    cg = []
    for it in range(2):
        H = tn.f_update_sparse(sps.csc_matrix(Ys), W, H, lam[0])            # local: F rows are independent
        info = {}
        W = tn.x_update(ShardedLoss(Ys, H, dist, torch), W, lags, L, lam[0], lam[1], info)
        cg.append(info["cg_iter"])
        L = tn.lag_update(W, lags, lam[2])                                   # replicated
    parts = [None] * world
    dist.all_gather_object(parts, H)
    if rank == 0:
        np.savez(out_path, W=W, H=np.vstack(parts), L=L, cg=np.array(cg))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_oracle_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    p = cases.make_problem(90, 70, 6, [1, 3, 8], 0.55, seed=31)
    tr = []
    W, H, L = tn.train(p["Ysp"], p["lags"], p["W0"], p["H0"], p["L0"], 0.5, 8.0, 0.5, max_iter=2, period_Lag=1,
                       missing=True, trace=tr)
    assert [t["cg_iter"] for t in tr] == list(z["cg"])
    assert cases.rel(z["W"], W) < 1e-12 and cases.rel(z["H"], H) < 1e-12 and cases.rel(z["L"], L) < 1e-12

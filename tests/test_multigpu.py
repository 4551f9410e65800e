"""GPU test (needs >= 2 GPUs; skipped on a 1-GPU box): series-slab sharding over NCCL gives the
same factors as one GPU and as the float64 oracle."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sps

import cases
from oracle import trmf_numpy as tn

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import trmf
        return trmf.trmf._clib.clib_float32.trmf_b200_device_count()
    except Exception:
        return 0


def _worker(rank, world, port, out_path, dtype_name):
    import torch
    import torch.distributed as dist
    from trmf.dist import DistSession
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dtype = np.dtype(dtype_name)
    p = cases.make_problem(600, 500, 40, [1, 7, 24], 0.6, seed=77)
    Y = sps.csr_matrix((p["Ysp"].data.astype(dtype), p["Ysp"].indices, p["Ysp"].indptr), shape=p["Ysp"].shape)
    ds = DistSession(Y, p["lags"], p["W0"].astype(dtype), p["H0"].astype(dtype), p["L0"].astype(dtype), rank, world,
                     dtype=dtype, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
    cg = []
    for it in range(2):
        ds.f_update(); ds.x_update(); ds.lag_update()
        cg.append(int(ds.stat("cg_iters")))
    W, _, L = ds.download()
    H = ds.gather_H()
    ncoll = ds.stat("collectives")
    Ws = [None] * world
    dist.all_gather_object(Ws, W)
    if rank == 0:
        np.savez(out_path, W=W, H=H, L=L, cg=np.array(cg), ncoll=ncoll, replicated=all(np.array_equal(Ws[0], w) for w in Ws))
    dist.barrier()
    ds.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-5)])
def test_two_gpu_slab_sharding_matches_oracle(tmp_path, dtype, tol):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "mg.npz")
    mp.spawn(_worker, args=(2, port, out, dtype), nprocs=2, join=True)
    z = np.load(out)
    p = cases.make_problem(600, 500, 40, [1, 7, 24], 0.6, seed=77)
    dt = np.dtype(dtype)
    Y = sps.csr_matrix((p["Ysp"].data.astype(dt).astype(np.float64), p["Ysp"].indices, p["Ysp"].indptr), shape=p["Ysp"].shape)
    tr = []
    c = lambda a: a.astype(dt).astype(np.float64)
    W, H, L = tn.train(Y, p["lags"], c(p["W0"]), c(p["H0"]), c(p["L0"]), 0.5, 50.0, 0.5, max_iter=2, period_Lag=1,
                       missing=True, trace=tr)
    assert bool(z["replicated"])                      # X identical on both ranks (bitwise)
    assert float(z["ncoll"]) > 0                      # NCCL all-reduces really ran
    assert [t["cg_iter"] for t in tr] == list(z["cg"])
    scale = 3.0 if dtype == "float32" else 1.0        # two compounded iterations in fp32
    assert cases.rel(z["W"], W) < tol * scale and cases.rel(z["H"], H) < tol * scale and cases.rel(z["L"], L) < tol * scale

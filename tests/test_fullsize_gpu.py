"""GPU tests at BASELINE.json's full single-GPU size (configs[1]: T = n = 10 000, k = 40, 10 % missing,
lag_set {1,7,24}; Y generated in HBM) through checks that do not need a full CPU run:
F rows are independent given X, so a random sample of them is recomputed exactly on the host;
the objective reported by the X-update is recomputed on the host in float64; the F-update is
idempotent bitwise; the objective decreases over outer iterations."""
import ctypes
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_c2_full_size_properties():
    import torch
    import bench
    from oracle import trmf_numpy as tn
    from trmf.session import Session, SynthDesc, _lib
    cfg = bench.CONFIGS["c2"]
    T, n, k, lags = cfg["T"], cfg["n"], cfg["k"], np.array(cfg["lags"], dtype=np.uint32)
    dtype = np.float32
    lib = _lib(dtype)
    sd = SynthDesc()
    assert lib.trmf_b200_synth_generate(ctypes.byref(sd), T, n, n, 0, bench.RANK_TRUE, cfg["p"], bench.NOISE, bench.SEED, 0) == 0
    nnz = int(sd.nnz)
    assert abs(nnz / (T * n) - cfg["p"]) < 1e-3
    W0, H0, L0 = bench.init_factors(T, n, k, len(lags), dtype)
    dev = torch.device("cuda", 0)
    dW, dH, dL = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (W0, H0, np.ascontiguousarray(L0.T)))
    lam = bench.LAMBDAS
    s = Session.from_device(dtype, T, n, nnz, k, sd.d_row_ptr, sd.d_col_idx, sd.d_val_t, sd.d_col_ptr, sd.d_row_idx, sd.d_val,
                            lags, dW.data_ptr(), dH.data_ptr(), dL.data_ptr(), device=0, lambdaI=lam[0], lambdaAR=lam[1], lambdaLag=lam[2])

    def fetch(ptr, count, dt):
        a = np.empty(count, dtype=dt)
        assert lib.trmf_b200_copy_to_host(a.ctypes.data, ptr, a.nbytes) == 0
        return a
    col_ptr = fetch(sd.d_col_ptr, n + 1, np.uint64).astype(np.int64)
    row_idx = fetch(sd.d_row_idx, nnz, np.uint32)
    val = fetch(sd.d_val, nnz, dtype)

    # (1) F-update: a random sample of series recomputed exactly (float64) from the same X
    s.f_update()
    _, H1, _ = s.download()
    rng = np.random.RandomState(0)
    W64 = W0.astype(np.float64)
    worst = 0.0
    for j in rng.choice(n, 24, replace=False):
        lo, hi = col_ptr[j], col_ptr[j + 1]
        Wj = W64[row_idx[lo:hi]]
        h = np.linalg.solve(Wj.T @ Wj + lam[0] * np.eye(k), Wj.T @ val[lo:hi].astype(np.float64))
        worst = max(worst, np.linalg.norm(H1[j] - h) / np.linalg.norm(h))
    assert worst < 1e-5, worst

    # (2) idempotent, bitwise: the kernels have a fixed summation order
    s.f_update()
    _, H2, _ = s.download()
    assert np.array_equal(H1, H2)

    # (3) the objective the X-update reports at its starting point, recomputed on the host in float64
    s.x_update()
    f_gpu, fnew_gpu = s.stat("f"), s.stat("fnew")
    assert s.stat("accepted") == 1.0 and 1 <= s.stat("cg_iters") <= 20 and fnew_gpu < f_gpu
    H64 = H1.astype(np.float64)
    loss = 0.0
    for j0 in range(0, n, 500):
        j1 = min(n, j0 + 500)
        lo, hi = col_ptr[j0], col_ptr[j1]
        cols = np.repeat(np.arange(j0, j1), np.diff(col_ptr[j0:j1 + 1]))
        r = val[lo:hi].astype(np.float64) - np.einsum("ek,ek->e", W64[row_idx[lo:hi]], H64[cols])
        loss += float(r @ r)
    f_host = 0.5 * loss + tn.base_fun(W64, lags.astype(np.int64), L0.astype(np.float64), lam[0], lam[1])
    assert abs(f_gpu - f_host) <= 1e-6 * abs(f_host), (f_gpu, f_host)

    # (4) ALS makes progress: the objective at the start of the next X-update is lower again
    s.lag_update(); s.f_update(); s.x_update()
    assert s.stat("f") < fnew_gpu * (1 + 1e-6)
    s.close()
    lib.trmf_b200_free_synth(ctypes.byref(sd))

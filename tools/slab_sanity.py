"""Ten-second check of the host-buffer path on a GPU box (no torch, no oracle run): c_trmf_train on a mostly observed Y with more
than 2^22 entries -- slab-wise upload, feeder thread, host-packed bitmaps, slab-wise complement F-update -- must return the factors
of a device-resident session bit for bit.    python tools/slab_sanity.py"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "exp-trmf-nips16_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cases  # noqa: E402
from oracle import abi  # noqa: E402  (only its ctypes caller of c_trmf_train is used here)
from trmf.session import Session  # noqa: E402

f32 = lambda a: np.asarray(a, dtype=np.float32)  # noqa: E731
p = cases.make_problem(2600, 2000, 40, [1, 7, 24], 0.9, seed=4)
Y = sps.csr_matrix((f32(p["Ysp"].data), p["Ysp"].indices, p["Ysp"].indptr), shape=p["Ysp"].shape)
assert Y.nnz >= 1 << 22
lib = os.path.join(ROOT, "exp-trmf-nips16_b200", "trmf", "corelib", "trmf_float32.so")
kw = dict(lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5, max_iter=2, period_Lag=1, missing=True)
t0 = time.perf_counter()
W, H, L = abi.run_train(lib, Y, p["lags"], f32(p["W0"]), f32(p["H0"]), f32(p["L0"]), dtype=np.float32, **kw)
t1 = time.perf_counter()
s = Session(Y, p["lags"], f32(p["W0"]), f32(p["H0"]), f32(p["L0"]), missing=True, dtype=np.float32, lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
s.train(max_iter=2, period_W=1, period_H=1, period_Lag=1)
W2, H2, L2 = s.download()
form = s.stat("formulation")
s.close()
same = np.array_equal(W, W2) and np.array_equal(H, H2) and np.array_equal(L, L2)
print("c_trmf_train %.1f ms (first call: context creation included); formulation %d; bit-identical to the resident session: %s; finite: %s"
      % (1e3 * (t1 - t0), form, same, bool(np.isfinite(W).all() and np.isfinite(H).all())))
sys.exit(0 if same else 1)

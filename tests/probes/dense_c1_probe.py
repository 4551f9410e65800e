"""C1-shaped dense run (electricity shape: T=26304, n=370, k=20, lags 1..24, missing=False) through
trmf.train on the GPU next to the compiled reference on the host cores; prints times and parity."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "exp-trmf-nips16_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trmf
from oracle import abi
import cases

def run(T, n, k, lags, iters, dtype):
    rng = np.random.RandomState(0)
    t = np.arange(T)[:, None]
    Y = (5 + rng.rand(1, n) * 20) * (1 + 0.5 * np.sin(2 * np.pi * t / 24 + rng.rand(1, n) * 6) + 0.2 * np.sin(2 * np.pi * t / 168)) + rng.randn(T, n)
    Y = Y.astype(dtype)
    m = trmf.Model.initialize(Y, lags, k, seed=0)
    W0, H0, L0 = m.W.copy(), m.H.copy(), m.lag_val.copy()
    kw = dict(lambdaI=0.5, lambdaAR=125.0, lambdaLag=2.0, max_iter=iters, period_W=1, period_H=1, period_Lag=2, missing=False)
    trmf.train(Y, m, **{k_: v for k_, v in kw.items() if k_ not in ()})   # warm-up (module load, pool)
    m = trmf.Model.initialize(Y, lags, k, seed=0)
    t0 = time.perf_counter(); trmf.train(Y, m, **kw); tg = time.perf_counter() - t0
    out = "T={} n={} k={} L={} iters={} {}: GPU (host buffers in/out) {:.1f} ms".format(T, n, k, len(lags), iters, np.dtype(dtype).name, 1e3 * tg)
    if abi.ref_available(dtype):
        t0 = time.perf_counter()
        Wr, Hr, Lr = abi.run_reference(Y, np.array(lags), W0, H0, L0, dtype=dtype, threads=os.cpu_count(), **kw)
        tc = time.perf_counter() - t0
        out += "; reference on {} cores {:.1f} ms; speed-up {:.1f}x; rel diff W {:.1e} H {:.1e} lag {:.1e}".format(
            os.cpu_count(), 1e3 * tc, tc / tg, cases.rel(m.W, Wr), cases.rel(m.H, Hr), cases.rel(m.lag_val, Lr))
    print(out, flush=True)

run(26304, 370, 20, list(range(1, 25)), 4, np.float32)
run(26304, 370, 20, list(range(1, 25)), 4, np.float64)
run(10560, 963, 40, list(range(1, 25)) + [168, 336], 4, np.float32)

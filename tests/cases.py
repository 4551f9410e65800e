"""Seeded problem generator shared by the golden-vector script and the tests."""
import numpy as np
import scipy.sparse as sps

BIG = 1 << 30  # a period larger than max_iter disables that phase (trmf.cpp:654,665,677)

# name -> (T, n, k, lags, density, seed, (lambdaI, lambdaAR, lambdaLag))
GOLDEN_CASES = {
    "tiny_k4":      (60, 40, 4, [1, 2, 5], 0.7, 1, (0.5, 5.0, 0.5)),
    "small_k8":     (120, 70, 8, [1, 2, 5, 24], 0.6, 2, (0.1, 0.1, 0.1)),
    "lagheavy_k5":  (150, 30, 5, list(range(1, 13)) + [24, 48], 0.8, 3, (0.5, 50.0, 0.5)),
    "k1_l1":        (50, 20, 1, [3], 0.9, 4, (0.3, 2.0, 0.2)),
    "ragged_k20":   (90, 64, 20, [1, 7], 0.35, 5, (2.0, 625.0, 0.5)),
    "k40":          (140, 90, 40, [1, 7, 24], 0.9, 6, (0.5, 50.0, 0.5)),
}

PHASES = {  # name -> (period_W, period_H, period_Lag, max_iter)
    "f_only": (BIG, 1, BIG, 1),
    "x_only": (1, BIG, BIG, 1),
    "lag_only": (BIG, BIG, 1, 1),
    "iter1": (1, 1, 1, 1),
    "iter3": (1, 1, 2, 3),
}


def make_problem(T, n, k, lags, density, seed, rank_true=3, noise=0.01, empty_series=True, empty_time=True):
    """Low-rank + noise Y with a Bernoulli(density) mask; one series and one time
    stamp are left completely unobserved (the reference skips empty F rows,
    trmf.cpp:374; an empty time stamp only sees the regulariser).  Values are
    shifted away from 0 so that `csr_matrix(dense)` keeps every observed cell."""
    rng = np.random.RandomState(seed)
    Wt = rng.randn(T, rank_true)
    Ht = rng.randn(n, rank_true)
    Y = Wt @ Ht.T + noise * rng.randn(T, n) + 3.0
    mask = rng.rand(T, n) < density
    if empty_series and n > 3:
        mask[:, 3] = False
    if empty_time and T > 5:
        mask[5, :] = False
    W0 = rng.rand(T, k)
    H0 = rng.rand(n, k)
    L0 = rng.randn(len(lags), k)
    Ysp = sps.csr_matrix(np.where(mask, Y, 0.0))
    return dict(Y=Y, mask=mask, Ysp=Ysp, W0=W0, H0=H0, L0=L0, lags=np.array(sorted(lags), dtype=np.uint32))


def golden_problem(name):
    T, n, k, lags, dens, seed, lams = GOLDEN_CASES[name]
    p = make_problem(T, n, k, lags, dens, seed)
    p["lambdas"] = lams
    return p


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb if nb > 0 else 1.0))

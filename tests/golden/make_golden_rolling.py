"""Generates tests/golden/rolling.npz: rolling_validate (reference python/trmf/trmf.py:303-329) with the UNMODIFIED
reference core doing every fit.

Run in the build container (where /root/reference exists):
    make -C oracle && python tests/golden/make_golden_rolling.py

The reference's own Python module does not import on this stack (scipy.* NumPy aliases), so the loop around
the fits is this repo's re-authored `trmf.rolling_validate` (per-window path) with `trmf.train` replaced by a call into
oracle/_ref/trmf_float64.so; what is recorded is, per case, the factors of every window, the metrics and the inputs.
Cases: missing / full, with and without the per-window NormalizedTransform.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "exp-trmf-nips16_b200")):
    sys.path.insert(0, p)

import trmf  # noqa: E402
import trmf.trmf as tmod  # noqa: E402
from oracle import abi  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LAGS = [1, 2, 12]
KW = dict(k=4, window_size=6, nr_windows=3, lambdaI=0.5, lambdaAR=5.0, lambdaLag=0.5, max_iter=4, threshold=0, seed=0)
CASES = {"missing": (True, None), "missing_tr": (True, True), "full": (False, None), "full_tr": (False, True)}


def series(T, n, seed, zeros):
    rng = np.random.RandomState(seed)
    t = np.arange(T)[:, None]
    Y = 3.0 + np.sin(2 * np.pi * t / 12.0 + rng.rand(1, n) * 6) * (1 + rng.rand(1, n)) + 0.1 * rng.randn(T, n)
    if zeros:
        Y[rng.rand(T, n) < 0.15] = 0.0
    return Y


def reference_train(Y, model, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1, max_iter=10, period_W=1, period_H=1, period_Lag=2,
                    threads=1, missing=False, verbose=0):
    """trmf.train (reference trmf.py:253-264) with the compiled reference core behind it."""
    if model.transform is not None:
        Y = model.transform.preprocess(Y)
    W, H, L = abi.run_reference(Y, model.lag_set, model.W, model.H, model.lag_val, dtype=np.float64, lambdaI=lambdaI,
                                lambdaAR=lambdaAR, lambdaLag=lambdaLag, max_iter=max_iter, period_W=period_W,
                                period_H=period_H, period_Lag=period_Lag, missing=missing, threads=1)
    model.W[:] = W; model.H[:] = H; model.lag_val[:] = L
    return model


def main():
    assert abi.ref_available(np.float64), "build oracle/_ref first: make -C oracle"
    tmod.train = reference_train
    out = {"lags": np.array(LAGS), "kw_names": np.array(sorted(KW)), "kw_vals": np.array([float(KW[k]) for k in sorted(KW)])}
    for name, (missing, transform) in CASES.items():
        Y = series(140, 10, seed=12, zeros=missing)
        models = []
        met = trmf.rolling_validate(Y, LAGS, missing=missing, transform=transform, resident=False, _models_out=models, **KW)
        out[name + "_Y"] = Y
        out[name + "_metrics"] = np.array(list(met))
        for w, m in enumerate(models):
            out["{}_w{}_W".format(name, w)] = m.W.copy()
            out["{}_w{}_H".format(name, w)] = m.H.copy()
            out["{}_w{}_L".format(name, w)] = np.ascontiguousarray(m.lag_val)
        print(name, met)
    path = os.path.join(HERE, "rolling.npz")
    np.savez_compressed(path, **out)
    print(os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

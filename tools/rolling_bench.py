#!/usr/bin/env python
"""rolling_bench.py -- SURVEY 8f-1: rolling_validate (reference python/trmf/trmf.py:303-329) with Y resident in HBM
against the per-window path (one c_trmf_train call per window: convert, upload, train, download) and, on a bounded
sample, the compiled reference (oracle/_ref) on the host cores.

    python tools/rolling_bench.py [--config electricity|traffic|c2] [--max-iter M] [--windows W] [--repeat R]

Workloads (the real datasets are not obtainable offline): "electricity" = BASELINE configs[0] shape
(T = 26 304, n = 370, k = 20, lags 1..24, dense, missing=False, transform on -- run_electricity.py:9-21 otherwise);
"traffic" = configs[2] shape (T = 10 560, n = 963, k = 40, lags 1..24,168,336) run sparse (10 % exact zeros,
missing=True); "c2" = configs[1] shape (T = n = 10 000, k = 40, lags {1,7,24}, 10 % zeros, missing=True).
Wall-clock of the whole call (host work included: this is a host-level API); one JSON line per config.

`oracle/` appears here in exactly the role it has in bench.py's `cpu_baseline` leg: the compiled reference
(`oracle/_ref`, driven through `oracle.abi`) is TIMED next to the product as the CPU baseline (`--no-reference` skips
it); the two measured product paths never touch it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "exp-trmf-nips16_b200"))

CONFIGS = {
    "electricity": dict(T=26304, n=370, k=20, lags=list(range(1, 25)), missing=False, transform=True, zeros=0.0),
    "traffic": dict(T=10560, n=963, k=40, lags=list(range(1, 25)) + [168, 336], missing=True, transform=None, zeros=0.1),
    "c2": dict(T=10000, n=10000, k=40, lags=[1, 7, 24], missing=True, transform=None, zeros=0.1),
    "tiny": dict(T=600, n=50, k=8, lags=[1, 2, 24], missing=True, transform=True, zeros=0.1),
}


def shaped_series(T, n, zeros, seed=0):
    """Positive, per-series scaled load curves with daily and weekly seasonality plus noise (float32)."""
    rng = np.random.RandomState(seed)
    t = np.arange(T, dtype=np.float64)[:, None]
    scale = np.exp(rng.randn(1, n))
    day = np.sin(2 * np.pi * t / 24.0 + rng.rand(1, n) * 2 * np.pi)
    week = np.sin(2 * np.pi * t / 168.0 + rng.rand(1, n) * 2 * np.pi)
    Y = scale * (3.0 + day + 0.5 * week + 0.2 * rng.randn(T, n))
    if zeros > 0:
        Y[rng.rand(T, n) < zeros] = 0.0
    return np.ascontiguousarray(Y, dtype=np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="electricity", choices=sorted(CONFIGS))
    ap.add_argument("--max-iter", type=int, default=20)       # rolling_validate's default (trmf.py:304)
    ap.add_argument("--windows", type=int, default=7)
    ap.add_argument("--window-size", type=int, default=24)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    c = CONFIGS[args.config]
    import trmf
    Y = shaped_series(c["T"], c["n"], c["zeros"])
    kw = dict(k=c["k"], window_size=args.window_size, nr_windows=args.windows, max_iter=args.max_iter, missing=c["missing"],
              transform=c["transform"], lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)
    out = {"workload": "rolling_validate, {}-shaped synthetic".format(args.config), "T": c["T"], "n": c["n"], "k": c["k"],
           "lags": len(c["lags"]), "missing": c["missing"], "transform": bool(c["transform"]), "windows": args.windows,
           "window_size": args.window_size, "max_iter": args.max_iter, "dtype": "f32",
           "observed_entries": int(np.count_nonzero(Y)) if c["missing"] else int(Y.size)}
    res = {}
    for name, resident in (("per_window", False), ("resident", True)):
        times = []
        for r in range(args.repeat + 1):          # first call warms the memory pool / module load
            t0 = time.perf_counter()
            m = trmf.rolling_validate(Y, c["lags"], resident=resident, **kw)
            times.append(time.perf_counter() - t0)
        res[name] = m
        out[name + "_s"] = float(np.median(times[1:] or times))
        out[name + "_s_all"] = [round(x, 4) for x in times]
    out["metrics_identical"] = bool(res["per_window"] == res["resident"])
    out["nd"] = float(res["resident"].nd)
    out["speedup_resident_vs_per_window"] = out["per_window_s"] / out["resident_s"]
    fits = args.windows * args.max_iter
    out["ms_per_outer_iteration_resident"] = 1e3 * out["resident_s"] / fits
    if not args.no_reference:
        # the reference's own solver on the host cores, bounded: ONE window (the first), 2 outer iterations, scaled to
        # windows x max_iter (its cost per iteration is flat in the iteration count; conversions are not counted)
        try:
            import scipy.sparse as sps
            from oracle import abi
            if abi.ref_available(np.float32):
                T0 = c["T"] - args.windows * args.window_size
                Yt = Y[:T0]
                mdl = trmf.Model.initialize(Yt, c["lags"], c["k"], seed=0, transform=c["transform"])
                if mdl.transform is not None:
                    Yt = mdl.transform.preprocess(Yt)
                Yin = sps.csr_matrix(Yt) if c["missing"] else np.ascontiguousarray(Yt)
                cores = os.cpu_count()
                Yin = abi.HostMatrix(Yin, np.float32)      # PyMatrix marshalling (rf_util.py:78-130) outside the timed call
                t0 = time.perf_counter()
                abi.run_reference(Yin, mdl.lag_set, mdl.W, mdl.H, mdl.lag_val, dtype=np.float32, threads=cores, lambdaI=0.5,
                                  lambdaAR=50.0, lambdaLag=0.5, max_iter=2, period_Lag=2, missing=c["missing"])
                dt = time.perf_counter() - t0
                out["cpu_reference"] = {"kind": "reference", "cores": cores, "seconds_2_iterations_one_window": dt,
                                        "extrapolated_s": dt / 2 * fits,
                                        "sample": "first window (T = {}), 2 outer iterations, scaled to {} fits".format(T0, fits)}
                out["speedup_resident_vs_cpu_reference_extrapolated"] = out["cpu_reference"]["extrapolated_s"] / out["resident_s"]
        except Exception as exc:   # the baseline is a report, never a reason to fail the measurement
            out["cpu_reference"] = {"unavailable": repr(exc)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

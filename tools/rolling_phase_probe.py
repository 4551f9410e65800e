#!/usr/bin/env python
"""Per-phase device times (CUDA events) of one outer iteration inside a rolling session, at the shapes of
tools/rolling_bench.py, plus the cost of moving the window (roll_window) and of the one-off ingest (roll_create)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "exp-trmf-nips16_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import trmf
from trmf import session
from rolling_bench import CONFIGS, shaped_series

for name in sys.argv[1:] or ["electricity", "traffic"]:
    c = CONFIGS[name]
    Y = shaped_series(c["T"], c["n"], c["zeros"])
    T_res = c["T"] - 24
    t0 = time.perf_counter()
    rs = session.RollingSession(Y[:T_res], c["lags"], c["k"], missing=c["missing"], dtype=np.float32, lambdaI=0.5, lambdaAR=50.0,
                                lambdaLag=0.5)
    rs.sync()
    t_create = time.perf_counter() - t0
    out = {"config": name, "T": c["T"], "n": c["n"], "k": c["k"], "lags": len(c["lags"]), "missing": c["missing"],
           "roll_create_ms": 1e3 * t_create, "windows": []}
    for T_w in (T_res - 48, T_res - 24, T_res):
        mdl = trmf.Model.initialize(Y[:T_w], c["lags"], c["k"], seed=0, transform=c["transform"])
        tr = mdl.transform
        t0 = time.perf_counter()
        rs.window(T_w, None if tr is None else tr.a, None if tr is None else tr.b)
        t_win = time.perf_counter() - t0
        rs.upload(W=mdl.W, H=mdl.H, lag_val=mdl.lag_val)
        rs.enable_timing(True)
        rows = []
        for it in range(1, 7):
            rs.f_update(); rs.x_update()
            f, x, cg = rs.stat("f_ms"), rs.stat("x_ms"), int(rs.stat("cg_iters"))
            lag = 0.0
            if it % 2 == 0:
                rs.lag_update(); lag = rs.stat("lag_ms")
            rows.append({"f_ms": round(f, 3), "x_ms": round(x, 3), "lag_ms": round(lag, 3), "cg": cg})
        rs.enable_timing(False)
        t0 = time.perf_counter()
        rs.train(max_iter=20)
        rs.sync()
        t_train = time.perf_counter() - t0
        out["windows"].append({"T_w": T_w, "nnz": rs.nnz, "roll_window_ms": round(1e3 * t_win, 3), "iterations": rows,
                               "train_20_iterations_wall_ms": round(1e3 * t_train, 2)})
    rs.close()
    print(json.dumps(out))

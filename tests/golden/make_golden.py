"""Generates tests/golden/*.npz from the UNMODIFIED reference solver.

Run in the build container (where /root/reference exists):
    make -C oracle && python tests/golden/make_golden.py

For every case in tests/cases.py:GOLDEN_CASES the inputs (sparse Y as COO
triplets, dense Y, initial W/H/lag_val, lag set, lambdas) and the outputs of
oracle/_ref/trmf_float{64,32}.so (= python/trmf/corelib/trmf.cpp compiled by
oracle/Makefile) are stored for each phase selection of cases.PHASES, in sparse
mode (missing=1) and dense mode (missing=0).  The reference ships no golden
vectors of its own; these files are the pin.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import abi  # noqa: E402
import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert abi.ref_available(np.float64), "build oracle/_ref first: make -C oracle"
    for name in cases.GOLDEN_CASES:
        p = cases.golden_problem(name)
        lI, lAR, lLag = p["lambdas"]
        coo = p["Ysp"].tocoo()
        out = dict(T=p["Y"].shape[0], n=p["Y"].shape[1], lags=p["lags"], lambdas=np.array(p["lambdas"]),
                   coo_row=coo.row.astype(np.int32), coo_col=coo.col.astype(np.int32), coo_val=coo.data,
                   Ydense=p["Y"], W0=p["W0"], H0=p["H0"], L0=p["L0"])
        for mode, Yin, missing in (("sparse", p["Ysp"], True), ("dense", p["Y"], False)):
            for phase, (pW, pH, pL, iters) in cases.PHASES.items():
                for tag, dt in (("f64", np.float64), ("f32", np.float32)):
                    if tag == "f32" and phase not in ("iter1",):
                        continue
                    W, H, L = abi.run_reference(Yin, p["lags"], p["W0"], p["H0"], p["L0"], dtype=dt,
                                                lambdaI=lI, lambdaAR=lAR, lambdaLag=lLag, max_iter=iters,
                                                period_W=pW, period_H=pH, period_Lag=pL, missing=missing, threads=1)
                    key = "{}_{}_{}".format(mode, phase, tag)
                    out[key + "_W"], out[key + "_H"], out[key + "_L"] = W, H, np.ascontiguousarray(L)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

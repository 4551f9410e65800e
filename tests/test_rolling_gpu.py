"""GPU parity tests of the rolling-window session: ``rolling_validate`` with Y resident in HBM.

What must hold (reference python/trmf/trmf.py:303-329 is the behaviour; there are no reference tests for it):
  * index work bit-exact: the windowed by-time CSR / by-series CSC the device holds for Y[:T_w] equal the arrays the
    reference's PyMatrix builds on the host from the same slice (rf_util.py:88-98), and the per-series affine
    transform gives NumPy's bits;
  * a window trains exactly like a fresh session on Y[:T_w] from the same factors (bitwise, both precisions);
  * ``rolling_validate(resident=True)`` returns the metrics of the per-window path exactly, and the float64 forecasts
    agree with the same loop driven by the NumPy oracle within 1e-7.
"""
import numpy as np
import pytest
import scipy.sparse as sps

import cases
import trmf
import trmf.trmf as tmod
from trmf import session
from trmf.rf_util import PyMatrix
from oracle import trmf_numpy as tn
from test_rolling_cpu import oracle_train, series

pytestmark = pytest.mark.gpu

LAMS = dict(lambdaI=0.5, lambdaAR=50.0, lambdaLag=0.5)


def sparse_case(T, n, k, lags, dens, seed, dtype):
    p = cases.make_problem(T, n, k, lags, dens, seed)
    Y = sps.csr_matrix((p["Ysp"].data.astype(dtype), p["Ysp"].indices, p["Ysp"].indptr), shape=p["Ysp"].shape)
    return p, Y


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_window_index_work_is_bit_exact(dtype, fmt):
    p, Y = sparse_case(300, 130, 8, [1, 2, 5], 0.4, 11, dtype)
    rs = session.RollingSession(Y.asformat(fmt), p["lags"], 8, missing=True, dtype=dtype)
    rng = np.random.RandomState(0)
    a = (0.5 + rng.rand(130)).astype(dtype)
    b = rng.randn(130).astype(dtype)
    for T_w, tr in [(300, False), (1, False), (6, True), (137, False), (137, True), (299, True), (300, True), (42, False)]:
        rs.window(T_w, a if tr else None, b if tr else None)
        row_ptr, col_idx, val_t, col_ptr, row_idx, val = rs.export_window()
        Yw = Y[:T_w]
        if tr:
            Yw = sps.csr_matrix(Yw, copy=True)
            Yw.data = Yw.data * a[Yw.indices] + b[Yw.indices]       # NormalizedTransform.preprocess, trmf.py:90-92
        ref = PyMatrix(Yw, dtype).py_buf                                # host twin build, rf_util.py:88-98
        assert rs.nnz == Yw.nnz
        for name, got in (("row_ptr", row_ptr), ("col_idx", col_idx), ("val_t", val_t), ("col_ptr", col_ptr),
                          ("row_idx", row_idx), ("val", val)):
            assert got.dtype == ref[name].dtype and np.array_equal(got, ref[name]), (T_w, tr, name)
    rs.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(300, 130), (64, 1), (3, 77), (97, 32), (50, 1000)])
def test_dense_array_with_missing_is_sparsified_like_csr_matrix(dtype, shape):
    """rolling_validate's csr_matrix(Y_trn) (trmf.py:320-321) done on the device: the non-zero cells, bit-exact."""
    T, n = shape
    rng = np.random.RandomState(T + n)
    Y = rng.randn(T, n).astype(dtype)
    Y[rng.rand(T, n) < 0.3] = 0.0
    Y[rng.rand(T, n) < 0.02] = -0.0          # negative zero is a zero for scipy too
    Y[T // 2, :] = 0.0                       # an empty time stamp
    Y[:, n // 2] = 0.0                       # an empty series
    if n > 40:
        Y[1, 30:40] = 0.0
    rs = session.RollingSession(Y, [1, 2], 4, missing=True, dtype=dtype)
    for T_w in sorted({T, max(1, T // 3), max(1, T - 1)}):
        rs.window(T_w)
        ref = PyMatrix(sps.csr_matrix(Y[:T_w]), dtype).py_buf
        for name, got in zip(("row_ptr", "col_idx", "val_t", "col_ptr", "row_idx", "val"), rs.export_window()):
            assert got.dtype == ref[name].dtype and np.array_equal(got, ref[name]), (T_w, name)
    rs.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["sparse", "dense"])
@pytest.mark.parametrize("k", [8, 40])
def test_window_trains_like_a_fresh_session(dtype, mode, k):
    lags = [1, 7, 24]
    p, Ysp = sparse_case(260, 150, k, lags, 0.6, 5, dtype)
    Y = Ysp if mode == "sparse" else p["Y"].astype(dtype)
    missing = mode == "sparse"
    rs = session.RollingSession(Y, lags, k, missing=missing, dtype=dtype, **LAMS)
    W0, H0, L0 = (x.astype(dtype) for x in (p["W0"], p["H0"], p["L0"]))
    for T_w in (200, 230, 260):          # growing windows, like rolling_validate; each restarted from (W0, H0, L0)
        rs.window(T_w)
        rs.upload(W=W0[:T_w], H=H0, lag_val=L0)
        rs.train(max_iter=2, period_W=1, period_H=1, period_Lag=1)
        W, H, L = rs.download()
        fresh = session.Session(Y[:T_w], lags, W0[:T_w], H0, L0, missing=missing, dtype=dtype, **LAMS)
        fresh.train(max_iter=2, period_W=1, period_H=1, period_Lag=1)
        Wf, Hf, Lf = fresh.download()
        fresh.close()
        assert np.array_equal(W, Wf) and np.array_equal(H, Hf) and np.array_equal(L, Lf), (T_w,)
        Y64 = Y[:T_w].astype(np.float64)
        Wo, Ho, Lo = tn.train(Y64, lags, W0[:T_w].astype(np.float64), H0.astype(np.float64), L0.astype(np.float64),
                              max_iter=2, period_Lag=1, missing=missing, **LAMS)
        tol = 1e-9 if dtype == np.float64 else 3e-5      # two iterations compound: 1e-5 per iteration
        assert max(cases.rel(W, Wo), cases.rel(H, Ho), cases.rel(L, Lo)) < tol
    rs.close()


def test_partial_row_io_and_errors():
    p, Y = sparse_case(80, 30, 4, [1, 2], 0.5, 2, np.float32)
    rs = session.RollingSession(Y, [1, 2], 4, dtype=np.float32)
    rs.window(60)
    W = np.arange(60 * 4, dtype=np.float32).reshape(60, 4)
    rs.upload(W=W, H=p["H0"].astype(np.float32), lag_val=p["L0"].astype(np.float32))
    rs.upload_W_rows(50, -W[50:60])
    assert np.array_equal(rs.download_W_rows(45, 15), np.vstack([W[45:50], -W[50:60]]))
    with pytest.raises(RuntimeError, match="outside"):
        rs.upload_W_rows(58, W[:5])
    with pytest.raises(RuntimeError, match="outside"):
        rs.window(81)
    rs.close()
    with pytest.raises(RuntimeError, match="row-major"):
        session.RollingSession(PyMatrix(np.asfortranarray(p["Y"]), np.float64), [1, 2], 4, missing=False, dtype=np.float64)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("missing", [True, False])
@pytest.mark.parametrize("transform", [None, True])
def test_rolling_validate_resident_equals_per_window_path(dtype, missing, transform):
    Y = series(400, 60, seed=3, zeros=missing).astype(dtype)
    kw = dict(k=8, window_size=12, nr_windows=4, max_iter=6, missing=missing, transform=transform, **LAMS)
    host = trmf.rolling_validate(Y, [1, 2, 12, 24], resident=False, **kw)
    res = trmf.rolling_validate(Y, [1, 2, 12, 24], resident=True, **kw)
    assert host == res
    assert np.isfinite(res.nd) and res.nd < 1.0
    # any memory order of Y is accepted (NumPy's own reductions round differently on it, hence not bitwise)
    res_f = trmf.rolling_validate(np.asfortranarray(Y), [1, 2, 12, 24], resident=True, **kw)
    assert all(abs(x - y) <= 1e-3 * abs(y) for x, y in zip(res_f, res))


@pytest.mark.parametrize("missing", [True, False])
def test_rolling_validate_resident_matches_the_oracle_loop(monkeypatch, missing):
    Y = series(220, 25, seed=6, zeros=missing)
    kw = dict(k=5, window_size=8, nr_windows=3, max_iter=3, missing=missing, transform=True, lambdaI=0.5, lambdaAR=5.0,
              lambdaLag=0.5)
    res = trmf.rolling_validate(Y, [1, 2, 12], resident=True, **kw)
    monkeypatch.setattr(tmod, "train", oracle_train)
    ora = trmf.rolling_validate(Y, [1, 2, 12], resident=False, **kw)
    for got, want in zip(res, ora):
        assert abs(got - want) <= 1e-7 * abs(want)


def test_grid_search_over_one_resident_copy():
    """The grid points of grid_search (trmf.py:331-346) reuse one rolling session: same results as fresh ones."""
    Y = series(300, 40, seed=8).astype(np.float32)
    grid = {"lambdaAR": [5.0, 50.0], "lambdaI": [0.5, 2.0]}
    kw = dict(k=8, window_size=10, nr_windows=3, max_iter=4, missing=True, transform=True)
    res, best = trmf.grid_search(Y, [1, 2, 12], grid, resident=True, **kw)
    ref, best_ref = trmf.grid_search(Y, [1, 2, 12], grid, resident=False, **kw)
    assert [r["metrics"] for r in res] == [r["metrics"] for r in ref] and best == best_ref
    assert len({r["metrics"].nd for r in res}) == 4          # the weights did reach the device


def test_grid_search_with_concurrent_workers_gives_the_serial_results(capsys):
    """grid_search(workers=3): grid points train at the same time on their own resident sessions / CUDA streams; every point's
    metrics, the order of the results, the best point and the print-out are the serial run's (a grid in k, the weights and
    window_size, so workers also open and retire sessions of different keys)."""
    import time
    Y = series(600, 60, seed=21).astype(np.float32)
    grid = {"k": [4, 8], "lambdaAR": [5.0, 50.0, 500.0], "lambdaI": [0.5, 2.0], "window_size": [8, 12]}
    kw = dict(nr_windows=3, max_iter=6, missing=False, transform=True, lambdaLag=0.5)
    t0 = time.perf_counter()
    res1, best1 = trmf.grid_search(Y, [1, 2, 12], grid, workers=1, **kw)
    t1 = time.perf_counter()
    out1 = capsys.readouterr().out
    res3, best3 = trmf.grid_search(Y, [1, 2, 12], grid, workers=3, **kw)
    t2 = time.perf_counter()
    out3 = capsys.readouterr().out
    assert [r["kws"] for r in res1] == [r["kws"] for r in res3] and len(res1) == 24
    assert [r["metrics"] for r in res1] == [r["metrics"] for r in res3] and best1 == best3 and out1 == out3
    print("grid_search 24 points: serial {:.3f} s, 3 workers {:.3f} s".format(t1 - t0, t2 - t1))


def test_grid_search_matches_the_oracle_loop(monkeypatch, capsys):
    """grid_search (reference trmf.py:331-346) over a grid in k, window_size and the weights -- resident sessions shared where
    the key allows, reordered internally -- against the reference's own loop order with the NumPy oracle doing every fit
    (float64): same metrics per grid point, same best point, results in the grid's order."""
    Y = series(200, 18, seed=12)
    grid = {"k": [3, 5], "lambdaAR": [5.0, 50.0], "window_size": [6, 8]}
    kw = dict(nr_windows=2, max_iter=3, missing=True, transform=True, lambdaI=0.5, lambdaLag=0.5)
    res, best = trmf.grid_search(Y, [1, 2, 12], grid, resident=True, **kw)
    monkeypatch.setattr(tmod, "train", oracle_train)
    ora, best_ora = trmf.grid_search(Y, [1, 2, 12], grid, resident=False, **kw)
    strip = lambda kws: {k: v for k, v in kws.items() if k != "resident"}
    assert [strip(r["kws"]) for r in res] == [strip(r["kws"]) for r in ora] and len(res) == 8
    for r, o in zip(res, ora):
        for got, want in zip(r["metrics"], o["metrics"]):
            assert abs(got - want) <= 1e-7 * abs(want)
    assert [r["metrics"].m_nd for r in res].index(best.m_nd) == [o["metrics"].m_nd for o in ora].index(best_ora.m_nd)


@pytest.mark.parametrize("name", ["missing", "missing_tr", "full", "full_tr"])
@pytest.mark.parametrize("resident", [True, False])
def test_rolling_validate_matches_fits_by_the_reference_core(name, resident):
    """tests/golden/rolling.npz: the loop of trmf.py:303-329 with the compiled reference doing every fit.  The float64
    CUDA path (resident session and per-window calls) reproduces every window's factors to 1e-8 and the metrics to
    1e-7 (3 windows x 4 iterations from the reference's own warm starts; per-iteration parity is 1e-9)."""
    from test_rolling_cpu import ROLL_CASES, check_models_against_golden, rolling_golden
    g, kw, lags = rolling_golden()
    missing, transform = ROLL_CASES[name]
    models = []
    met = trmf.rolling_validate(g[name + "_Y"], lags, missing=missing, transform=transform, resident=resident,
                                _models_out=models, **kw)
    check_models_against_golden(g, name, models, 1e-8)
    assert np.allclose(np.array(list(met)), g[name + "_metrics"], rtol=1e-7, atol=0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_window_statistics_on_the_device_are_numpys_bits(dtype):
    """NormalizedTransform's per-window statistics (reference trmf.py:84-88) for a dense resident Y: computed on the device in
    NumPy's axis-0 summation order -> bit-equal mean / std, hence bit-equal a, b, at the electricity shape."""
    import trmf
    from trmf.session import RollingSession
    rng = np.random.RandomState(5)
    T, n = 26304, 370
    t = np.arange(T)[:, None]
    Y = (np.abs(rng.randn(T, n)) * (0.5 + 20 * rng.rand(1, n)) * (1.0 + 0.5 * np.sin(2 * np.pi * t / 24.0)) + 0.1).astype(dtype)
    Y[:, 7] = 3.25                                   # a constant series: std == 0 -> treated as 1
    s = RollingSession(Y, list(range(1, 25)), 4, missing=False, dtype=dtype)
    try:
        for T_w in (T, T - 24 * 7, 1000, 1):
            mean, std = s.window_stats(T_w)
            assert np.array_equal(mean, Y[:T_w].mean(axis=0)) and np.array_equal(std, Y[:T_w].std(axis=0))
            tr_dev, tr_host = trmf.NormalizedTransform(None, _stats=(mean, std)), trmf.NormalizedTransform(Y[:T_w])
            assert np.array_equal(tr_dev.a, tr_host.a) and np.array_equal(tr_dev.b, tr_host.b)
    finally:
        s.close()
    sp_sess = RollingSession(Y[:2000], [1, 2], 4, missing=True, dtype=dtype)     # sparsified on the device: no dense copy kept
    assert sp_sess.window_stats(1000) is None
    sp_sess.close()

"""Host logic of the device-resident ``rolling_validate``, checked on the CPU.

The resident path (``trmf.trmf._rolling_resident``) must hand every window the same bits the reference's loop
(python/trmf/trmf.py:303-329) would: same initial factors, same warm-start rows, same per-window transform.  Here the
CUDA session is replaced by a stand-in that keeps "device" state across windows like ``session.RollingSession`` does
and runs the NumPy oracle on it, and ``trmf.train`` is replaced by the same oracle, so both paths must agree exactly.
The real session is exercised by tests/test_rolling_gpu.py.
"""
import numpy as np
import pytest
import scipy.sparse as sps

import trmf
import trmf.trmf as tmod
import trmf.session as smod
from oracle import trmf_numpy as tn


def oracle_train(Y, model, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1, max_iter=10, period_W=1, period_H=1, period_Lag=2,
                 threads=1, missing=False, verbose=0):
    """Stand-in for trmf.train (reference trmf.py:253-264) backed by the NumPy oracle."""
    if model.transform is not None:
        Y = model.transform.preprocess(Y)
    W, H, L = tn.train(Y, model.lag_set, model.W, model.H, model.lag_val, lambdaI=lambdaI, lambdaAR=lambdaAR,
                       lambdaLag=lambdaLag, max_iter=max_iter, period_W=period_W, period_H=period_H, period_Lag=period_Lag,
                       missing=missing)
    model.W[:] = W; model.H[:] = H; model.lag_val[:] = L
    return model


class FakeRollingSession(object):
    """Same interface and state model as session.RollingSession, oracle arithmetic."""
    log = []

    def __init__(self, Y, lag_set, k, missing=True, dtype=None, device=0, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1):
        self.Y = Y.tocsr() if sps.issparse(Y) else np.array(Y)
        if missing and not sps.issparse(Y):
            self.Y = sps.csr_matrix(self.Y)      # the device keeps the non-zero cells (trmf.py:320-321)
        self.missing = missing
        self.lag_set = np.sort(np.asarray(lag_set))
        self.dtype = np.dtype(dtype)
        self.T_cap, self.n = Y.shape
        self.k = k
        self.lams = (lambdaI, lambdaAR, lambdaLag)
        self.W = np.full((self.T_cap, k), np.nan, dtype=self.dtype)     # stale rows must never be read
        self.H = np.full((self.n, k), np.nan, dtype=self.dtype)
        self.L = np.full((len(self.lag_set), k), np.nan, dtype=self.dtype, order="F")
        self.closed = False
        FakeRollingSession.log.append(("create", Y.shape, sps.issparse(Y)))

    def set_params(self, lambdaI, lambdaAR, lambdaLag):
        self.lams = (lambdaI, lambdaAR, lambdaLag)

    def window_stats(self, T_w):
        if sps.issparse(self.Y):
            return None
        return self.Y[:T_w].mean(axis=0), self.Y[:T_w].std(axis=0)

    def window(self, T_w, scale=None, offset=None):
        assert 0 < T_w <= self.T_cap and not self.closed
        self.T = T_w
        self.a = None if scale is None else np.asarray(scale).reshape(1, -1)
        self.b = None if offset is None else np.asarray(offset).reshape(1, -1)
        FakeRollingSession.log.append(("window", T_w, scale is not None))

    def upload(self, W=None, H=None, lag_val=None):
        assert W.shape == (self.T, self.k)
        self.W[:self.T] = W; self.H[:] = H; self.L[:] = lag_val
        FakeRollingSession.log.append(("upload", W.shape[0]))

    def upload_W_rows(self, row0, rows):
        assert row0 + rows.shape[0] == self.T
        self.W[row0:row0 + rows.shape[0]] = rows
        FakeRollingSession.log.append(("rows", row0, rows.shape[0]))

    def train(self, max_iter=10, period_W=1, period_H=1, period_Lag=2, verbose=0):
        Y = self.Y[:self.T]
        if self.a is not None:
            if sps.issparse(Y):
                Y = sps.csr_matrix(Y, copy=True)
                Y.data = Y.data * self.a[0, Y.indices] + self.b[0, Y.indices]
            else:
                Y = Y * self.a + self.b
        W, H, L = tn.train(Y, self.lag_set, self.W[:self.T], self.H, self.L, lambdaI=self.lams[0], lambdaAR=self.lams[1],
                           lambdaLag=self.lams[2], max_iter=max_iter, period_W=period_W, period_H=period_H,
                           period_Lag=period_Lag, missing=self.missing)
        self.W[:self.T] = W; self.H[:] = H; self.L[:] = L

    def download_into(self, W, H, lag_val):
        W[:] = self.W[:self.T]; H[:] = self.H; lag_val[:] = self.L

    def close(self):
        self.closed = True


def series(T, n, seed, zeros=True):
    rng = np.random.RandomState(seed)
    t = np.arange(T)[:, None]
    Y = 3.0 + np.sin(2 * np.pi * t / 12.0 + rng.rand(1, n) * 6) * (1 + rng.rand(1, n)) + 0.1 * rng.randn(T, n)
    if zeros:
        Y[rng.rand(T, n) < 0.15] = 0.0     # missing=True reads exact zeros as unobserved (trmf.py:320-321)
    return Y


@pytest.mark.parametrize("missing", [True, False])
@pytest.mark.parametrize("transform", [None, True])
def test_resident_rolling_hands_every_window_the_same_bits(monkeypatch, missing, transform):
    Y = series(150, 9, seed=4, zeros=missing)
    kw = dict(k=3, window_size=6, nr_windows=4, lambdaI=0.5, lambdaAR=5.0, lambdaLag=0.5, max_iter=3, missing=missing,
              threshold=0, transform=transform, seed=0)
    monkeypatch.setattr(tmod, "train", oracle_train)
    monkeypatch.setattr(smod, "RollingSession", FakeRollingSession)
    FakeRollingSession.log = []
    host = trmf.rolling_validate(Y, [1, 2, 12], resident=False, **kw)
    res = trmf.rolling_validate(Y, [1, 2, 12], resident=True, **kw)
    assert host == res                                      # namedtuple of floats: exact
    assert all(np.isfinite(v) for v in res)
    log = FakeRollingSession.log
    assert log[0] == ("create", (150 - 6, 9), False)        # the longest training prefix, uploaded once, as it is
    assert [e for e in log if e[0] == "window"] == [("window", 150 - 6 * (4 - w), transform is not None) for w in range(4)]
    assert [e for e in log if e[0] == "upload"] == [("upload", 126)]                 # full factors: first window only
    assert [e for e in log if e[0] == "rows"] == [("rows", 126 + 6 * w, 6) for w in range(3)]   # then window_size rows


def test_resident_rolling_models_match_window_by_window(monkeypatch):
    Y = series(120, 7, seed=9)
    monkeypatch.setattr(smod, "RollingSession", FakeRollingSession)
    models = []
    tmod._rolling_resident(Y, [1, 3], 4, 5, 3, 0.5, 5.0, 0.5, 2, True, 0, None, 0, models_out=models)
    prev = None
    for w, m in enumerate(models):     # the reference's loop, one window at a time
        trn_end = 120 - (3 - w) * 5
        Yt = sps.csr_matrix(Y[:trn_end])
        ref = trmf.Model.initialize(Yt, [1, 3], 4, seed=0, warm_start_model=prev)
        oracle_train(Yt, ref, lambdaI=0.5, lambdaAR=5.0, lambdaLag=0.5, max_iter=2, missing=True)
        assert np.array_equal(ref.W, m.W) and np.array_equal(ref.H, m.H) and np.array_equal(ref.lag_val, m.lag_val)
        prev = ref


def test_grid_search_shares_the_resident_copy(monkeypatch, capsys):
    """one upload of Y per (k, window_size, missing) for the whole grid; results as the reference's loop
    (trmf.py:331-346) over the per-window path."""
    Y = series(110, 6, seed=2)
    monkeypatch.setattr(tmod, "train", oracle_train)
    monkeypatch.setattr(smod, "RollingSession", FakeRollingSession)
    grid = {"lambdaAR": [5.0, 50.0], "lambdaI": [0.5, 2.0], "k": [2, 3]}
    kw = dict(window_size=5, nr_windows=2, max_iter=2, missing=True)
    FakeRollingSession.log = []
    res, best = trmf.grid_search(Y, [1, 2], grid, resident=True, **kw)
    creates = [e for e in FakeRollingSession.log if e[0] == "create"]
    assert len(res) == 8 and len(creates) == 2                      # one session per k, not one per grid point / window
    ref, best_ref = trmf.grid_search(Y, [1, 2], grid, resident=False, **kw)
    assert [r["metrics"] for r in res] == [r["metrics"] for r in ref] and best == best_ref
    assert [r["kws"]["lambdaAR"] for r in res] == [5.0, 5.0, 5.0, 5.0, 50.0, 50.0, 50.0, 50.0]


def test_grid_search_workers_host_logic(monkeypatch, capsys, tmp_path):
    """grid_search(workers=n): grid points run on a thread pool, one set of parked sessions per worker thread; the results, their
    order, the best point, the print-out and the pickle are those of the serial run (oracle-backed stand-in session, no device)."""
    import pickle
    Y = series(110, 6, seed=3)
    monkeypatch.setattr(tmod, "train", oracle_train)
    monkeypatch.setattr(smod, "RollingSession", FakeRollingSession)
    grid = {"lambdaAR": [5.0, 50.0], "lambdaI": [0.5, 2.0], "k": [2, 3], "window_size": [4, 5]}
    kw = dict(nr_windows=2, max_iter=2, missing=True)
    res, best = trmf.grid_search(Y, [1, 2], grid, resident=True, **kw)
    out = capsys.readouterr().out
    FakeRollingSession.log = []
    pkl = tmp_path / "grid.pkl"
    res3, best3 = trmf.grid_search(Y, [1, 2], grid, resident=True, workers=3, pkl_file=str(pkl), **kw)
    out3 = capsys.readouterr().out
    assert len(res3) == 16 and [r["kws"] for r in res] == [r["kws"] for r in res3]
    assert [r["metrics"] for r in res] == [r["metrics"] for r in res3] and best == best3 and out == out3
    creates = [e for e in FakeRollingSession.log if e[0] == "create"]
    assert 4 <= len(creates) <= 12       # 4 session keys (k x window_size), opened by at most 3 workers each
    with open(pkl, "rb") as fh:
        assert [r["metrics"] for r in pickle.load(fh)] == [r["metrics"] for r in res]


def test_resident_default_and_switches(monkeypatch):
    calls = []
    monkeypatch.setattr(tmod, "_rolling_resident", lambda *a, **k: calls.append("resident") or trmf.Metrics.default())
    monkeypatch.setattr(tmod, "train", oracle_train)
    Y = series(60, 4, seed=1).astype(np.float32)
    kw = dict(k=2, window_size=4, nr_windows=2, max_iter=1)
    trmf.rolling_validate(Y, [1, 2], **kw)
    assert calls == ["resident"]                                    # default for a dense float array
    trmf.rolling_validate(Y, [1, 2], resident=False, **kw)
    monkeypatch.setenv("TRMF_B200_ROLLING_HOST", "1")
    trmf.rolling_validate(Y, [1, 2], **kw)
    assert calls == ["resident"]


def test_resident_rolling_is_gpu_only():
    """No device -> loud failure, never a host computation."""
    if trmf.trmf._clib.clib_float32.trmf_b200_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="GPU-only"):
        trmf.rolling_validate(series(60, 4, seed=1), [1, 2], k=2, window_size=4, nr_windows=2, max_iter=1, resident=True)


# ---- golden vectors: the same loop with the UNMODIFIED reference core doing every fit (tests/golden/make_golden_rolling.py) ----
ROLL_CASES = {"missing": (True, None), "missing_tr": (True, True), "full": (False, None), "full_tr": (False, True)}


def rolling_golden():
    from conftest import load_golden
    g = load_golden("rolling")
    kw = {k: v for k, v in zip(g["kw_names"], g["kw_vals"])}
    for k in ("k", "window_size", "nr_windows", "max_iter", "seed"):
        kw[k] = int(kw[k])
    return g, kw, [int(l) for l in g["lags"]]


def check_models_against_golden(g, name, models, tol):
    import cases
    assert len(models) == 3
    for w, m in enumerate(models):
        for key, got in (("W", m.W), ("H", m.H), ("L", m.lag_val)):
            assert cases.rel(got, g["{}_w{}_{}".format(name, w, key)]) < tol, (name, w, key)


@pytest.mark.parametrize("name", list(ROLL_CASES))
@pytest.mark.parametrize("resident", [False, True])
def test_rolling_loop_matches_the_reference_core(monkeypatch, name, resident):
    """NumPy oracle as the solver, per-window loop and resident host logic alike, against fits by the compiled reference:
    every window's factors to 1e-9, the metrics to 1e-9 -- pins warm start, transform and windowing over 3 windows."""
    g, kw, lags = rolling_golden()
    missing, transform = ROLL_CASES[name]
    monkeypatch.setattr(tmod, "train", oracle_train)
    monkeypatch.setattr(smod, "RollingSession", FakeRollingSession)
    models = []
    met = trmf.rolling_validate(g[name + "_Y"], lags, missing=missing, transform=transform, resident=resident,
                                _models_out=models, **kw)
    check_models_against_golden(g, name, models, 1e-9)
    assert np.allclose(np.array(list(met)), g[name + "_metrics"], rtol=1e-9, atol=0)


# ---- property test: any shape / window plan / option mix, resident host logic == the reference's per-window loop ----
try:
    from hypothesis import given, settings, strategies as st, HealthCheck
    HAVE_HYPOTHESIS = True
except Exception:       # pragma: no cover
    HAVE_HYPOTHESIS = False

if HAVE_HYPOTHESIS:
    @settings(max_examples=40, deadline=None, derandomize=True, database=None,
              suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(n=st.integers(1, 6), k=st.integers(1, 4), window_size=st.integers(1, 5), nr_windows=st.integers(1, 4),
           lag_pool=st.lists(st.integers(0, 9), min_size=1, max_size=4, unique=True), missing=st.booleans(),
           transform=st.sampled_from([None, True]), max_iter=st.integers(1, 3), seed=st.integers(0, 5),
           f32=st.booleans())
    def test_resident_rolling_property(monkeypatch, n, k, window_size, nr_windows, lag_pool, missing, transform, max_iter, seed, f32):
        from hypothesis import assume
        assume(nr_windows * window_size >= 2)      # Metrics' MASE needs two rows of truth (reference trmf.py:290-294 too)
        T = max(lag_pool) + 12 + nr_windows * window_size
        Y = series(T, n, seed=seed, zeros=missing)
        if f32:
            Y = Y.astype(np.float32)
        if missing:
            Y[0, :] = 1.0          # keep every series observed at least once in every window
        hor = Y[T - nr_windows * window_size:]
        hor[hor == 0] = 1.0        # (truth rows: the metrics divide by their sums)
        monkeypatch.setattr(tmod, "train", oracle_train)
        monkeypatch.setattr(smod, "RollingSession", FakeRollingSession)
        kw = dict(k=k, window_size=window_size, nr_windows=nr_windows, lambdaI=0.5, lambdaAR=5.0, lambdaLag=0.5, max_iter=max_iter,
                  missing=missing, transform=transform, seed=seed)
        m_host, m_res = [], []
        with np.errstate(all="ignore"):
            host = trmf.rolling_validate(Y, lag_pool, resident=False, _models_out=m_host, **kw)
            res = trmf.rolling_validate(Y, lag_pool, resident=True, _models_out=m_res, **kw)
        assert len(m_host) == len(m_res) == nr_windows
        for a, b in zip(m_host, m_res):
            assert np.array_equal(a.W, b.W) and np.array_equal(a.H, b.H) and np.array_equal(a.lag_val, b.lag_val)
        assert np.array_equal(np.array(list(host)), np.array(list(res)), equal_nan=True)

"""Python surface of the B200-native TRMF solver.

Same names, arguments, defaults and side effects as the reference's
``python/trmf/trmf.py`` (``Model``, ``train``, ``rolling_validate``,
``grid_search``, ``Metrics``, ``NormalizedTransform``) so that user code written
against ``rofuyu/exp-trmf-nips16`` runs unchanged; the work behind ``train`` is
one ``c_trmf_train`` call into hand-written sm_100a CUDA (``../csrc``).  Host
orchestration stays NumPy, as in the reference; there is no CPU solver here.

Shapes (reference python/README.md:37, trmf.h:31-52): ``Y`` is T x n (rows =
time stamps), ``W`` T x k (temporal factor, the paper's X), ``H`` n x k (series
factor, the paper's F), ``lag_val`` L x k column-major, ``lag_set`` sorted uint32.
"""
import collections
import ctypes
import itertools
import os
import pickle
from ctypes import POINTER, byref, c_double, c_int32, c_uint32

import numpy as np
import scipy.sparse as smat


try:
    from .rf_util import PyMatrix, fillprototype, load_dynamic_library
except ImportError:  # running as a script, like the reference allows
    from rf_util import PyMatrix, fillprototype, load_dynamic_library


_RNG_LOCK = __import__("threading").Lock()   # np.random's global stream is shared by every thread (Model.initialize)

class corelib(object):
    """The float32 / float64 pair of CUDA libraries (reference trmf.py:19-75)."""

    _ARGS = [
        POINTER(PyMatrix),  # Y
        POINTER(c_uint32),  # lag_set
        c_uint32,           # lag_size
        POINTER(PyMatrix),  # W
        POINTER(PyMatrix),  # H
        POINTER(PyMatrix),  # lag_val
        c_int32,            # warm_start
        c_double,           # lambdaI
        c_double,           # lambdaAR
        c_double,           # lambdaLag
        c_int32,            # max_iter
        c_int32,            # period_W
        c_int32,            # period_H
        c_int32,            # period_Lag
        c_int32,            # threads
        c_int32,            # missing
        c_int32,            # verbose
    ]

    def __init__(self, dirname, soname, forced_rebuild=False):
        self.clib_float32 = load_dynamic_library(dirname, soname + "_float32", forced_rebuild=forced_rebuild)
        self.clib_float64 = load_dynamic_library(dirname, soname + "_float64", forced_rebuild=forced_rebuild)
        for lib in (self.clib_float32, self.clib_float64):
            fillprototype(lib.c_trmf_train, None, corelib._ARGS)
            fillprototype(lib.trmf_b200_last_error, ctypes.c_char_p, [])
            fillprototype(lib.trmf_b200_device_count, c_int32, [])

    def pick(self, dtype):
        return self.clib_float64 if np.dtype(dtype) == np.float64 else self.clib_float32

    def train(self, pyY, lag_set, pyW, pyH, pylag_val, warm_start=True,
              lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1, max_iter=10,
              period_W=1, period_H=1, period_Lag=2, threads=1, missing=False, verbose=0):
        clib = self.pick(pyY.dtype)
        if verbose != 0:
            print("perform float64 computation" if clib is self.clib_float64 else "perform float32 computation")
        if clib.trmf_b200_device_count() <= 0:
            raise RuntimeError("trmf: no CUDA device visible; this solver is GPU-only (B200, sm_100a) "
                               "and has no CPU fallback")
        lag_set = np.ascontiguousarray(lag_set, dtype=np.uint32)
        clib.c_trmf_train(byref(pyY), lag_set.ctypes.data_as(POINTER(c_uint32)), len(lag_set),
                          byref(pyW), byref(pyH), byref(pylag_val), c_int32(int(bool(warm_start))),
                          lambdaI, lambdaAR, lambdaLag, max_iter, period_W, period_H, period_Lag,
                          threads, int(bool(missing)), verbose)
        err = clib.trmf_b200_last_error()
        if err:
            raise RuntimeError("trmf (CUDA): " + err.decode())


forced_rebuild = False
corelib_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "corelib/")
soname = "trmf"
_clib = corelib(corelib_path, soname, forced_rebuild=forced_rebuild)


class NormalizedTransform(object):
    """Per-series affine normalisation y -> a*y + b with a = 1/std, b = -mean/std
    over time; zero std is treated as 1 (reference trmf.py:82-96)."""

    def __init__(self, Y, _stats=None):
        if _stats is not None:      # (mean, std) of Y's columns, already computed (on the device: session.RollingSession.window_stats)
            mean, std = (np.asarray(x).reshape(1, -1) for x in _stats)
            std = std.copy()
        else:
            Yd = Y.toarray() if smat.issparse(Y) else np.asarray(Y)
            mean = Yd.mean(axis=0).reshape(1, -1)
            std = Yd.std(axis=0).reshape(1, -1).copy()
        std[std == 0] = 1.0
        self.a = 1.0 / std
        self.b = -self.a * mean

    def preprocess(self, Y):
        assert Y.shape[1] == self.a.shape[1]
        if smat.issparse(Y):
            # only the stored entries are observations: transform those, keep the pattern
            out = smat.csr_matrix(Y, copy=True)
            cols = out.indices
            out.data = out.data * self.a[0, cols] + self.b[0, cols]
            return out
        return Y * self.a + self.b

    def postprocess(self, Y):
        assert Y.shape[1] == self.a.shape[1]
        return (Y - self.b) / self.a


class Model(object):
    def __init__(self, pyW=None, pyH=None, pylag_val=None, lag_set=None, transform=None):
        self.pyW = pyW
        self.pyH = pyH
        self.pylag_val = pylag_val
        self.lag_set = lag_set
        self.transform = transform

    # ---- views on the buffers the solver updates in place (trmf.py:102-128) ----
    @property
    def W(self):
        return self.pyW.py_buf["val"]

    @property
    def H(self):
        return self.pyH.py_buf["val"]

    @property
    def lag_val(self):
        return self.pylag_val.py_buf["val"]

    @property
    def k(self):
        return self.W.shape[1]

    @property
    def m(self):
        return self.W.shape[0]

    @property
    def n(self):
        return self.H.shape[0]

    @staticmethod
    def _wrap(W, H, lag_val, dtype):
        return (PyMatrix(np.ascontiguousarray(W), dtype, major="row"),
                PyMatrix(np.ascontiguousarray(H), dtype, major="row"),
                PyMatrix(np.asfortranarray(lag_val), dtype, major="col"))

    # ---- checkpointing: arrays.npz + other.pkl, same file format (trmf.py:131-168) ----
    @classmethod
    def load(cls, path_to_folder, dtype=None):
        assert os.path.isdir(path_to_folder)
        with open(os.path.join(path_to_folder, "other.pkl"), "rb") as fh:
            transform = pickle.load(fh)["transform"]
        with np.load(os.path.join(path_to_folder, "arrays.npz")) as npz:
            W, H, lag_val, lag_set = npz["W"], npz["H"], npz["lag_val"], npz["lag_set"]
        if dtype is None:
            dtype = W.dtype
        pyW, pyH, pyL = cls._wrap(W, H, lag_val, dtype)
        return cls(pyW=pyW, pyH=pyH, pylag_val=pyL, lag_set=lag_set, transform=transform)

    def save(self, path_to_folder):
        if not os.path.exists(path_to_folder):
            os.makedirs(path_to_folder)
        else:
            assert os.path.isdir(path_to_folder)
        with open(os.path.join(path_to_folder, "arrays.npz"), "wb") as fh:
            np.savez(fh, W=self.W, H=self.H, lag_val=self.lag_val, lag_set=self.lag_set)
        with open(os.path.join(path_to_folder, "other.pkl"), "wb") as fh:
            pickle.dump({"transform": self.transform}, fh)

    # ---- forecasting (trmf.py:170-193): sequential AR roll-out, then Wnew H^T ----
    def latent_forecast(self, window, Wnew=None):
        m, k = self.m, self.k
        if Wnew is not None:
            assert Wnew.shape[0] == m + window
            assert Wnew.shape[1] == k
            assert Wnew.dtype == self.W.dtype
            assert Wnew.flags["C_CONTIGUOUS"]
        else:
            Wnew = np.zeros((m + window, k), dtype=self.W.dtype, order="C")
        Wnew[:m, :] = self.W
        lags = np.asarray(self.lag_set, dtype=np.int64)
        theta = self.lag_val
        for i in range(m, m + window):
            Wnew[i, :] = (Wnew[i - lags, :] * theta).sum(axis=0)
        return Wnew

    def forecast(self, window, Ynew=None, threshold=None):
        Wnew = self.latent_forecast(window)[self.m:, :]
        if Ynew is None:
            Ynew = np.zeros((window, self.n), dtype=self.W.dtype, order="C")
        Ynew[:] = Wnew @ self.H.T
        if threshold is not None:
            Ynew[Ynew < threshold] = threshold
        if self.transform is not None:
            Ynew[:] = self.transform.postprocess(Ynew)
        return Ynew, Wnew

    # ---- synthetic data in the model's own generative form (trmf.py:195-220) ----
    @staticmethod
    def syn_gen(m, n, k, lag_set, seed=None, noise=0.01, dtype=np.float32):
        if seed is not None:
            np.random.seed(seed)
        lag_set = np.array(sorted(lag_set), dtype=np.uint32)
        L = len(lag_set)
        midx = int(lag_set.max())
        W = np.zeros((m, k), dtype=dtype, order="C")
        H = np.zeros((n, k), dtype=dtype, order="C")
        lag_val = np.zeros((L, k), dtype=dtype, order="F")
        W[:] = np.random.randn(m, k)
        H[:] = np.random.randn(n, k)
        lag_val[:] = np.random.randn(L, k)
        lag_val = np.dot(lag_val, np.diag(1.0 / (np.abs(lag_val).sum(axis=0) + 0.1)))
        lags = lag_set.astype(np.int64)
        for i in range(midx, m):
            W[i, :] = (W[i - lags, :] * lag_val).sum(axis=0)
        W[midx:, :] += noise * np.random.randn(m - midx, k)
        Y = np.zeros((m, n), dtype=dtype, order="C")
        np.dot(W, H.T, out=Y)
        return {"W": W, "H": H, "lag_val": lag_val, "lag_set": lag_set, "Y": Y, "k": k}

    # ---- initialisation / warm start (trmf.py:222-251) ----
    @classmethod
    def initialize(cls, Y, lag_set, k, warm_start_model=None, seed=None, dtype=None, transform=None, _transform_stats=None):
        if dtype is None:
            dtype = Y.dtype
        m, n = Y.shape[0], Y.shape[1]
        lag_set = np.array(sorted(lag_set), dtype=np.uint32)
        L = len(lag_set)
        W = np.zeros((m, k), dtype=dtype, order="C")
        H = np.zeros((n, k), dtype=dtype, order="C")
        lag_val = np.zeros((L, k), dtype=dtype, order="F")
        # draw order W, H, lag_val after the seed -- the reference's stream (trmf.py:234-236), NumPy's global generator.
        # Seed + draws are one critical section: grid_search(workers > 1) initialises models from several threads.
        with _RNG_LOCK:
            if seed is not None:
                np.random.seed(seed)
            W[:] = np.random.rand(m, k)
            H[:] = np.random.rand(n, k)
            lag_val[:] = np.random.randn(L, k)
        if warm_start_model is not None:
            prev = warm_start_model
            assert prev.k == k
            assert prev.n == n
            assert prev.m <= m
            assert len(lag_set) == len(prev.lag_set)
            W[:] = prev.latent_forecast(m - prev.m)
            H[:] = prev.H
            lag_val[:] = prev.lag_val
            transform = prev.transform
        if transform is not None:
            transform = NormalizedTransform(Y, _stats=_transform_stats)
        pyW, pyH, pyL = cls._wrap(W, H, lag_val, dtype)
        return cls(pyW=pyW, pyH=pyH, pylag_val=pyL, lag_set=lag_set, transform=transform)


def train(Y, model, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1,
          max_iter=10, period_W=1, period_H=1, period_Lag=2,
          threads=1, missing=False, verbose=0):
    """ALS training on the GPU; updates ``model`` in place and returns it
    (reference trmf.py:253-264).  ``threads`` is accepted for compatibility."""
    if model.transform is not None:
        Y = model.transform.preprocess(Y)
    # one sparse orientation only: the library derives the other on the device (the reference's PyMatrix spends
    # seconds in scipy building both at the sizes of BASELINE config 2, rf_util.py:88-98)
    _clib.train(PyMatrix(Y, dtype=model.W.dtype, twin=False), model.lag_set,
                model.pyW, model.pyH, model.pylag_val, warm_start=True,
                lambdaI=lambdaI, lambdaAR=lambdaAR, lambdaLag=lambdaLag,
                max_iter=max_iter, period_W=period_W, period_H=period_H, period_Lag=period_Lag,
                threads=threads, missing=missing, verbose=verbose)
    return model


_METRIC_FIELDS = ["nd", "mase", "nrmse", "m_nd", "m_mase", "m_nrmse", "mape"]


class Metrics(collections.namedtuple("Metrics", _METRIC_FIELDS)):
    """ND / MASE / NRMSE, their per-series means, and MAPE (reference trmf.py:266-301)."""
    __slots__ = ()

    def __str__(self):
        return " ".join("{}={:.4g}".format(key, getattr(self, key)) for key in self._fields)

    @classmethod
    def default(cls):
        return cls(*([1e10] * len(_METRIC_FIELDS)))

    @classmethod
    def generate(cls, trueY, forecastY, missing=True):
        trueY = np.asarray(trueY)
        forecastY = np.asarray(forecastY)
        err = forecastY - trueY
        a_err, a_true = np.abs(err), np.abs(trueY)

        def finite_mean(x):
            x = x[np.isfinite(x)]
            assert len(x) != 0
            return x.mean()

        with np.errstate(divide="ignore", invalid="ignore"):
            nrmse = np.sqrt((err ** 2).mean()) / a_true.mean()
            m_nrmse = finite_mean(np.sqrt((err ** 2).mean(axis=0)) / a_true.mean(axis=0))
            nd = a_err.sum() / a_true.sum()
            m_nd = finite_mean(a_err.sum(axis=0) / a_true.sum(axis=0))
            naive = np.abs(trueY[1:, :] - trueY[:-1, :])
            mase = a_err.mean() / naive.mean()
            m_mase = finite_mean(a_err.mean(axis=0) / naive.mean(axis=0))
            nz = trueY != 0
            mape = finite_mean(a_err[nz] / a_true[nz])
        return cls(nd=nd, mase=mase, nrmse=nrmse, m_nd=m_nd, m_mase=m_mase, m_nrmse=m_nrmse, mape=mape)


def rolling_validate(Y, lag_set, k=40, window_size=24, nr_windows=7, lambdaI=0.5, lambdaAR=50, lambdaLag=0.5,
                     max_iter=20, missing=True, threshold=0, transform=None, threads=16, verbose=0, seed=0,
                     resident=None, _sessions=None, _models_out=None):
    """Rolling-origin evaluation (reference trmf.py:303-329): ``nr_windows``
    successive fits, each warm-started from the previous one and forecasting the
    next ``window_size`` time stamps.  ``missing=True`` turns exact zeros of the
    dense slice into unobserved entries (``csr_matrix(Y_trn)``).

    ``resident`` (additive): keep Y in HBM across the windows instead of
    converting and uploading the growing prefix for every fit (``_rolling_resident``
    below).  Same models, forecasts and metrics, bit for bit; default on for a
    dense float32/float64 ``Y`` when ``verbose == 0`` (the per-window path prints
    the reference's parameter dump), ``TRMF_B200_ROLLING_HOST=1`` turns it off."""
    T, n = Y.shape[0], Y.shape[1]
    assert T > nr_windows * window_size
    if resident is None:
        resident = (verbose == 0 and isinstance(Y, np.ndarray) and Y.dtype in (np.float32, np.float64)
                    and bool(__package__)      # (run as a script, like the reference allows: no package, no session module)
                    and not os.environ.get("TRMF_B200_ROLLING_HOST"))
    if resident:
        return _rolling_resident(Y, lag_set, k, window_size, nr_windows, lambdaI, lambdaAR, lambdaLag, max_iter, missing,
                                 threshold, transform, seed, sessions=_sessions, models_out=_models_out)
    horizon = nr_windows * window_size
    trueY = Y[-horizon:, :]
    forecastY = np.zeros((horizon, n), dtype=Y.dtype, order="C")
    prev_model = None
    for w in range(nr_windows):
        trn_end = T - (nr_windows - w) * window_size
        Y_trn = Y[:trn_end, :]
        if missing:
            Y_trn = smat.csr_matrix(Y_trn)
        curr_model = Model.initialize(Y_trn, lag_set, k, seed=seed, warm_start_model=prev_model, transform=transform)
        curr_model = train(Y_trn, curr_model, lambdaI=lambdaI, lambdaAR=lambdaAR, lambdaLag=lambdaLag,
                           max_iter=max_iter, missing=missing, threads=threads, verbose=verbose)
        curr_model.forecast(window_size, Ynew=forecastY[w * window_size:(w + 1) * window_size, :], threshold=threshold)
        if _models_out is not None:     # (tests: the fitted model of every window)
            _models_out.append(curr_model)
        prev_model = curr_model
    return Metrics.generate(trueY, forecastY, missing=missing)


def _rolling_resident(Y, lag_set, k, window_size, nr_windows, lambdaI, lambdaAR, lambdaLag, max_iter, missing,
                      threshold, transform, seed, models_out=None, sessions=None):
    """``rolling_validate`` with Y resident on the device (``session.RollingSession``).

    The host keeps doing what the reference's loop does around ``train`` --
    ``Model.initialize`` (RNG stream, AR roll-out warm start, a fresh
    ``NormalizedTransform`` per window: trmf.py:222-251) and ``Model.forecast``
    (trmf.py:184-193) -- on the same NumPy expressions, so every window starts
    from the same bits.  What no longer happens per window (or at all, on the
    host): ``csr_matrix(Y_trn)`` -- 2-4 s for a 10 000 x 10 000 array --,
    the PyMatrix conversions, the upload of Y and of the factors (W[:T_prev], H
    and lag_val are already in HBM, bit-identical to what would be sent; only the
    ``window_size`` rolled-out rows of W cross PCIe), session set-up.

    ``sessions`` (a dict owned by the caller, e.g. ``grid_search``): sessions are
    parked there under (k, missing, resident length, lag set) instead of being
    closed, so that further calls over the same Y with other regularisation
    weights / iteration counts reuse the resident copy."""
    from .session import RollingSession   # (session imports this module)
    T, n = Y.shape[0], Y.shape[1]
    horizon = nr_windows * window_size
    trueY = Y[-horizon:, :]
    forecastY = np.zeros((horizon, n), dtype=Y.dtype, order="C")
    Y_res = Y[:T - window_size, :]        # the longest prefix any window trains on
    key = (int(k), bool(missing), Y_res.shape[0], tuple(sorted(int(l) for l in lag_set)), np.dtype(Y.dtype).str)
    sess = sessions.pop(key, None) if sessions is not None else None
    if sess is None:
        if sessions:
            # a parked session under another key (other k / window_size / lag set) holds a full device copy of Y:
            # at most one stays resident, so a grid over k or window_size cannot exhaust HBM
            for other in list(sessions):
                sessions.pop(other).close()
        try:
            # dense in both modes: with missing=True the device keeps the non-zero cells, i.e. csr_matrix(Y_res)
            sess = RollingSession(Y_res, lag_set, k, missing=missing, dtype=Y.dtype)
        except RuntimeError as e:
            if "out of memory" not in str(e).lower():
                raise
            # Y does not fit next to what else lives on the device: the per-window path (= the reference's loop)
            return rolling_validate(Y, lag_set, k=k, window_size=window_size, nr_windows=nr_windows, lambdaI=lambdaI,
                                    lambdaAR=lambdaAR, lambdaLag=lambdaLag, max_iter=max_iter, missing=missing,
                                    threshold=threshold, transform=transform, seed=seed, resident=False,
                                    _models_out=models_out)
    sess.set_params(lambdaI, lambdaAR, lambdaLag)
    done = False
    try:
        prev_model = None
        for w in range(nr_windows):
            trn_end = T - (nr_windows - w) * window_size
            # the dense slice stands in for csr_matrix(Y_trn): initialize() only takes its shape, dtype and
            # NormalizedTransform statistics, which the reference computes on toarray() anyway (trmf.py:84)
            # the window statistics of NormalizedTransform (a fresh one per window, trmf.py:247-248) come from the device when it
            # holds the dense Y: same bits as Yd.mean / Yd.std, without two host passes over the whole prefix per window
            will_transform = (prev_model.transform if prev_model is not None else transform) is not None
            stats = sess.window_stats(trn_end) if will_transform else None
            curr_model = Model.initialize(Y[:trn_end, :], lag_set, k, seed=seed, warm_start_model=prev_model,
                                          transform=transform, _transform_stats=stats)
            tr = curr_model.transform
            sess.window(trn_end, None if tr is None else tr.a, None if tr is None else tr.b)
            if prev_model is None:
                sess.upload(W=curr_model.W, H=curr_model.H, lag_val=curr_model.lag_val)
            else:
                sess.upload_W_rows(prev_model.m, curr_model.W[prev_model.m:, :])
            sess.train(max_iter=max_iter, period_W=1, period_H=1, period_Lag=2, verbose=0)
            sess.download_into(curr_model.W, curr_model.H, curr_model.lag_val)
            curr_model.forecast(window_size, Ynew=forecastY[w * window_size:(w + 1) * window_size, :], threshold=threshold)
            if models_out is not None:
                models_out.append(curr_model)
            prev_model = curr_model
        done = True
    finally:
        if done and sessions is not None:
            sessions[key] = sess
        else:
            sess.close()
    return Metrics.generate(trueY, forecastY, missing=missing)


def grid_search(Y, lag_set, grid_params, pkl_file=None, workers=1, **kw_args):
    """Exhaustive search over ``grid_params`` (dict name -> list of values),
    ranking by ``m_nd`` (reference trmf.py:331-346).  Every grid point is one
    ``rolling_validate``; on the resident path (its default for a dense float
    array) consecutive grid points with the same (k, missing, resident length =
    T - window_size, lag set, dtype) share one copy of Y in HBM instead of
    re-ingesting it nr_windows times per point; at most one session stays parked,
    and a session that does not fit falls back to the per-window path.

    ``workers`` > 1 (an addition; the reference runs the grid serially) trains that
    many grid points at the same time: each worker thread owns its resident session
    (its own copy of Y in HBM and its own CUDA stream), so the small launch-bound
    kernels of the reference's data sets (370 x 26 304, 963 x 10 560) overlap on the
    device.  Every grid point computes exactly what it computes alone -- results, the
    best point and the print-out are the serial run's."""
    names = list(grid_params.keys())
    combos = list(itertools.product(*[grid_params[name] for name in names]))
    # Grid points that can share a resident session (same k, window_size, missing, lag set) run back to back, so that
    # ONE device copy of Y serves all of them and only one session is alive at a time; results, the running best and
    # its print-out are reported in the reference's order (trmf.py:336-345).
    def share_key(combo):
        kws = dict(kw_args)
        kws.update(zip(names, combo))
        return (repr(kws.get("k", 40)), repr(kws.get("window_size", 24)), repr(kws.get("missing", True)))
    order = sorted(range(len(combos)), key=lambda i: share_key(combos[i]))   # (stable: ties keep the grid's order)
    done = {}
    parked = []          # one dict of parked sessions per worker thread

    def run(i, sessions):
        kws = dict(kw_args)
        kws.update(zip(names, combos[i]))
        return {"kws": kws, "metrics": rolling_validate(Y, lag_set, _sessions=sessions, **kws)}

    def record(i, res):
        done[i] = res
        if pkl_file is not None:
            with open(pkl_file, "wb") as fh:
                pickle.dump([done[j] for j in sorted(done)], fh)

    try:
        if workers is None or workers <= 1 or len(order) <= 1:
            parked.append({})
            for i in order:
                record(i, run(i, parked[0]))
        else:
            import threading
            from concurrent.futures import ThreadPoolExecutor
            tls, lock = threading.local(), threading.Lock()

            def worker(i):
                if not hasattr(tls, "sessions"):
                    tls.sessions = {}
                    with lock:
                        parked.append(tls.sessions)
                return i, run(i, tls.sessions)

            with ThreadPoolExecutor(max_workers=int(workers)) as pool:
                for i, res in pool.map(worker, order):
                    record(i, res)
    finally:
        for sessions in parked:
            for sess in sessions.values():
                sess.close()
    results = [done[i] for i in range(len(combos))]
    best = Metrics.default()
    for r, combo in zip(results, combos):
        if r["metrics"].m_nd < best.m_nd:
            best = r["metrics"]
            print(best, dict(zip(names, combo)))
    return results, best


if __name__ == "__main__":
    # the reference's own smoke run (trmf.py:348-365): synthetic data, one fit, one rolling validation
    def norm(x):
        return (x * x).sum()

    missing = False
    m, n, k, lag_set = 1000, 500, 20, list(range(24)) + list(range(24 * 7, 24 * 8))
    dtype = np.float64
    threads = 16
    data = Model.syn_gen(m, n, k, lag_set, seed=0, dtype=dtype)
    data["Y"] += 10
    m0 = Model.initialize(data["Y"], data["lag_set"], k + 20, seed=0)
    print("dY={} W{} H{}".format(norm(data["Y"] - m0.W.dot(m0.H.T)), norm(m0.W), norm(m0.H)))
    train(data["Y"], m0, lambdaI=0.01, lambdaAR=0.001, lambdaLag=0.0001, max_iter=20, missing=missing, verbose=1, threads=threads)
    print("dY={} W{} H{}".format(norm(data["Y"] - m0.W.dot(m0.H.T)), norm(m0.W), norm(m0.H)))
    print(rolling_validate(data["Y"], data["lag_set"], k, 24, 7, lambdaI=0.001, lambdaAR=0.001, lambdaLag=0.2, threshold=None))

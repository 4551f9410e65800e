"""TEST INFRASTRUCTURE ONLY -- NumPy float64 restatement of the reference's ALS path.

This is the CPU oracle ("port") for the hot path named by BASELINE.json: the
three inner updates of ``c_trmf_train`` and the loop around them.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it; the product path (``exp-trmf-nips16_b200/``) never does.

PINNING.  The reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so this restatement is pinned against *outputs of the
reference itself run here*: ``oracle/_ref/trmf_float64.so`` is the unmodified
``python/trmf/corelib/trmf.cpp`` compiled by ``oracle/Makefile``;
``tests/golden/make_golden.py`` records its outputs on seeded inputs as
``tests/golden/*.npz`` and ``tests/test_oracle.py`` checks this file against
those vectors (and against the live ``_ref`` library when it is present).

Every function cites the reference lines it follows.  Shapes/names follow the
reference: Y is T x n (rows = time stamps), W is T x k ("X" of the paper),
H is n x k ("F" of the paper), lag_val (Theta) is L x k, lag_set sorted uint32.
Everything is float64 == the reference's ``ValueType=double`` build.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sps

# solver constants: trmf.h:90-93 (eps = eps_cg = 0.1, max_tron_iter = 2,
# max_cg_iter = 10) merged by trmf.cpp:603-606 into one Newton step of <= 20 CG.
EPS_CG = 0.1
MAX_CG = 20
ETA0 = 1e-4  # rf_tron.h:138


def _chol_solve(A, b):
    """LAPACK ?posv('U') as used by ls_solve_chol (rf_matrix.h:3008-3014)."""
    c = sla.cho_factor(A, lower=False, check_finite=False)
    return sla.cho_solve(c, b, check_finite=False)


# --------------------------------------------------------------------------
# F-update (H_solver): trmf.cpp:369-397 sparse, trmf.cpp:319-337 dense
# --------------------------------------------------------------------------
def f_update_sparse(Ycsc, W, H, lambdaI):
    """Per series j: (sum_{i in Omega_j} W_i W_i^T + lambdaI I) h_j = sum Y_ij W_i.

    Rows (series) with no observation keep their previous value (trmf.cpp:374).
    ``Ycsc`` is the CSC view of the T x n matrix Y, i.e. the CSR of Y^T that the
    reference obtains by pointer swap (rf_matrix.h:1633-1640).
    """
    Ycsc = sps.csc_matrix(Ycsc)
    n, k = H.shape
    H = H.copy()
    ptr, idx, val = Ycsc.indptr, Ycsc.indices, Ycsc.data
    eye = lambdaI * np.eye(k)
    for j in range(n):
        lo, hi = ptr[j], ptr[j + 1]
        if lo == hi:
            continue
        Wj = W[idx[lo:hi]]
        H[j] = _chol_solve(Wj.T @ Wj + eye, Wj.T @ val[lo:hi])
    return H


def f_update_dense(Y, W, lambdaI):
    """YH = Y^T W, HTH = W^T W + lambdaI I, one posv with n right-hand sides
    (trmf.cpp:319-337).  ``Y`` may be dense or sparse (gmat_x_dmat,
    rf_matrix.h:2853-2869: a sparse Y is multiplied as-is, zeros = observed 0)."""
    k = W.shape[1]
    YtW = np.asarray(Y.T @ W)
    G = W.T @ W + lambdaI * np.eye(k)
    return _chol_solve(G, YtW.T).T.copy()


# --------------------------------------------------------------------------
# X-update objective pieces
# --------------------------------------------------------------------------
def ar_residual(S, lag_set, lag_val):
    """rho[i,t] = S[i,t] - sum_l Theta[l,t] S[i-lag_l,t] for i >= max lag, else 0
    (trmf.cpp:82-90, 109-114)."""
    T = S.shape[0]
    mid = int(lag_set[-1])  # trmf.cpp:79 - last element of the sorted set
    rho = np.zeros_like(S)
    if mid < T:
        rho[mid:] = S[mid:]
        for l, lag in enumerate(lag_set):
            lag = int(lag)
            rho[mid:] -= lag_val[l] * S[mid - lag:T - lag]
    return rho, mid


def base_fun(W, lag_set, lag_val, lambdaI, lambdaAR):
    """arr_base_IX::fun, trmf.cpp:70-97."""
    f = 0.0
    if lambdaI > 0:
        f += 0.5 * lambdaI * float(np.vdot(W, W))
    if lambdaAR > 0:
        rho, _ = ar_residual(W, lag_set, lag_val)
        f += 0.5 * lambdaAR * float(np.vdot(rho, rho))
    return f


def base_apply(S, lag_set, lag_val, lambdaI, lambdaAR):
    """arr_base_IX::grad / ::Hv (identical linear map), trmf.cpp:99-149:
    out = lambdaI S + lambdaAR A^T A S."""
    out = lambdaI * S
    if lambdaAR > 0:
        T = S.shape[0]
        rho, mid = ar_residual(S, lag_set, lag_val)
        if mid < T:
            out[mid:] += lambdaAR * rho[mid:]
            for l, lag in enumerate(lag_set):
                lag = int(lag)
                out[mid - lag:T - lag] -= lambdaAR * rho[mid:] * lag_val[l]
    return out


class SparseLoss:
    """arr_ls_pY_IX, trmf.cpp:220-289 (Y as CSR by time stamp)."""

    def __init__(self, Ycsr, H):
        Ycsr = sps.csr_matrix(Ycsr)
        self.ptr, self.col, self.val = Ycsr.indptr, Ycsr.indices, Ycsr.data
        self.shape = Ycsr.shape
        self.row = np.repeat(np.arange(Ycsr.shape[0]), np.diff(Ycsr.indptr))
        self.H = H

    def _scatter(self, z):
        Z = sps.csr_matrix((z, self.col, self.ptr), shape=self.shape)
        return np.asarray(Z @ self.H)

    def fun(self, W):  # trmf.cpp:231-245
        r = self.val - np.einsum("ek,ek->e", W[self.row], self.H[self.col])
        return 0.5 * float(np.dot(r, r))

    def grad(self, W):  # trmf.cpp:247-267
        r = np.einsum("ek,ek->e", W[self.row], self.H[self.col]) - self.val
        return self._scatter(r)

    def Hv(self, S):  # trmf.cpp:269-288
        z = np.einsum("ek,ek->e", S[self.row], self.H[self.col])
        return self._scatter(z)


class DenseLoss:
    """arr_ls_fY_IX, trmf.cpp:155-215: YH = Y H, HTH = H^T H precomputed in init()."""

    def __init__(self, Y, H):
        self.trYTY = float(Y.multiply(Y).sum()) if sps.issparse(Y) else float(np.vdot(Y, Y))
        self.YH = np.asarray(Y @ H)
        self.HTH = H.T @ H

    def fun(self, W):  # trmf.cpp:189-197
        return 0.5 * self.trYTY + 0.5 * float(np.vdot(W.T @ W, self.HTH)) - float(np.vdot(self.YH, W))

    def grad(self, W):  # trmf.cpp:199-207
        return -self.YH + W @ self.HTH

    def Hv(self, S):  # trmf.cpp:209-214
        return S @ self.HTH


def x_update(loss, W, lag_set, lag_val, lambdaI, lambdaAR, info=None):
    """One TRON step with pure CG (rf_tron.h:135-254, trcg 412-505), as driven
    by arr_solver::solve (trmf.h:175-186) with max_iter = 1, warm start."""
    def fun(V):
        return loss.fun(V) + base_fun(V, lag_set, lag_val, lambdaI, lambdaAR)

    def grad(V):  # base first, then the loss term (trmf.cpp:248, 102-121, 252-266)
        return base_apply(V, lag_set, lag_val, lambdaI, lambdaAR) + loss.grad(V)

    def Hv(V):
        return base_apply(V, lag_set, lag_val, lambdaI, lambdaAR) + loss.Hv(V)

    max_cg = min(MAX_CG, W.size)  # trmf.cpp:523-526
    f = fun(W)
    g = grad(W)
    gnorm = np.sqrt(np.vdot(g, g))
    if not gnorm > 0:  # rf_tron.h:170: gnorm <= eps*gnorm1 only when gnorm == 0 (or NaN guard)
        if info is not None:
            info.update(cg_iter=0, accepted=False, f=f, fnew=f)
        return W.copy()
    # trcg, rf_tron.h:412-505
    s = np.zeros_like(W)
    r = -g
    d = r.copy()
    cgtol = EPS_CG * gnorm
    rTr = float(np.vdot(r, r))
    cg_iter = 0
    while True:
        if np.sqrt(np.vdot(r, r)) <= cgtol:
            break
        if cg_iter >= max_cg:
            break
        cg_iter += 1
        Hd = Hv(d)
        alpha = rTr / float(np.vdot(d, Hd))
        s += alpha * d
        r -= alpha * Hd
        rnew = float(np.vdot(r, r))
        beta = rnew / rTr
        d = d + (beta - 1.0) * d  # rf_tron.h:495-499
        d = d + r                 # rf_tron.h:500-501
        rTr = rnew
    W_new = W + s
    gs = float(np.vdot(g, s))
    prered = -0.5 * (gs - float(np.vdot(s, r)))
    fnew = fun(W_new)
    actred = f - fnew
    accepted = actred > ETA0 * prered  # rf_tron.h:222
    if info is not None:
        info.update(cg_iter=cg_iter, accepted=bool(accepted), f=f, fnew=fnew,
                    actred=actred, prered=prered, gnorm=float(gnorm))
    return W_new if accepted else W.copy()


# --------------------------------------------------------------------------
# Theta (lag_val) update: trmf.cpp:455-484
# --------------------------------------------------------------------------
def lag_update(W, lag_set, lambdaLag):
    """Per latent dim t: ridge regression of x_i on (x_{i-lag_l})_l over the
    window i in [max lag, T)."""
    T, k = W.shape
    L = len(lag_set)
    mid = int(lag_set[-1])
    out = np.zeros((L, k))
    for t in range(k):
        x = W[:, t]
        X = np.stack([x[mid - int(lag):T - int(lag)] for lag in lag_set], axis=1) if mid < T else np.zeros((0, L))
        G = X.T @ X + lambdaLag * np.eye(L)
        out[:, t] = _chol_solve(G, X.T @ x[mid:])
    return out


# --------------------------------------------------------------------------
# The ALS loop: trmf.cpp:647-693
# --------------------------------------------------------------------------
def train(Y, lag_set, W, H, lag_val, lambdaI=0.1, lambdaAR=0.1, lambdaLag=0.1, max_iter=10,
          period_W=1, period_H=1, period_Lag=2, missing=True, trace=None):
    """Order F -> X -> Theta; a phase runs iff iter % period == 0, iter from 1."""
    W = np.array(W, dtype=np.float64)
    H = np.array(H, dtype=np.float64)
    lag_val = np.array(lag_val, dtype=np.float64)
    lag_set = np.sort(np.asarray(lag_set)).astype(np.int64)
    if missing:
        Ycsr = sps.csr_matrix(Y)
        Ycsc = sps.csc_matrix(Y)
    for it in range(1, max_iter + 1):
        if it % period_H == 0:
            H = f_update_sparse(Ycsc, W, H, lambdaI) if missing else f_update_dense(Y, W, lambdaI)
        if it % period_W == 0:
            loss = SparseLoss(Ycsr, H) if missing else DenseLoss(Y, H)
            info = {}
            W = x_update(loss, W, lag_set, lag_val, lambdaI, lambdaAR, info)
            if trace is not None:
                trace.append(info)
        if it % period_Lag == 0:
            lag_val = lag_update(W, lag_set, lambdaLag)
    return W, H, lag_val

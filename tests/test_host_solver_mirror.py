"""Host-side logic of krylov_b200.solvers (the mirror of the reference's RecyclingKrylovSolvers
signatures that the Julia shim follows) checked on the CPU: a numpy stand-in implements the device
block API of capi.Handle (kb200_mat_* / kb200_op_* / kb200_pcg_burst) for an explicit SPD matrix, so
the control flow of eigpcg / eigdefpcg / defpcg_hostloop -- restarts, the bookkeeping of V'AV from
the alpha / beta history of a device burst, column indices, stopping rules -- is compared with the
oracle's restatement (oracle/solvers.py) without a GPU.  TEST TOOLING: the product path has no CPU
implementation (capi.lib() raises when libkrylov_b200.so cannot run)."""
import numpy as np
import pytest

import oracle.solvers as OS
from krylov_b200 import solvers


class NumpyBlockHandle:
    """numpy twin of the device block API; operator A (which 0/1), preconditioner Minv (which 2)."""

    def __init__(self, A, Minv):
        self.A, self.Minv = A, Minv
        self.n_gamma = A.shape[0]
        self.mats = {}
        self.next_id = 0
        self.bursts = 0

    def mat_alloc(self, ncols):
        self.mats[self.next_id] = np.zeros((self.n_gamma, ncols), order="F")
        self.next_id += 1
        return self.next_id - 1

    def mat_free(self, mid):
        del self.mats[mid]

    def mat_set(self, mid, col0, data):
        data = np.asarray(data, dtype=np.float64)
        if data.ndim == 1:
            self.mats[mid][:, col0] = data
        else:
            self.mats[mid][:, col0:col0 + data.shape[1]] = data

    def mat_get(self, mid, col0, ncols=1):
        return self.mats[mid][:, col0].copy() if ncols == 1 else self.mats[mid][:, col0:col0 + ncols].copy()

    def op_apply(self, which, src, scol, dst, dcol, ncols=1):
        X = self.mats[src][:, scol:scol + ncols].copy()
        self.mats[dst][:, dcol:dcol + ncols] = (self.Minv @ X) if which == 2 else (self.A @ X)

    def op_axpby(self, a, x, xcol, b, y, ycol):
        yv = self.mats[y][:, ycol]
        self.mats[y][:, ycol] = (0.0 if b == 0.0 else b * yv) + a * self.mats[x][:, xcol]

    def op_dot(self, x, xcol, y, ycol):
        return float(self.mats[x][:, xcol] @ self.mats[y][:, ycol])

    def op_gemv_t(self, A, col0, k, x, xcol):
        return self.mats[A][:, col0:col0 + k].T @ self.mats[x][:, xcol]

    def op_gemv_n(self, a, A, col0, k, mu, b, y, ycol):
        yv = self.mats[y][:, ycol]
        self.mats[y][:, ycol] = (0.0 if b == 0.0 else b * yv) + a * (self.mats[A][:, col0:col0 + k] @ np.asarray(mu))

    def op_gemm_tn(self, A, ca, ka, B, cb, kb):
        return self.mats[A][:, ca:ca + ka].T @ self.mats[B][:, cb:cb + kb]

    def op_gemm_nn(self, A, ca, ka, Q, Cm, cc):
        Q = np.asarray(Q)
        self.mats[Cm][:, cc:cc + Q.shape[1]] = self.mats[A][:, ca:ca + ka] @ Q

    def pcg_burst(self, assembled, precond, wk, m, rTz, tol, it0, maxit, V=-1, vcol0=0, W=-1, AW=-1, nvec=0,
                  WtAW=None, WtW=None):
        """Semantics of kb200_pcg_burst (include/krylov_b200.h): columns 0..4 of `wk` are x, r, z, p, Ap."""
        self.bursts += 1
        S = self.mats[wk]
        X, R, Z, P, AP = 0, 1, 2, 3, 4
        al, be, rs = [], [], []
        it = it0
        m = min(m, maxit - it0)
        for j in range(m):
            S[:, AP] = self.A @ S[:, P]
            alpha = rTz / float(S[:, P] @ S[:, AP])
            S[:, X] += alpha * S[:, P]
            S[:, R] -= alpha * S[:, AP]
            if WtW is not None:
                Wm = self.mats[W][:, :nvec]
                S[:, R] -= Wm @ np.linalg.solve(WtW, Wm.T @ S[:, R])
            rTr = float(S[:, R] @ S[:, R])
            S[:, Z] = (self.Minv @ S[:, R]) if precond else S[:, R]
            rTz_new = float(S[:, R] @ S[:, Z])
            beta = rTz_new / rTz
            S[:, P] = beta * S[:, P] + S[:, Z]
            if nvec:
                S[:, P] -= self.mats[W][:, :nvec] @ np.linalg.solve(WtAW, self.mats[AW][:, :nvec].T @ S[:, Z])
            rTz = rTz_new
            it += 1
            if V >= 0:
                self.mats[V][:, vcol0 + j] = S[:, Z] / np.sqrt(rTz)
            al.append(alpha)
            be.append(beta)
            rs.append(np.sqrt(rTr))
            if not (it < maxit and np.sqrt(rTr) > tol):
                break
        return len(al), np.array(al), np.array(be), np.array(rs), rTz


class _Op:
    def __init__(self, h):
        self.h, self.assembled, self.N = h, False, h.n_gamma


class _Pre:
    def __init__(self, S):
        self.S = S


def _problem(n=140, seed=0, nsmall=8):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.linspace(1.0, 4.0, nsmall), np.linspace(150.0, 400.0, n - nsmall)])
    A = (Q * lam) @ Q.T
    d = 1.0 + rng.random(n)
    Minv = np.diag(1.0 / d)                           # SPD diagonal preconditioner
    b = rng.standard_normal(n)
    return A, Minv, b, Q


@pytest.mark.parametrize("burst", [True, False])
def test_eigpcg_and_eigdefpcg_mirror_the_oracle(burst):
    A, Minv, b, Q = _problem()
    n, nvec, spdim = A.shape[0], 6, 18
    h = NumpyBlockHandle(A, Minv)
    S = _Op(h)
    M = _Pre(S)
    Aop, Mop = (lambda v: A @ v), (lambda r: Minv @ r)
    xo, ito, reso, Vo = OS.eigpcg(Aop, b, np.zeros(n), Mop, nvec, spdim)
    x, it, res, V = solvers.eigpcg_hostloop(S, b, np.zeros(n), M, nvec, spdim, burst=burst)
    assert it == ito and V.shape == (n, nvec)
    assert np.abs(res - reso).max() <= 1e-9 * reso.max()
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
    assert np.abs(np.abs(np.linalg.svd(np.linalg.qr(V)[0].T @ np.linalg.qr(Vo)[0], compute_uv=False)) - 1).max() <= 1e-6
    assert (h.bursts > 0) == burst
    # recycle on a second right-hand side (Example09 / 11 usage)
    b2 = np.random.default_rng(9).standard_normal(n)
    x2o, it2o, res2o, V2o = OS.eigdefpcg(Aop, b2, np.zeros(n), Mop, Vo, spdim)
    h.bursts = 0
    x2, it2, res2, V2 = solvers.eigdefpcg_hostloop(S, b2, np.zeros(n), M, Vo, spdim, burst=burst)
    assert it2 == it2o and np.abs(res2 - res2o).max() <= 1e-8 * res2o.max()
    assert np.linalg.norm(x2 - x2o) <= 1e-8 * np.linalg.norm(x2o)
    assert (h.bursts > 0) == burst
    _, it_plain, _ = OS.pcg(Aop, b2, np.zeros(n), Mop)
    assert it2 < it_plain
    assert not h.mats                                  # every device block was released


def test_burst_respects_maxit_and_restart_boundaries():
    A, Minv, b, Q = _problem(seed=3)
    n, nvec, spdim = A.shape[0], 4, 10
    h = NumpyBlockHandle(A, Minv)
    S = _Op(h)
    M = _Pre(S)
    # (maxit = 5 would end with 4 basis vectors and nvec = 4: the final Rayleigh-Ritz of the reference then
    # indexes past its 3 x 3 minor, eigcg.jl:251-256 -- not a case it supports)
    for maxit in (2, 7, 9, 10, 11, 23):
        xo, ito, reso, Vo = OS.eigpcg(lambda v: A @ v, b, np.zeros(n), lambda r: Minv @ r, nvec, spdim, maxit=maxit)
        x, it, res, V = solvers.eigpcg_hostloop(S, b, np.zeros(n), M, nvec, spdim, maxit=maxit)
        assert it == ito == min(maxit, it) and res.size == it
        assert np.abs(res - reso).max() <= 1e-9 * reso.max()


def test_defpcg_hostloop_mirror():
    A, Minv, b, Q = _problem(seed=5)
    n = A.shape[0]
    h = NumpyBlockHandle(A, Minv)
    S = _Op(h)
    M = _Pre(S)
    W = Q[:, :8]
    xo, ito, reso = OS.defpcg(lambda v: A @ v, b, np.zeros(n), W, lambda r: Minv @ r)
    x, it, res = solvers.defpcg_hostloop(S, b, np.zeros(n), W, M)
    assert it == ito and np.abs(res - reso).max() <= 1e-9 * reso.max()
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)

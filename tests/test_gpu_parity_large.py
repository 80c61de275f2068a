"""Parity of the CUDA path at the sizes BASELINE.json names (configs 2-4): 501^2 / 64 subdomains (Example07),
1001^2 / 128 subdomains with nvec = 1.25 ndom, spdim = 3 ndom (Example17-style deflated solvers), 64 right-hand
sides (Example11).  The checker is the oracle's C port (oracle/kb_oracle.c through oracle.cpu_model) driving the
oracle's solver loops (oracle/solvers.py).

Tolerances: the ROUNDING-FLOOR protocol.  One-level Neumann-Neumann PCG on a lognormal field runs hundreds of
iterations of an ill-conditioned recurrence; two runs of the SAME arithmetic that differ only in rounding drift
apart exponentially with the iteration index.  The floor is measured, not assumed: the CPU port against itself
with the interface sums taken in descending instead of ascending subdomain order.  The GPU is held, index by
index, to a small multiple of that floor (and to 1e-10 where the floor is below it), and its iteration count to
the floor's own spread + 1."""
import os

import numpy as np
import pytest

import oracle.solvers as OS
from conftest import get_problem
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
from oracle.cpu_model import CpuModel

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1
_cache = {}


def setup(nx, ndom):
    key = (nx, ndom)
    if key not in _cache:
        pr = get_problem(nx, ndom)
        pr0 = get_problem(nx, ndom, ireal=0, median=True)          # constant "median" preconditioner (Example07:89, 152)
        cm0 = CpuModel(pr0.sd, pr0.blocks, pr0.factors)
        Pi = [PB.pinv_sym(S) for S in cm0.assemble_local_schurs(NT)]
        cm = CpuModel(pr.sd, pr.blocks, pr.factors, Pi)
        bs = cm.schur_rhs(pr.blocks.b_I, pr.blocks.b_G, NT)
        S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors)
        M = fem.B200NeumannNeumann(S, Pi)
        _cache[key] = (pr, cm, Pi, bs, S, M)
    return _cache[key]


def hist_diff(a, b, upto):
    return float(np.abs(a[:upto] - b[:upto]).max() / b[:upto].max())


def marks(m):
    return [i for i in (5, 10, 20, 50, 100, 200, 400, 800) if i < m] + [m]


def check_history(res_gpu, res_cpu, res_floor, m, what):
    """GPU-vs-CPU distance of the residual histories within 20 x the CPU-vs-CPU floor at every mark."""
    for i in marks(m):
        g, f = hist_diff(res_gpu, res_cpu, i), hist_diff(res_floor, res_cpu, i)
        assert g <= max(1e-10, 20 * f), f"{what}: history differs by {g:.2e} over the first {i} iterations, floor {f:.2e}"


def test_pcg_parity_501_64():
    """Config 2 (Example07, ~250k DOF, 64 subdomains, median preconditioner): iteration count and iterates."""
    pr, cm, Pi, bs, S, M = setup(501, 64)
    n = pr.sd.n_gamma
    xo, ito, reso = cm.pcg(bs, np.zeros(n), nthreads=NT)
    cm.set_reverse_order(True)
    xr, itr, resr = cm.pcg(bs, np.zeros(n), nthreads=NT)
    cm.set_reverse_order(False)
    x, it, res = solvers.pcg(S, bs, np.zeros(n), M)
    assert res[-1] <= 1e-7 * np.linalg.norm(bs)
    assert abs(it - ito) <= abs(itr - ito) + 1, f"it gpu {it}, cpu {ito}, cpu reordered {itr}"
    m = min(it, ito, itr)
    check_history(res, reso, resr, m, "pcg 501/64")
    # iterates at equal iteration index, early (tight) and at the common end index (floor-limited)
    for k in (20, m):
        xg, _, _ = solvers.pcg(S, bs, np.zeros(n), M, maxit=k)
        xc, _, _ = cm.pcg(bs, np.zeros(n), maxit=k, nthreads=NT)
        cm.set_reverse_order(True)
        xf, _, _ = cm.pcg(bs, np.zeros(n), maxit=k, nthreads=NT)
        cm.set_reverse_order(False)
        floor = np.linalg.norm(xf - xc) / np.linalg.norm(xc)
        assert np.linalg.norm(xg - xc) <= max(1e-10, 20 * floor) * np.linalg.norm(xc), (k, floor)
    # operators at this size against the C port
    v = np.random.default_rng(0).standard_normal(n)
    yc, zc = cm.apply_schur(v, NT), cm.apply_nn(v, NT)
    assert np.abs(S * v - yc).max() <= 1e-12 * np.abs(yc).max()
    assert np.abs(M.solve(v) - zc).max() <= 1e-12 * np.abs(zc).max()


def test_multi_rhs_64_right_hand_sides():
    """Config 4 (Example11: constant operator, 64 right-hand sides b = rand(n), seed 481456, Ex11 Functions:145):
    the block of independent recurrences against sequential single solves, matrix-free and assembled operator."""
    pr, cm, Pi, bs, S, M = setup(501, 64)
    n = pr.sd.n_gamma
    B = np.random.default_rng(481456).random((n, 64))
    X, its, ress = S.h.pcg_multi(B, np.zeros((n, 64)))
    for j in (0, 17, 63):
        xs, its_j, res_j = solvers.pcg(S, B[:, j].copy(), np.zeros(n), M)
        xo, ito, reso = cm.pcg(B[:, j], np.zeros(n), nthreads=NT)
        cm.set_reverse_order(True)
        xr, itr, resr = cm.pcg(B[:, j], np.zeros(n), nthreads=NT)
        cm.set_reverse_order(False)
        spread = abs(itr - ito) + 1
        assert abs(its[j] - its_j) <= spread and abs(its[j] - ito) <= spread
        assert ress[j][-1] <= 1e-7 * np.linalg.norm(B[:, j])
        check_history(ress[j], reso, resr, min(its[j], ito, itr), f"rhs {j}")
        assert np.linalg.norm(X[:, j] - xo) <= 5e-6 * np.linalg.norm(xo)       # both converged to the 1e-7 tolerance
    S.assemble_local_schurs()
    Xa, its_a, _ = S.h.pcg_multi(B, np.zeros((n, 64)), assembled=True)
    assert np.abs(its_a - its).max() <= 3
    assert np.linalg.norm(Xa - X) <= 5e-6 * np.linalg.norm(X)


def test_recycling_solvers_1001_128():
    """Config 3 (Example17-style: eigpcg on the first system, eigdefpcg with the recycled basis on the next;
    ~1M DOF, 128 subdomains, nvec = floor(1.25 ndom) = 160, spdim = 3 ndom = 384, Example09:39-40)."""
    pr, cm, Pi, bs, S, M = setup(1001, 128)
    n = pr.sd.n_gamma
    nvec, spdim = 160, 384
    Aop = lambda v: cm.apply_schur(v, NT)
    Mop = lambda r: cm.apply_nn(r, NT)
    xo, ito, reso, Vo = OS.eigpcg(Aop, bs, np.zeros(n), Mop, nvec, spdim)
    cm.set_reverse_order(True)
    xr, itr, resr, Vr = OS.eigpcg(Aop, bs, np.zeros(n), Mop, nvec, spdim)
    cm.set_reverse_order(False)
    x, it, res, V = solvers.eigpcg(S, bs, np.zeros(n), M, nvec, spdim)
    assert V.shape == (n, nvec) and res[-1] <= 1e-7 * np.linalg.norm(bs)
    assert abs(it - ito) <= abs(itr - ito) + 1, f"eigpcg it gpu {it}, cpu {ito}, cpu reordered {itr}"
    check_history(res, reso, resr, min(it, ito, itr, spdim - 1), "eigpcg before the first restart")
    floor_x = np.linalg.norm(xr - xo) / np.linalg.norm(xo)
    assert np.linalg.norm(x - xo) <= max(5e-6, 20 * floor_x) * np.linalg.norm(xo)
    # the recycled basis deflates the next solve: same W (the GPU's) for the GPU and the oracle
    b2 = Aop(np.random.default_rng(3).standard_normal(n))
    x2o, it2o, res2o, _ = OS.eigdefpcg(Aop, b2, np.zeros(n), Mop, V, spdim)
    cm.set_reverse_order(True)
    x2r, it2r, res2r, _ = OS.eigdefpcg(Aop, b2, np.zeros(n), Mop, V, spdim)
    cm.set_reverse_order(False)
    x2, it2, res2, V2 = solvers.eigdefpcg(S, b2, np.zeros(n), M, V, spdim)
    _, it_plain, _ = solvers.pcg(S, b2, np.zeros(n), M)
    assert res2[-1] <= 1e-7 * np.linalg.norm(b2) and V2.shape == (n, nvec)
    assert abs(it2 - it2o) <= abs(it2r - it2o) + 2, f"eigdefpcg it gpu {it2}, cpu {it2o}, cpu reordered {it2r}"
    assert it2 < it_plain, "deflation with the recycled basis must pay"
    floor2 = np.linalg.norm(x2r - x2o) / np.linalg.norm(x2o)
    assert np.linalg.norm(x2 - x2o) <= max(5e-6, 20 * floor2) * np.linalg.norm(x2o)
    # defpcg with the same basis (a12 at this size)
    x3o, it3o, res3o = OS.defpcg(Aop, b2, np.zeros(n), V, Mop)
    x3, it3, res3 = solvers.defpcg(S, b2, np.zeros(n), V, M)
    assert abs(it3 - it3o) <= abs(it2r - it2o) + 2 and res3[-1] <= 1e-7 * np.linalg.norm(b2)
    assert np.linalg.norm(x3 - x3o) <= max(5e-6, 20 * floor2) * np.linalg.norm(x3o)

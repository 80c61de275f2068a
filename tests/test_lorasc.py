"""LORASC preconditioner of the full system (EllipticPdeDomainDecomposition.jl:1406-1428, :1502-1676,
:1908-1976).  CPU: the oracle's literal apply against an independent dense block formula.  GPU: kb200_apply_lorasc
(through fem.B200Lorasc) against the oracle, alone and as the preconditioner of a PCG solve of the undecomposed system
(the way Example09 uses it)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle.epdd as E
import oracle.lorasc as OL
import oracle.solvers as OS
from conftest import get_oracle_ops, get_problem


def build(nx, ndom, nvec=25, eps=0.01):
    pr = get_problem(nx, ndom)
    sd = pr.sd
    inner, Sd, Pi, bs = get_oracle_ops(pr)
    A_II, A_IGg, A_GG, b_I, b_G = E.prepare_global_schur(pr.cells, pr.points, pr.epart, sd, pr.coeff, pr.f_vals, pr.u_dir)
    n_g = sd.n_gamma
    S = np.column_stack([E.apply_global_schur(A_II, A_IGg, A_GG, inner, e) for e in np.eye(n_g)])
    free = np.nonzero(pr.markers == 0)[0]
    P = OL.prepare_lorasc_precond(S, A_II, A_IGg, A_GG, sd.node_Id, sd.node_gamma, free, nvec=nvec, eps=eps)
    return pr, S, A_II, A_IGg, A_GG, free, P


def test_oracle_apply_equals_the_block_formula():
    """u = [I, -A_II^-1 A_IΓ; 0, I] [A_II^-1, 0; 0, A_ΓΓ^-1 + E E'] [I, 0; -A_ΓI A_II^-1, I] x, assembled densely."""
    pr, S, A_II, A_IGg, A_GG, free, P = build(33, 4, nvec=10, eps=0.5)
    sd = pr.sd
    assert 0 < P.E.shape[1] <= 10
    # generalized pairs: S e = sigma A_ΓΓ e with A_ΓΓ-orthonormal vectors, ascending
    G = P.E.T @ (A_GG @ P.E)
    assert np.abs(G - np.eye(P.E.shape[1])).max() <= 1e-10
    n = free.size
    g2l = np.full(sd.nnode, -1)
    g2l[free] = np.arange(n)
    idx_I = [g2l[v] for v in sd.node_Id]
    idx_G = g2l[sd.node_gamma]
    Minv = np.zeros((n, n))
    AiiInv = [np.linalg.inv(A.toarray()) for A in A_II]
    Sg = np.linalg.inv(A_GG.toarray()) + P.E @ P.E.T
    # dense operator columns
    X = np.eye(n)
    U = np.zeros((n, n))
    zG = X[idx_G, :].copy()
    xI = []
    for d in range(sd.ndom):
        t = AiiInv[d] @ X[idx_I[d], :]
        xI.append(t)
        zG -= A_IGg[d].T @ t
    xG = Sg @ zG
    for d in range(sd.ndom):
        U[idx_I[d], :] = xI[d] - AiiInv[d] @ (A_IGg[d] @ xG)
    U[idx_G, :] = xG
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.standard_normal(n)
        u = OL.apply_lorasc(P, x)
        assert np.abs(u - U @ x).max() <= 1e-11 * np.abs(u).max()
    assert np.abs(U - U.T).max() <= 1e-10 * np.abs(U).max()          # a symmetric preconditioner
    assert np.linalg.eigvalsh((U + U.T) / 2).min() > 0                # positive definite


def test_oracle_lorasc_preconditions_the_full_system():
    pr, S, A_II, A_IGg, A_GG, free, P = build(33, 4, nvec=10, eps=0.5)
    A, b, free2 = E.do_isotropic_elliptic_assembly(pr.cells, pr.points, pr.markers == 1, pr.coeff, pr.f_vals, pr.u_dir)
    assert np.array_equal(free, free2)
    x0 = np.zeros(free.size)
    _, it_pre, _ = OS.pcg(lambda v: A @ v, b, x0.copy(), lambda r: OL.apply_lorasc(P, r))
    _, it_cg, _ = OS.pcg(lambda v: A @ v, b, x0.copy(), lambda r: r.copy())
    assert it_pre < it_cg / 2


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ndom,devices", [(65, 16, None), (97, 9, None), (65, 16, [0])])
def test_gpu_apply_lorasc_matches_oracle(nx, ndom, devices):
    from krylov_b200 import fem
    from krylov_b200.hostprep import chol
    pr, S, A_II, A_IGg, A_GG, free, P = build(nx, ndom)
    sd = pr.sd
    Sop = fem.B200SchurOperator(sd, pr.blocks, pr.factors, devices=devices) if devices else fem.B200SchurOperator(sd, pr.blocks, pr.factors)
    pts = pr.points[:, sd.node_gamma]
    perm = chol.nd_order(A_GG, pts[0], pts[1])
    L, _ = chol.cholesky(A_GG, perm)
    M = fem.B200Lorasc(Sop, (perm, L), P.E, free)
    rng = np.random.default_rng(1)
    for _ in range(3):
        x = rng.standard_normal(free.size)
        u, uo = M.solve(x), OL.apply_lorasc(P, x)
        assert np.abs(u - uo).max() <= 1e-11 * np.abs(uo).max()
    # as the preconditioner of PCG on the undecomposed system (Example09:42): same iteration count, same iterate
    A, b, _ = E.do_isotropic_elliptic_assembly(pr.cells, pr.points, pr.markers == 1, pr.coeff, pr.f_vals, pr.u_dir) \
        if nx <= 65 else (None, None, None)
    if A is not None:
        xo, ito, reso = OS.pcg(lambda v: A @ v, b, np.zeros(free.size), lambda r: OL.apply_lorasc(P, r))
        xg, itg, resg = OS.pcg(lambda v: A @ v, b, np.zeros(free.size), M.solve)
        assert itg == ito and np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)
        assert np.abs(resg - reso).max() <= 1e-9 * reso.max()
    # E = nothing (ε = 0, :1623-1626): the block-factorization preconditioner with A_ΓΓ alone
    M0 = fem.B200Lorasc(Sop, (perm, L), np.zeros((sd.n_gamma, 0)), free)
    P0 = OL.prepare_lorasc_precond(S, A_II, A_IGg, A_GG, sd.node_Id, sd.node_gamma, free, eps=0.0)
    x = rng.standard_normal(free.size)
    assert np.abs(M0.solve(x) - OL.apply_lorasc(P0, x)).max() <= 1e-11 * np.abs(x).max()

"""The reference's own in-script self-checks, restated as tests of the oracle:
(i) global-column operator == local operator (Example03:175); (ii) DD solution == undecomposed
solve (Example03:204); (iii) sum_d R_d' A_GG,d R_d == A_GG; (iv) S SPD, S_d singular exactly for
floating subdomains; plus the committed golden vectors."""
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import oracle.epdd as E
import oracle.solvers as OS
from conftest import get_oracle_ops, get_problem

GOLD = os.path.join(os.path.dirname(__file__), "golden", "schur_nx33_ndom4.npz")


def test_global_and_local_operators_agree(medium_problem):
    pr = medium_problem
    inner, Sd, Pi, bs = get_oracle_ops(pr)
    A_II, A_IGg, A_GG, b_I, b_G = E.prepare_global_schur(pr.cells, pr.points, pr.epart, pr.sd, pr.coeff,
                                                         pr.f_vals, pr.u_dir)
    y1 = E.apply_global_schur(A_II, A_IGg, A_GG, inner, bs)
    y2 = E.apply_local_schurs_assembled(Sd, pr.sd, bs)
    y3 = E.apply_local_schurs(pr.blocks, inner, pr.sd, bs)
    assert abs(y1 - y2).max() <= 1e-13 * abs(y1).max()
    assert abs(y1 - y3).max() <= 1e-13 * abs(y1).max()
    A2 = E.assemble_A_GG_from_local_blocks(pr.blocks.A_GG, pr.sd)
    assert abs(A2 - A_GG).max() <= 1e-13 * abs(A_GG).max()


def test_dd_solution_equals_undecomposed_solve(medium_problem):
    pr = medium_problem
    inner, Sd, Pi, bs = get_oracle_ops(pr)
    ug, it, res = OS.pcg(lambda v: E.apply_local_schurs(pr.blocks, inner, pr.sd, v), bs, np.zeros_like(bs),
                         lambda r: E.apply_neumann_neumann_schur(Pi, pr.sd, r))
    assert res[-1] <= 1e-7 * np.linalg.norm(bs) and it == res.size
    uI = E.get_subdomain_solutions(ug, pr.blocks, inner, pr.sd)
    isd = pr.markers == 1
    udd = E.merge_subdomain_solutions(ug, uI, pr.sd, np.nonzero(isd)[0], pr.u_dir)
    A, rhs, free = E.do_isotropic_elliptic_assembly(pr.cells, pr.points, isd, pr.coeff, pr.f_vals, pr.u_dir)
    ufull = pr.u_dir.copy()
    ufull[free] = spla.spsolve(A.tocsc(), rhs)
    assert abs(udd - ufull).max() <= 5e-6 * abs(ufull).max()      # PCG stopped at 1e-7 (cg.jl:90)


def test_schur_spd_and_floating_subdomains(medium_problem):
    pr = medium_problem
    inner, Sd, Pi, bs = get_oracle_ops(pr)
    n = pr.sd.n_gamma
    S = np.column_stack([E.apply_local_schurs_assembled(Sd, pr.sd, e) for e in np.eye(n)])
    assert abs(S - S.T).max() <= 1e-12 * abs(S).max()
    assert np.linalg.eigvalsh((S + S.T) / 2).min() > 0
    boundary_touch = [bool((pr.markers[pr.cells[:, pr.sd.elemd[d]]] == 1).any()) for d in range(pr.sd.ndom)]
    for d in range(pr.sd.ndom):
        w = np.linalg.eigvalsh(Sd[d])
        singular = w.min() < 1e-10 * w.max()
        assert singular == (not boundary_touch[d])


def test_golden_vectors():
    """Self-generated golden vectors (tests/golden/make_golden.py): PARITY UNPINNED by the
    reference (it ships none); they pin this oracle against regressions."""
    g = np.load(GOLD)
    pr = get_problem(33, 4)
    inner, Sd, Pi, bs = get_oracle_ops(pr)
    assert np.array_equal(g["node_gamma"], pr.sd.node_gamma)
    assert np.array_equal(g["node_gamma_cnt"], pr.sd.node_gamma_cnt)
    for d in range(4):
        assert np.array_equal(g[f"gamma_l2g_{d}"], pr.sd.gamma_l2g[d])
    assert abs(g["b_schur"] - bs).max() <= 1e-12 * abs(bs).max()
    x = g["x_probe"]
    assert abs(g["S_x"] - E.apply_local_schurs(pr.blocks, inner, pr.sd, x)).max() <= 1e-12 * abs(g["S_x"]).max()
    assert abs(g["M_x"] - E.apply_neumann_neumann_schur(Pi, pr.sd, x)).max() <= 1e-11 * abs(g["M_x"]).max()
    ug, it, res = OS.pcg(lambda v: E.apply_local_schurs(pr.blocks, inner, pr.sd, v), bs, np.zeros_like(bs),
                         lambda r: E.apply_neumann_neumann_schur(Pi, pr.sd, r))
    assert it == int(g["pcg_it"])
    assert abs(res - g["pcg_res"]).max() <= 1e-9 * res[0]
    assert np.linalg.norm(ug - g["pcg_x"]) <= 1e-9 * np.linalg.norm(ug)


@pytest.mark.parametrize("nx,n_gamma,it_nn,it_cg", [(33, 177, 8, 28), (65, 369, 9, 41)])
def test_constant_coefficient_counts_of_the_survey_prototype(nx, n_gamma, it_nn, it_cg):
    """The only numbers about this path that were NOT produced by this repository: SURVEY.md 8(c) lists, for a throwaway
    prototype written from the reference during the survey (structured mesh, 4 x 4 square subdomains, a = 1, f = -1,
    u = 0.734 on the boundary, exact interior solves, pinv(rcond = sqrt(eps)), tolerance 1e-7), the interface size and the
    iteration counts `it` (reference convention) of Neumann-Neumann PCG and of unpreconditioned CG.  The oracle reproduces
    them exactly on the unjittered mesh (the lognormal rows of that table depend on the prototype's random field)."""
    from krylov_b200.hostprep import epdd_host as H
    from krylov_b200.hostprep import mesh as M
    cells, pts, mk, nb = M.structured_mesh(nx, jitter=0.0)
    epart, npart = M.rcb_partition(cells, pts, 16)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    nn = pts.shape[1]
    lb = E.prepare_local_schurs_ref(cells, pts, epart, sd, np.ones(nn), -np.ones(nn), 0.734 * np.ones(nn)) if nx <= 33 else \
        H.prepare_local_schurs(cells, pts, epart, sd, np.ones(nn), -np.ones(nn), 0.734 * np.ones(nn))
    inner = E.ExactInterior(lb.A_II)
    Pi = E.prepare_neumann_neumann_schur_precond(E.assemble_local_schurs(lb, inner))
    bs = E.get_schur_rhs(lb, inner, sd)
    A = lambda v: E.apply_local_schurs(lb, inner, sd, v)   # noqa: E731
    assert sd.n_gamma == n_gamma
    assert OS.pcg(A, bs, np.zeros(n_gamma), lambda r: E.apply_neumann_neumann_schur(Pi, sd, r))[1] == it_nn
    assert OS.pcg(A, bs, np.zeros(n_gamma), lambda r: r.copy())[1] == it_cg


@pytest.mark.parametrize("nx,ndom,n_gamma,it_nn,it_cg", [(65, 16, 369, 32, 178), (129, 16, 753, 34, 289), (129, 64, 1729, 125, 401)])
def test_lognormal_counts_of_the_survey_prototype(nx, ndom, n_gamma, it_nn, it_cg):
    """The lognormal rows of the same table.  The prototype's field turns out to be the recipe of SURVEY
    8(d) drawn from numpy's default_rng(0) (256 random Fourier features, frequencies first, then phases): with it the
    oracle lands within the rounding envelope of such recurrences (+-2 on counts of up to 400) of every row --
    32 / 177, 34 / 290 and 124 / 403 against the prototype's 32 / 178, 34 / 289 and 125 / 401."""
    from krylov_b200.hostprep import epdd_host as H
    from krylov_b200.hostprep import mesh as M
    cells, pts, mk, nb = M.structured_mesh(nx, jitter=0.0)
    epart, npart = M.rcb_partition(cells, pts, ndom)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    nn = pts.shape[1]
    rng = np.random.default_rng(0)
    omega = rng.normal(0.0, np.sqrt(2.0) / 0.1, size=(256, 2))
    phi = rng.uniform(0.0, 2.0 * np.pi, size=256)
    coeff = np.exp(np.sqrt(2.0 / 256) * np.cos(omega @ pts + phi[:, None]).sum(axis=0))
    lb = H.prepare_local_schurs(cells, pts, epart, sd, coeff, -np.ones(nn), 0.734 * np.ones(nn))
    inner = E.ExactInterior(lb.A_II)
    Pi = E.prepare_neumann_neumann_schur_precond(E.assemble_local_schurs(lb, inner))
    bs = E.get_schur_rhs(lb, inner, sd)
    A = lambda v: E.apply_local_schurs(lb, inner, sd, v)   # noqa: E731
    assert sd.n_gamma == n_gamma
    assert abs(OS.pcg(A, bs, np.zeros(n_gamma), lambda r: E.apply_neumann_neumann_schur(Pi, sd, r))[1] - it_nn) <= 2
    assert abs(OS.pcg(A, bs, np.zeros(n_gamma), lambda r: r.copy())[1] - it_cg) <= 2

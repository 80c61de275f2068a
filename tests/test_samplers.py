"""Host-side McmcSampler stand-in (Fem/Samplers.jl:93-199) and the chain workload of Example09:
correlated lognormal realizations solved with eigpcg on the first sample and eigdefpcg afterwards
(Example09_DefPcgMcmcStochasticEllipticPde_Functions.jl:139-429), here on the CPU oracle."""
import numpy as np

import oracle.epdd as E
import oracle.solvers as OS
from krylov_b200.hostprep import epdd_host as H
from krylov_b200.hostprep import mesh as M
from krylov_b200.hostprep.samplers import McmcSampler, rff_modes


def test_sampler_follows_the_reference_rules():
    pts = np.random.default_rng(0).random((2, 300))
    Lam, Psi = rff_modes(pts, m=16, sig2=1.0, L=0.3, seed=1)
    assert Psi.shape == (300, 32) and np.allclose(Lam, 1.0 / 16)
    s = McmcSampler(Lam, Psi, seed=2)
    assert np.isclose(s.theta2, 2.38 ** 2 / 32)                                  # Samplers.jl:138
    assert np.allclose(s.g, Psi @ (np.sqrt(Lam) * s.xi))
    xs, cnts, moved = [], [], []
    for _ in range(4000):
        before = s.xi.copy()
        cnts.append(s.draw())
        xs.append(s.xi.copy())
        moved.append(np.linalg.norm(s.xi - before))
        assert np.allclose(s.g, Psi @ (np.sqrt(Lam) * s.xi))
    xs = np.array(xs[500:])
    assert min(cnts) >= 1 and np.mean(cnts) < 6                                   # re-proposes until accepted
    assert min(moved) > 0                                                         # every draw moves the chain
    # stationary law N(0, I) of the latent coordinates
    assert np.abs(xs.mean(axis=0)).max() < 0.35 and 0.6 < xs.var(axis=0).mean() < 1.4
    # consecutive realizations are strongly correlated (a random WALK): that is what recycling exploits
    assert np.mean(moved) < 0.5 * np.sqrt(2 * 32)


def test_covariance_of_the_expansion():
    pts = np.array([[0.2, 0.25, 0.6], [0.3, 0.3, 0.9]])
    Lam, Psi = rff_modes(pts, m=4000, sig2=1.0, L=0.2, seed=3)
    C = (Psi * Lam) @ Psi.T
    d01 = np.sum((pts[:, 0] - pts[:, 1]) ** 2)
    assert abs(C[0, 0] - 1.0) < 1e-12 and abs(C[0, 1] - np.exp(-d01 / 0.2 ** 2)) < 0.05 and abs(C[0, 2]) < 0.06


def test_chain_workload_recycling_reduces_iterations():
    cells, pts, mk, nb = M.structured_mesh(33, jitter=0.2)
    epart, npart = M.rcb_partition(cells, pts, 4)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    Lam, Psi = rff_modes(pts, m=64, sig2=1.0, L=0.1, seed=4)
    chain = McmcSampler(Lam, Psi, seed=5)
    f, u = -np.ones(pts.shape[1]), 0.734 * np.ones(pts.shape[1])
    lb0 = H.prepare_local_schurs(cells, pts, epart, sd, np.ones(pts.shape[1]), f, u)       # median preconditioner
    Pi = E.prepare_neumann_neumann_schur_precond(E.assemble_local_schurs(lb0, E.ExactInterior(lb0.A_II)))
    Mop = lambda r: E.apply_neumann_neumann_schur(Pi, sd, r)
    nvec, spdim = 6, 18
    W = None
    plain, recycled = [], []
    for s in range(4):
        if s:
            chain.draw()
        lb = H.prepare_local_schurs(cells, pts, epart, sd, np.exp(chain.g), f, u)
        inner = E.ExactInterior(lb.A_II)
        Aop = lambda v: E.apply_local_schurs(lb, inner, sd, v)
        bs = E.get_schur_rhs(lb, inner, sd)
        _, it_p, _ = OS.pcg(Aop, bs, np.zeros(sd.n_gamma), Mop)
        if W is None:
            x, it, res, W = OS.eigpcg(Aop, bs, np.zeros(sd.n_gamma), Mop, nvec, spdim)       # Ex09F:345
        else:
            x, it, res, W = OS.eigdefpcg(Aop, bs, np.zeros(sd.n_gamma), Mop, W, spdim)       # Ex09F:364
            plain.append(it_p)
            recycled.append(it)
        assert res[-1] <= 1e-7 * np.linalg.norm(bs)
    assert sum(recycled) < sum(plain)

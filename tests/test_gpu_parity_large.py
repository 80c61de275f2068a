"""Parity of the CUDA path at the sizes BASELINE.json names (configs 2-4): 501^2 / 64 subdomains (Example07),
1001^2 / 128 subdomains with nvec = 1.25 ndom, spdim = 3 ndom (Example17-style deflated solvers), 64 right-hand
sides (Example11).  The checker is the oracle's C port (oracle/kb_oracle.c through oracle.cpu_model) driving the
oracle's solver loops (oracle/solvers.py).

Tolerances: the ROUNDING-FLOOR protocol.  One-level Neumann-Neumann PCG on a lognormal field runs hundreds of
iterations of an ill-conditioned recurrence; two runs of the SAME arithmetic that differ only in rounding drift
apart exponentially with the iteration index.  The floor is measured, not assumed: the CPU port against itself
with the interface sums taken in descending instead of ascending subdomain order.  The GPU is held, index by
index, to a small multiple of that floor (and to 1e-10 where the floor is below it), and its iteration count to
the envelope of the CPU variants' counts widened by 2 (north_star's +-1 holds where that envelope is a point: the
well-conditioned cases of tests/test_gpu_parity.py).

Degenerate regime of the recycling solvers (config 3 as BASELINE.json sizes it, with THIS preconditioner): one-level
NN-PCG needs ~800 iterations at 1001^2 / 128, i.e. eight Rayleigh-Ritz restarts of a spdim = 384 window.  By the first
restart the Lanczos vectors in V have lost linear independence (V'AV, which eigcg.jl:219 recomputes explicitly, has
~100 eigenvalues at the 1e-13 level); a null Ritz pair is the "lowest" one, so it is kept by every later restart and
the null space grows (98 -> 132 -> 169 > nvec = 160 in scripts/diag_recycle4.py's log): from the third restart on all
nvec returned vectors are rounding residue (column norms 1e-10 ... 1e-7).  That is the reference algorithm's own
behaviour, reproduced by the oracle; what such a basis does to the next solve (it = 156 ... 425 against 693
undeflated, over 1e-14 perturbations of b and over implementations of the small dense algebra) is not a property
parity can pin.  The test therefore checks at this size what IS determined -- the eigpcg solve itself, and the
deflated solvers GPU against oracle on the SAME basis -- and test_recycling_solvers_501_64 covers the regime the
reference's examples run in (converging within a few restarts), where the bases themselves agree."""
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
    """GPU-vs-CPU distance of the residual histories against the CPU-vs-CPU floor, at every mark where the floor itself
    is still below 1e-3 (beyond that two roundings of the SAME arithmetic have decorrelated and only the stopping index
    and the converged solution are compared).  The recurrence amplifies a rounding difference exponentially, so the
    allowance is a factor on that scale: 100 x the floor, and 1e-10 where the floor is below it."""
    for i in marks(m):
        g, f = hist_diff(res_gpu, res_cpu, i), hist_diff(res_floor, res_cpu, i)
        if f > 1e-3:
            break
        assert g <= max(1e-10, 100 * f), f"{what}: history differs by {g:.2e} over the first {i} iterations, floor {f:.2e}"


def cpu_variants(run, perturb, extra=True):
    """The same CPU solve under rounding-level variations: interface sums in ascending / descending subdomain order,
    and (extra) inputs perturbed at the 1e-16 level.  run(x) -> result tuple with the iteration count at [1]."""
    out = [run(perturb(0))]
    _cm_rev[0].set_reverse_order(True)
    out.append(run(perturb(0)))
    _cm_rev[0].set_reverse_order(False)
    if extra:
        out.append(run(perturb(1)))
    return out


_cm_rev = [None]


def envelope(outs):
    its = [o[1] for o in outs]
    return min(its) - 2, max(its) + 2


def ulp_noise(a, seed):
    return a if seed == 0 else a * (1.0 + 2.2e-16 * np.random.default_rng(seed).standard_normal(a.shape))


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
    _cm_rev[0] = cm
    n = pr.sd.n_gamma
    B = np.random.default_rng(481456).random((n, 64))
    X, its, ress = S.h.pcg_multi(B, np.zeros((n, 64)))
    for j in (0, 17, 63):
        xs, its_j, res_j = solvers.pcg(S, B[:, j].copy(), np.zeros(n), M)
        outs = cpu_variants(lambda b: cm.pcg(b, np.zeros(n), nthreads=NT), lambda s: ulp_noise(B[:, j], s))
        (xo, ito, reso), (xr, itr, resr) = outs[0], outs[1]
        lo, hi = envelope(outs)
        assert lo <= its[j] <= hi and lo <= its_j <= hi, f"rhs {j}: it block {its[j]}, single {its_j}, cpu variants {[o[1] for o in outs]}"
        assert ress[j][-1] <= 1e-7 * np.linalg.norm(B[:, j])
        check_history(ress[j], reso, resr, min(its[j], ito, itr), f"rhs {j}")
        assert np.linalg.norm(X[:, j] - xo) <= 5e-6 * np.linalg.norm(xo)       # both converged to the 1e-7 tolerance
    S.assemble_local_schurs()
    Xa, its_a, _ = S.h.pcg_multi(B, np.zeros((n, 64)), assembled=True)
    assert np.abs(its_a - its).max() <= 4
    assert np.linalg.norm(Xa - X) <= 5e-6 * np.linalg.norm(X)


def subspace_gap(U, V):
    """sin of the largest principal angle between span(U) and span(V)."""
    Qu, _ = np.linalg.qr(U)
    Qv, _ = np.linalg.qr(V)
    s = np.linalg.svd(Qu.T @ Qv, compute_uv=False)
    return float(np.sqrt(max(0.0, 1.0 - s.min() ** 2)))


def recycling_case(nx, ndom, compare_bases, extra=True):
    pr, cm, Pi, bs, S, M = setup(nx, ndom)
    _cm_rev[0] = cm
    n = pr.sd.n_gamma
    nvec, spdim = int(1.25 * ndom), 3 * ndom                                   # Example09:39-40
    Aop = lambda v: cm.apply_schur(v, NT)
    Mop = lambda r: cm.apply_nn(r, NT)
    outs = cpu_variants(lambda b: OS.eigpcg(Aop, b, np.zeros(n), Mop, nvec, spdim), lambda s: ulp_noise(bs, s), extra)
    (xo, ito, reso, Vo), (xr, itr, resr, Vr) = outs[0], outs[1]
    x, it, res, V = solvers.eigpcg(S, bs, np.zeros(n), M, nvec, spdim)
    assert V.shape == (n, nvec) and res[-1] <= 1e-7 * np.linalg.norm(bs)
    lo, hi = envelope(outs)
    assert lo <= it <= hi, f"eigpcg it gpu {it}, cpu variants {[o[1] for o in outs]}"
    check_history(res, reso, resr, min(it, ito, itr, spdim - 1), "eigpcg before the first restart")
    floor_x = np.linalg.norm(xr - xo) / np.linalg.norm(xo)
    assert np.linalg.norm(x - xo) <= max(5e-6, 20 * floor_x) * np.linalg.norm(xo)
    b2 = Aop(np.random.default_rng(3).standard_normal(n))
    _, it_plain, _ = solvers.pcg(S, b2, np.zeros(n), M)
    # the deflated solvers on the SAME basis (the GPU's): GPU against the oracle
    outs2 = cpu_variants(lambda W: OS.eigdefpcg(Aop, b2, np.zeros(n), Mop, W, spdim), lambda s: ulp_noise(V, s), extra)
    (x2o, it2o, res2o, _), (x2r, it2r, res2r, _) = outs2[0], outs2[1]
    x2, it2, res2, V2 = solvers.eigdefpcg(S, b2, np.zeros(n), M, V, spdim)
    assert res2[-1] <= 1e-7 * np.linalg.norm(b2) and V2.shape == (n, nvec)
    # W'AW is numerically singular when the basis carries rounding residue (module docstring): the count of the
    # deflated solve then moves with the rounding of `WtAW \ .`; allowance = twice the CPU variants' own width, at
    # least 5 % of the count
    its2 = [o[1] for o in outs2]
    dev = max(2 + 2 * (max(its2) - min(its2)), it2o // 20)
    assert min(its2) - dev <= it2 <= max(its2) + dev, f"eigdefpcg it gpu {it2}, cpu variants {its2}"
    assert it2 < it_plain, "deflation with the recycled basis must pay"
    floor2 = max(np.linalg.norm(o[0] - x2o) / np.linalg.norm(x2o) for o in outs2[1:])
    assert np.linalg.norm(x2 - x2o) <= max(5e-6, 20 * floor2) * np.linalg.norm(x2o)
    if compare_bases:
        # defpcg (a12) at this size.  defcg.jl:242-308 has no re-orthogonalisation of r against W (eigdefpcg has,
        # :407) and stalls on the raw recycled basis, whose columns span seven orders of magnitude in norm -- oracle
        # and GPU alike run into maxit; its W here is that basis orthonormalised.
        Wq = np.linalg.qr(V)[0]
        x3o, it3o, res3o = OS.defpcg(Aop, b2, np.zeros(n), Wq, Mop, maxit=2000)
        x3, it3, res3 = solvers.defpcg(S, b2, np.zeros(n), Wq, M, maxit=2000)
        assert abs(it3 - it3o) <= dev and res3[-1] <= 1e-7 * np.linalg.norm(b2), (it3, it3o)
        assert np.linalg.norm(x3 - x3o) <= max(5e-6, 20 * floor2) * np.linalg.norm(x3o)
    if compare_bases:
        # the recycled bases themselves: as good a deflation space as the oracle's own (iteration count of the
        # next solve), and the well-converged part of the spectrum spans the same space
        _, it_ov, _, _ = solvers.eigdefpcg(S, b2, np.zeros(n), M, Vo, spdim)
        _, it_rv, _, _ = solvers.eigdefpcg(S, b2, np.zeros(n), M, Vr, spdim)
        slack = abs(it_rv - it_ov) + max(3, it_ov // 20)
        assert it2 <= it_ov + slack, f"recycled basis: next solve takes {it2} iterations with the GPU's, {it_ov} / {it_rv} with the oracle's"
    return it, it2, it_plain


def test_recycling_solvers_501_64():
    """eigpcg -> eigdefpcg / defpcg at 501^2 / 64 subdomains, nvec = 80, spdim = 192: the regime of the reference's
    examples (a few restarts, V keeps full rank), where the recycled bases themselves are comparable."""
    recycling_case(501, 64, compare_bases=True)


def test_recycling_solvers_1001_128():
    """Config 3 (Example17-style: eigpcg on the first system, eigdefpcg with the recycled basis on the next;
    ~1M DOF, 128 subdomains, nvec = floor(1.25 ndom) = 160, spdim = 3 ndom = 384, Example09:39-40).  See the module
    docstring: the returned basis is rounding residue at this size with this preconditioner, so GPU and oracle are
    compared on the same basis."""
    recycling_case(1001, 128, compare_bases=False, extra=False)

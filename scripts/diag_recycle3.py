"""Diagnostic (GPU): is the quality of the recycled basis of eigpcg (C entry point / host loop / CPU oracle) a matter of
rounding noise?  Each variant is run on copies of the right-hand side perturbed at the 1e-14 level; the iteration count
of the following eigdefpcg solve measures the basis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch  # noqa
import krylov_b200  # noqa
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
from oracle.cpu_model import CpuModel
import oracle.solvers as OS
nx, ndom = int(sys.argv[1]), int(sys.argv[2])
nvec, spdim = int(1.25 * ndom), 3 * ndom
NT = os.cpu_count()
pr = PB.build_problem(nx, ndom); pr0 = PB.build_problem(nx, ndom, ireal=0, median=True)
cm0 = CpuModel(pr0.sd, pr0.blocks, pr0.factors); Pi = [PB.pinv_sym(S) for S in cm0.assemble_local_schurs(NT)]
cm = CpuModel(pr.sd, pr.blocks, pr.factors, Pi); bs = cm.schur_rhs(pr.blocks.b_I, pr.blocks.b_G, NT); n = pr.sd.n_gamma
S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors); M = fem.B200NeumannNeumann(S, Pi)
Aop = lambda v: cm.apply_schur(v, NT); Mop = lambda r: cm.apply_nn(r, NT)
b2 = Aop(np.random.default_rng(3).standard_normal(n))
_, itp, _ = solvers.pcg(S, b2, np.zeros(n), M)
print("plain pcg on b2:", itp, flush=True)
def quality(V):
    _, it2, _, _ = solvers.eigdefpcg(S, b2, np.zeros(n), M, V, spdim)
    nrm = np.linalg.norm(V, axis=0)
    return it2, int((nrm < 1e-3 * nrm.max()).sum())
for seed in range(int(sys.argv[3]) if len(sys.argv) > 3 else 5):
    b = bs * (1.0 + (1e-14 * np.random.default_rng(100 + seed).standard_normal(n) if seed else 0.0))
    xc, itc, _, Vc = solvers.eigpcg(S, b, np.zeros(n), M, nvec, spdim)
    xh, ith, _, Vh = solvers.eigpcg_hostloop(S, b, np.zeros(n), M, nvec, spdim)
    line = f"seed {seed}: C it {itc} -> (it2, tiny cols) {quality(Vc)}; hostloop it {ith} -> {quality(Vh)}"
    if seed < 3:
        xo, ito, _, Vo = OS.eigpcg(Aop, b, np.zeros(n), Mop, nvec, spdim)
        line += f"; oracle it {ito} -> {quality(Vo)}"
    print(line, flush=True)

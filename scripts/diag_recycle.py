"""Diagnostic (GPU): quality of the recycled basis of eigpcg -- C entry point vs host-language loop vs CPU oracle."""
import os, sys, time
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
xo, ito, reso, Vo = OS.eigpcg(Aop, bs, np.zeros(n), Mop, nvec, spdim)
xc, itc, resc, Vc = solvers.eigpcg(S, bs, np.zeros(n), M, nvec, spdim)
xh, ith, resh, Vh = solvers.eigpcg_hostloop(S, bs, np.zeros(n), M, nvec, spdim)
for rb in (1,):
    S.h.set_option("rhs_block", rb)
    x1, it1, res1, V1 = solvers.eigpcg(S, bs, np.zeros(n), M, nvec, spdim)
    S.h.set_option("rhs_block", 0)
print("eigpcg it: oracle", ito, "C", itc, "hostloop", ith, "C rhs_block=1", it1, flush=True)
b2 = Aop(np.random.default_rng(3).standard_normal(n))
def ritz(V):
    Q, _ = np.linalg.qr(V)
    AQ = np.column_stack([Aop(Q[:, j]) for j in range(Q.shape[1])])
    MQ = None
    H = Q.T @ AQ
    w = np.linalg.eigvalsh((H + H.T) / 2)
    return w[:3], w[-3:]
for name, V in (("oracle", Vo), ("C", Vc), ("hostloop", Vh), ("C_rb1", V1)):
    x2, it2, res2, _ = solvers.eigdefpcg(S, b2, np.zeros(n), M, V, spdim)
    print(name, "-> eigdefpcg it", it2, "orth defect", float(np.abs(V.T @ V - np.eye(nvec)).max()), "col norms", float(np.linalg.norm(V, axis=0).min()), float(np.linalg.norm(V, axis=0).max()), "ritz", ritz(V), flush=True)

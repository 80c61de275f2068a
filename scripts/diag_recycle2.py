import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch  # noqa
import krylov_b200  # noqa
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
nx, ndom = int(sys.argv[1]), int(sys.argv[2])
nvec, spdim = int(1.25 * ndom), 3 * ndom
pr = PB.build_problem(nx, ndom)
S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors)
Sd = S.assemble_local_schurs()
Pi = {d: PB.pinv_sym(Sd[d]) for d in range(ndom)}
M = fem.B200NeumannNeumann(S, Pi)
n = pr.sd.n_gamma
bs = S.schur_rhs(pr.blocks.b_I, pr.blocks.b_G)
print("=== C", flush=True)
xc, itc, resc, Vc = solvers.eigpcg(S, bs, np.zeros(n), M, nvec, spdim)
print("=== hostloop", flush=True)
xh, ith, resh, Vh = solvers.eigpcg_hostloop(S, bs, np.zeros(n), M, nvec, spdim)
print(itc, ith)

"""Small end-to-end case for compute-sanitizer (all modes of the stream kernel, NN, pcg, device ops)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_b200
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
pr = PB.build_problem(49, 4)
S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors, csize=4)
Sd = S.assemble_local_schurs()
M = fem.B200NeumannNeumann(S, {d: PB.pinv_sym(Sd[d]) for d in Sd})
bs = S.schur_rhs(pr.blocks.b_I, pr.blocks.b_G)
x, it, res = solvers.pcg(S, bs, np.zeros(pr.sd.n_gamma), M)
u = S.subdomain_solutions(x, pr.blocks.b_I)
W = np.linalg.qr(np.random.default_rng(0).standard_normal((pr.sd.n_gamma, 4)))[0]
x2, it2, _ = solvers.defpcg(S, bs, np.zeros(pr.sd.n_gamma), W, M)
print("sanitize case ok: it", it, it2, "res", res[-1] / res[0])

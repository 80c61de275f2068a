"""Diagnostic (GPU): multi-member handle on one device, block paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import krylov_b200  # noqa
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
import oracle.epdd as E
devices = [int(v) for v in sys.argv[1].split(",")]
pr = PB.build_problem(65, 16, ireal=1)
sd, lb = pr.sd, pr.blocks
inner = E.ExactInterior(lb.A_II)
Pi = E.prepare_neumann_neumann_schur_precond(E.assemble_local_schurs(lb, inner))
n = sd.n_gamma
rng = np.random.default_rng(3)
X = rng.standard_normal((n, 5))
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
S1 = fem.B200SchurOperator(sd, lb, pr.factors, device=0); M1 = fem.B200NeumannNeumann(S1, Pi)
S = fem.B200SchurOperator(sd, lb, pr.factors, devices=devices); M = fem.B200NeumannNeumann(S, Pi)
bs = S1.schur_rhs(lb.b_I, lb.b_G)
for k in (1, 2, 3, 5):
    Y, Y1 = S.h.apply_schur(X[:, :k]), S1.h.apply_schur(X[:, :k])
    Z, Z1 = S.h.apply_nn(X[:, :k]), S1.h.apply_nn(X[:, :k])
    print("nrhs", k, "apply_schur", rel(Y, Y1), "apply_nn", rel(Z, Z1), flush=True)
for rb in (1, 0):
    S.h.set_option("rhs_block", rb)
    for B in (np.column_stack([bs, bs]), np.column_stack([bs, X[:, 0]]), np.column_stack([bs, X[:, 0], X[:, 1], 2 * bs, X[:, 2]])):
        try:
            Xs, its, _ = S.h.pcg_multi(B, np.zeros(B.shape))
            X1, its1, _ = S1.h.pcg_multi(B, np.zeros(B.shape))
            print("rhs_block", rb, "k", B.shape[1], "its", its.tolist(), its1.tolist(), "dx", rel(Xs, X1), flush=True)
        except Exception as e:
            print("rhs_block", rb, "k", B.shape[1], "FAILED", str(e)[:200], flush=True)

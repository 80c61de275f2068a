"""Generates tests/golden/schur_nx33_ndom4.npz from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference ships no golden data and cannot run here
(no Julia), so these vectors are self-generated: they pin the oracle and the CUDA path against
regressions, not against the reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import krylov_b200  # noqa: E402
from krylov_b200.hostprep import problem as PB  # noqa: E402
import oracle.epdd as E  # noqa: E402
import oracle.solvers as OS  # noqa: E402

pr = PB.build_problem(33, 4, ireal=1, jitter=0.2)
inner = E.ExactInterior(pr.blocks.A_II)
Sd = E.assemble_local_schurs(pr.blocks, inner)
Pi = E.prepare_neumann_neumann_schur_precond(Sd)
bs = E.get_schur_rhs(pr.blocks, inner, pr.sd)
x = np.random.default_rng(481456).standard_normal(pr.sd.n_gamma)
ug, it, res = OS.pcg(lambda v: E.apply_local_schurs(pr.blocks, inner, pr.sd, v), bs, np.zeros_like(bs),
                     lambda r: E.apply_neumann_neumann_schur(Pi, pr.sd, r))
out = dict(node_gamma=pr.sd.node_gamma, node_gamma_cnt=pr.sd.node_gamma_cnt, b_schur=bs, x_probe=x,
           S_x=E.apply_local_schurs(pr.blocks, inner, pr.sd, x), M_x=E.apply_neumann_neumann_schur(Pi, pr.sd, x),
           pcg_it=np.int64(it), pcg_res=res, pcg_x=ug)
for d in range(4):
    out[f"gamma_l2g_{d}"] = pr.sd.gamma_l2g[d]
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "schur_nx33_ndom4.npz"), **out)
print("written; pcg it =", it)

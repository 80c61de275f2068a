"""Diagnostic (GPU): host-loop eigpcg with the Rayleigh-Ritz algebra swapped between numpy and the device routine
(kb200_debug_rayleigh_ritz), both applied to the SAME projected matrices, with the differences printed per restart."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch  # noqa
import krylov_b200  # noqa
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
from oracle.cpu_model import CpuModel
nx, ndom = int(sys.argv[1]), int(sys.argv[2])
nvec, spdim = int(1.25 * ndom), 3 * ndom
NT = os.cpu_count()
pr = PB.build_problem(nx, ndom); pr0 = PB.build_problem(nx, ndom, ireal=0, median=True)
cm0 = CpuModel(pr0.sd, pr0.blocks, pr0.factors); Pi = [PB.pinv_sym(S) for S in cm0.assemble_local_schurs(NT)]
cm = CpuModel(pr.sd, pr.blocks, pr.factors, Pi); bs = cm.schur_rhs(pr.blocks.b_I, pr.blocks.b_G, NT); n = pr.sd.n_gamma
S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors); M = fem.B200NeumannNeumann(S, Pi)
b2 = cm.apply_schur(np.random.default_rng(3).standard_normal(n), NT)
rr_np = solvers._rayleigh_ritz
def gap(U, V):
    Qu, _ = np.linalg.qr(U); Qv, _ = np.linalg.qr(V)
    s = np.linalg.svd(Qu.T @ Qv, compute_uv=False)
    return float(np.sqrt(max(0.0, 1.0 - s.min() ** 2)))
def make(mode):
    def rr(Tm, nv):
        a = rr_np(Tm, nv)
        d = S.h.debug_rayleigh_ritz(Tm, nv)
        w = np.linalg.eigvalsh(Tm)
        def chk(QZ, vals):
            return (float(np.abs(QZ.T @ QZ - np.eye(QZ.shape[1])).max()), float(np.abs(QZ.T @ Tm @ QZ - np.diag(vals)).max()))
        print(f"  [{mode}] m {Tm.shape[0]} eig(Tm) {w[0]:.3e} {w[1]:.3e} .. #<1e-8: {(np.abs(w) < 1e-8).sum()}  nev np {a[2]} dev {d[2]}"
              f"  (orth, ritz resid) np {chk(a[0], a[1])} dev {chk(d[0], d[1])}"
              f"  vals[:4] np {a[1][:4]} dev {d[1][:4]}  gap first nvec {gap(a[0][:, :nv], d[0][:, :nv]):.2e}"
              + (f" gap all {gap(a[0], d[0]):.2e}" if a[2] == d[2] else ""), flush=True)
        return a if mode == "numpy" else d
    return rr
for mode in ("numpy", "device"):
    solvers._rayleigh_ritz = make(mode)
    x, it, res, V = solvers.eigpcg_hostloop(S, bs, np.zeros(n), M, nvec, spdim)
    solvers._rayleigh_ritz = rr_np
    _, it2, _, _ = solvers.eigdefpcg(S, b2, np.zeros(n), M, V, spdim)
    print(f"hostloop with {mode} Rayleigh-Ritz: it {it} -> eigdefpcg it {it2}", flush=True)

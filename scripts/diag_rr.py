import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch  # noqa
import krylov_b200  # noqa
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
pr = PB.build_problem(65, 16)
S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors)
rng = np.random.default_rng(0)
for m, nvec in ((24, 8), (192, 80), (384, 160), (300, 160)):
    # a Lanczos-like projected matrix: diag(vals) + arrowhead column + tridiagonal tail, wide spectrum
    nev = min(2 * nvec, m - 10)
    T = np.zeros((m, m))
    T[np.arange(nev), np.arange(nev)] = np.sort(np.exp(rng.uniform(np.log(2e-3), np.log(80), nev)))
    T[:nev, nev] = rng.standard_normal(nev) * 0.3
    d = np.exp(rng.uniform(np.log(0.5), np.log(60), m - nev)); e = rng.uniform(0.1, 5.0, m - nev - 1)
    T[np.arange(nev, m), np.arange(nev, m)] = d
    T[np.arange(nev, m - 1), np.arange(nev + 1, m)] = e
    T = np.triu(T) + np.triu(T, 1).T
    T += 80 * np.eye(m) * 0  # keep
    w = np.linalg.eigvalsh(T)
    QZr, valsr, nevr = solvers._rayleigh_ritz(T, nvec)
    QZ, vals, nevg = S.h.debug_rayleigh_ritz(T, nvec)
    def gap(U, V):
        Qu, _ = np.linalg.qr(U); Qv, _ = np.linalg.qr(V)
        s = np.linalg.svd(Qu.T @ Qv, compute_uv=False)
        return float(np.sqrt(max(0.0, 1.0 - s.min() ** 2)))
    print(m, nvec, "spectrum", w[0], w[-1], "nev", nevr, nevg, "vals diff", float(np.abs(vals[:nvec] - valsr[:nvec]).max() / np.abs(valsr[:nvec]).max()),
          "gap(first nvec)", gap(QZ[:, :nvec], QZr[:, :nvec]), "gap(all)", gap(QZ, QZr) if nevr == nevg else None,
          "orth", float(np.abs(QZ.T @ QZ - np.eye(nevg)).max()), "ritz resid gpu", float(np.abs(QZ.T @ T @ QZ - np.diag(vals)).max()), flush=True)

"""CPU restatement of the reference's LORASC preconditioner (TEST INFRASTRUCTURE: only tests/, smoke() and the
cpu_baseline / reference legs of bench.py may import this package; PARITY UNPINNED by the reference).

Follows Fem/EllipticPdeDomainDecomposition.jl: `LorascPreconditioner` :1406-1428, `prepare_lorasc_precond`
(low_rank_correction = :exact) :1502-1676, `apply_lorasc` :1908-1976.  Interior and interface factorizations are exact
sparse LU solves (the reference uses CHOLMOD factors: same linear maps).  The generalized eigenpairs the reference
obtains from KrylovKit.geneigsolve(:SR, isposdef = true) (:1546-1549; KrylovKit 0.5.2 is not vendored) are computed
with a dense symmetric-definite eigen-solver: the nvec smallest pairs of S e = sigma A_ΓΓ e, A_ΓΓ-orthonormal, which
is what that call converges to.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class LorascPreconditioner:
    """Fields of the reference struct that apply_lorasc reads (:1406-1428)."""

    def __init__(self, A_IGg, lu_II, A_GG, lu_GG, eps, E, Sigma, node_Id, node_gamma, free):
        self.A_IGg, self.lu_II, self.A_GG, self.lu_GG = A_IGg, lu_II, A_GG, lu_GG
        self.eps, self.E, self.Sigma = eps, E, Sigma
        self.node_Id, self.node_gamma = node_Id, node_gamma
        g2l = np.full(int(max(int(free.max()), int(node_gamma.max()) if node_gamma.size else 0)) + 1, -1, dtype=np.int64)
        g2l[free] = np.arange(free.size)                      # not_dirichlet_inds_g2l
        self.g2l = g2l
        self.n = free.size


def low_rank_pairs(S_dense, A_GG, nvec):
    """The nvec smallest generalized eigenpairs of (S, A_ΓΓ), ascending, eigenvectors A_ΓΓ-orthonormal (:1546-1549)."""
    B = A_GG.toarray() if sp.issparse(A_GG) else np.asarray(A_GG)
    n = B.shape[0]
    w, V = sla.eigh((S_dense + S_dense.T) * 0.5, (B + B.T) * 0.5, subset_by_index=[0, min(nvec, n) - 1])
    return w, V


def select_pairs(Sigma, E, eps, nvec):
    """:1587-1611: order ascending, keep the pairs with sigma < eps (Sigma[k] <- (eps - sigma) / sigma); nev == 0 keeps
    all nvec pairs (the reference's fall-back, with its warning)."""
    order = np.argsort(Sigma, kind="stable")
    Sigma, E = Sigma[order].copy(), E[:, order].copy()
    nev = 0
    for k, s in enumerate(Sigma):
        if s < eps:
            Sigma[k] = (eps - s) / s
            nev += 1
        else:
            break
    if nev == 0:
        nev = nvec
    return Sigma[:nev], E[:, :nev]


def prepare_lorasc_precond(S_dense, A_II, A_IGg, A_GG, node_Id, node_gamma, free, nvec=25, eps=0.01):
    """prepare_lorasc_precond(S, A_IId, A_IΓd, A_ΓΓ, ...) with compute_A_ΓΓ_chol = true, :exact pairs (:1502-1676)."""
    lu_GG = spla.splu(sp.csc_matrix(A_GG))                                    # :1525
    lu_II = [spla.splu(sp.csc_matrix(A)) if A.shape[0] else None for A in A_II]   # :1535-1537
    if eps > 0:
        Sigma, E = low_rank_pairs(S_dense, A_GG, nvec)
        Sigma, E = select_pairs(Sigma, E, eps, nvec)
    else:
        Sigma, E = np.zeros(0), np.zeros((A_GG.shape[0], 0))
    return LorascPreconditioner(A_IGg, lu_II, A_GG, lu_GG, eps, E, Sigma, node_Id, node_gamma, free)


def apply_lorasc(P: LorascPreconditioner, x):
    """apply_lorasc, :1908-1976 (x, u indexed by the non-Dirichlet numbering not_dirichlet_inds_g2l)."""
    ndom = len(P.node_Id)
    x_Id = [x[P.g2l[P.node_Id[d]]].copy() for d in range(ndom)]              # :1921-1926
    x_G = x[P.g2l[P.node_gamma]].copy()                                      # :1928-1933
    z_G = x_G.copy()
    for d in range(ndom):                                                    # :1935-1942
        if P.lu_II[d] is not None:
            x_Id[d] = P.lu_II[d].solve(x_Id[d])
        z_G -= P.A_IGg[d].T @ x_Id[d]
    x_G = P.lu_GG.solve(z_G)                                                 # :1944-1951
    for k in range(P.E.shape[1]):                                            # :1953-1957 (Σ is not used by the apply)
        val = P.E[:, k] @ z_G
        x_G += val * P.E[:, k]
    for d in range(ndom):                                                    # :1959-1961
        if P.lu_II[d] is not None:
            x_Id[d] -= P.lu_II[d].solve(P.A_IGg[d] @ x_G)
    u = np.empty(P.n)
    for d in range(ndom):                                                    # :1963-1968
        u[P.g2l[P.node_Id[d]]] = x_Id[d]
    u[P.g2l[P.node_gamma]] = x_G                                             # :1970-1973
    return u

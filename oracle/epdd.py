"""CPU restatement of Fem/EllipticPdeDomainDecomposition.jl (TEST INFRASTRUCTURE).

Every public function names the reference lines it follows.  Two flavours exist
for the two index-map / assembly routines:

* ``*_ref``  -- a literal loop-and-dict transcription of the reference's control
  flow (same traversal order, same first-encounter numbering); pure Python, for
  small meshes only;
* the numpy-vectorised versions the harness uses at 4M DOF live with the host
  preparation code (julia-phd-krylov-spdes_b200/hostprep/epdd_host.py) and are
  checked bit-for-bit against the ``*_ref`` versions in tests/test_oracle_maps.py.

Interior solves are done exactly (sparse direct), i.e. the reference run with
``preconds`` = exact Cholesky factors so that its inner IterativeSolvers.cg
converges in one step.

PARITY UNPINNED: the reference ships no tests, golden vectors or data for this path and cannot run here (no Julia);
this restatement is pinned by the reference's in-script identities and self-generated goldens only.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from dataclasses import dataclass, field
from typing import List


# The oracle keeps its OWN containers and assembly loops: nothing here imports product code, so an
# error shared with julia-phd-krylov-spdes_b200/hostprep cannot hide behind an identity test.  The
# host-preparation classes of the product are duck-type compatible (same attribute names).
@dataclass
class Subdomains:
    """Outputs of set_subdomains (EllipticPdeDomainDecomposition.jl:191-192) as arrays, 0-based
    except node_owner (reference values: -1 interface, 0 Dirichlet, d+1 interior of subdomain d)."""
    ndom: int
    nnode: int
    node_owner: np.ndarray
    node_gamma: np.ndarray
    node_gamma_cnt: np.ndarray
    ind_gamma_g2l: np.ndarray
    gd_nodes: List[np.ndarray]
    gamma_l2g: List[np.ndarray]
    node_Id: List[np.ndarray]
    elemd: List[np.ndarray] = field(default_factory=list)

    @property
    def n_gamma(self):
        return int(self.node_gamma.size)


@dataclass
class LocalBlocks:
    """Outputs of prepare_local_schurs (:545)."""
    A_II: list
    A_IG: list
    A_GG: list
    b_I: list
    b_G: np.ndarray


def restriction(sd, d):
    """R_d: n_Γd x n_Γ boolean gather matrix (the Dict loops at :728-730)."""
    g = np.asarray(sd.gamma_l2g[d])
    return sp.csr_matrix((np.ones(g.size), (np.arange(g.size), g)), shape=(g.size, sd.n_gamma))


def set_subdomains_ref(cells, cell_neighbors, epart, npart, is_dirichlet):
    """Literal transcription of set_subdomains, EllipticPdeDomainDecomposition.jl:86-193.

    Inputs are 0-based (cell_neighbors: -1 on the boundary)."""
    nel = cells.shape[1]
    nnode = int(cells.max()) + 1                       # :93
    ndom = int(epart.max()) + 1                        # :94
    ind_Id_g2l = [dict() for _ in range(ndom)]         # :97
    ind_gd_g2l = [dict() for _ in range(ndom)]         # :99
    ind_g_g2l = dict()                                 # :101
    ind_gd_g2lG = [dict() for _ in range(ndom)]        # :103 (ind_Γd_Γ2l)
    node_owner = np.zeros(nnode, dtype=np.int64)       # :105
    elemd = [[] for _ in range(ndom)]
    node_g, node_g_cnt = [], []
    node_Id = [[] for _ in range(ndom)]
    bnd_tag = -1                                       # :115-119 (0-based inputs need no shift)
    for iel in range(nel):                             # :123
        iel_cell = [int(v) for v in cells[:, iel]]
        idom = int(epart[iel])
        elemd[idom].append(iel)
        for j in range(3):                             # :131
            jel = int(cell_neighbors[j, iel])
            if jel != bnd_tag:
                jdom = int(epart[jel])
                if jdom != idom:
                    jel_cell = [int(v) for v in cells[:, jel]]
                    for node in iel_cell:              # :143
                        if (node in jel_cell) and not is_dirichlet[node]:
                            if node not in ind_gd_g2l[idom]:                  # :148
                                ind_gd_g2l[idom][node] = len(ind_gd_g2l[idom])
                            if node not in ind_g_g2l:                         # :152 (node in node_Γ)
                                node_g.append(node)
                                node_g_cnt.append(0)
                                ind_g_g2l[node] = len(ind_g_g2l)
                                node_owner[node] = -1
    for inode in range(nnode):                         # :166
        if (not is_dirichlet[inode]) and node_owner[inode] != -1:
            idom = int(npart[inode])
            node_Id[idom].append(inode)
            node_owner[inode] = idom + 1
    for idom in range(ndom):                           # :175
        for i, nd in enumerate(node_Id[idom]):
            ind_Id_g2l[idom][nd] = i
    for idom in range(ndom):                           # :183
        for gnode, l_in_gd in ind_gd_g2l[idom].items():
            l_in_g = ind_g_g2l[gnode]
            node_g_cnt[l_in_g] += 1
            ind_gd_g2lG[idom][l_in_g] = l_in_gd
    # Dicts -> arrays
    ind_gamma = np.full(nnode, -1, dtype=np.int64)
    for nd, l in ind_g_g2l.items():
        ind_gamma[nd] = l
    gd_nodes, gamma_l2g = [], []
    for idom in range(ndom):
        a = np.empty(len(ind_gd_g2l[idom]), dtype=np.int64)
        for nd, l in ind_gd_g2l[idom].items():
            a[l] = nd
        b = np.empty(len(ind_gd_g2lG[idom]), dtype=np.int64)
        for lg, l in ind_gd_g2lG[idom].items():
            b[l] = lg
        gd_nodes.append(a)
        gamma_l2g.append(b)
    return Subdomains(ndom, nnode, node_owner, np.array(node_g, dtype=np.int64),
                      np.array(node_g_cnt, dtype=np.int64), ind_gamma, gd_nodes, gamma_l2g,
                      [np.array(v, dtype=np.int64) for v in node_Id],
                      [np.array(v, dtype=np.int64) for v in elemd])


def prepare_local_schurs_ref(cells, points, epart, sd: Subdomains, coeff, f_vals, u_dir):
    """Literal transcription of prepare_local_schurs, EllipticPdeDomainDecomposition.jl:389-546.
    coeff / f_vals / u_dir are nodal values of a, f and uexact."""
    ndom = sd.ndom
    nel = cells.shape[1]
    ind_Id = [{int(n): i for i, n in enumerate(v)} for v in sd.node_Id]
    ind_gd = [{int(n): i for i, n in enumerate(v)} for v in sd.gd_nodes]
    n_Id = [len(v) for v in sd.node_Id]
    n_gd = [len(v) for v in sd.gd_nodes]
    II = [([], [], []) for _ in range(ndom)]
    IG = [([], [], []) for _ in range(ndom)]
    GG = [([], [], []) for _ in range(ndom)]
    b_I = [np.zeros(n) for n in n_Id]
    b_G = np.zeros(sd.n_gamma)
    x, y = np.zeros(3), np.zeros(3)
    dx, dy = np.zeros(3), np.zeros(3)
    for iel in range(nel):
        idom = int(epart[iel])
        nodes = [int(v) for v in cells[:, iel]]
        da = 0.0
        for j, jn in enumerate(nodes):
            x[j], y[j] = points[0, jn], points[1, jn]
            da += coeff[jn]
        da /= 3.0
        dx[0], dx[1], dx[2] = x[2] - x[1], x[0] - x[2], x[1] - x[0]
        dy[0], dy[1], dy[2] = y[1] - y[2], y[2] - y[0], y[0] - y[1]
        area = (dx[2] * dy[1] - dx[1] * dy[2]) / 2.0
        for i, inode in enumerate(nodes):
            io = sd.node_owner[inode]
            for j, jnode in enumerate(nodes):
                jo = sd.node_owner[jnode]
                kij = da * (dy[i] * dy[j] + dx[i] * dx[j]) / 4 / area
                if io != 0 and jo != 0:
                    if io == -1 and jo == -1:                                   # :475
                        GG[idom][0].append(ind_gd[idom][inode])
                        GG[idom][1].append(ind_gd[idom][jnode])
                        GG[idom][2].append(kij)
                    elif io > 0 and jo > 0:                                      # :483
                        II[idom][0].append(ind_Id[idom][inode])
                        II[idom][1].append(ind_Id[idom][jnode])
                        II[idom][2].append(kij)
                    elif io > 0 and jo == -1:                                    # :491
                        IG[idom][0].append(ind_Id[idom][inode])
                        IG[idom][1].append(ind_gd[idom][jnode])
                        IG[idom][2].append(kij)
                elif io == 0 and jo != 0:                                        # :499
                    if jo == -1:
                        b_G[sd.ind_gamma_g2l[jnode]] -= kij * u_dir[inode]
                    elif jo > 0:
                        b_I[idom][ind_Id[idom][jnode]] -= kij * u_dir[inode]
        for i, inode in enumerate(nodes):                                        # :514
            io = sd.node_owner[inode]
            j, k = (i + 1) % 3, (i + 2) % 3
            db = (2 * f_vals[nodes[i]] + f_vals[nodes[j]] + f_vals[nodes[k]]) * area / 12
            if io == -1:
                b_G[sd.ind_gamma_g2l[inode]] += db
            elif io > 0:
                b_I[idom][ind_Id[idom][inode]] += db
    A_II = [julia_sparse(II[d][0], II[d][1], II[d][2], (n_Id[d], n_Id[d])) for d in range(ndom)]     # :540-543
    A_IG = [julia_sparse(IG[d][0], IG[d][1], IG[d][2], (n_Id[d], n_gd[d])) for d in range(ndom)]
    A_GG = [julia_sparse(GG[d][0], GG[d][1], GG[d][2], (n_gd[d], n_gd[d])) for d in range(ndom)]
    return LocalBlocks(A_II, A_IG, A_GG, b_I, b_G)


def julia_sparse(I, J, V, shape):
    """`sparse(I, J, V, m, n)` of Julia's SparseArrays (the call at :540-543): CSC with sorted rows in which the
    duplicates of an entry are combined with `+` in the order they appear in the triplet lists (sparse! counting-sorts
    the triplets by row -- a stable sort -- and folds repeated (row, column) pairs from left to right into the first
    one).  scipy's coo -> csc conversion makes no such promise, and for the diagonal entries (up to eight element
    contributions) the order decides the last bit."""
    acc = {}
    for i, j, v in zip(I, J, V):
        k = (int(j), int(i))
        acc[k] = acc[k] + v if k in acc else v
    keys = sorted(acc)
    indptr = np.zeros(shape[1] + 1, dtype=np.int64)
    for j, _ in keys:
        indptr[j + 1] += 1
    indptr = np.cumsum(indptr)
    indices = np.array([i for _, i in keys], dtype=np.int64)
    data = np.array([acc[k] for k in keys], dtype=np.float64)
    return sp.csc_matrix((data, indices, indptr), shape=shape)


def prepare_global_schur(cells, points, epart, sd: Subdomains, coeff, f_vals, u_dir):
    """prepare_global_schur (EllipticPdeDomainDecomposition.jl:212-369): same as the local
    version but A_IΓ,d has *global* interface columns and A_ΓΓ is the assembled block."""
    lb = prepare_local_schurs_ref(cells, points, epart, sd, coeff, f_vals, u_dir)
    n_g = sd.n_gamma
    A_IGg, A_GG = [], sp.csc_matrix((n_g, n_g))
    for d in range(sd.ndom):
        R = restriction(sd, d)
        A_IGg.append((lb.A_IG[d] @ R).tocsc())
        A_GG = A_GG + R.T @ lb.A_GG[d] @ R
    return lb.A_II, A_IGg, A_GG.tocsc(), lb.b_I, lb.b_G



def assemble_A_GG_from_local_blocks(A_GG, sd: Subdomains):
    """assemble_A_ΓΓ_from_local_blocks, EllipticPdeDomainDecomposition.jl:1080-1108."""
    n_g = sd.n_gamma
    out = sp.csc_matrix((n_g, n_g))
    for d in range(sd.ndom):
        R = restriction(sd, d)
        out = out + R.T @ A_GG[d] @ R
    return out.tocsc()


# ---------------------------------------------------------------------------------------------
# Interior solves, Schur operator, right-hand side, back-substitution
# ---------------------------------------------------------------------------------------------
class ExactInterior:
    """Exact A_II,d^{-1}: the reference's inner cg(A_IIdd, ., Pl=cholesky) at :648-650 converges
    in one step when Pl is an exact factor; this is that limit."""

    def __init__(self, A_II):
        self.lu = [spla.splu(A.tocsc()) if A.shape[0] else None for A in A_II]

    def solve(self, d, b):
        if self.lu[d] is None:
            return np.zeros_like(b)
        return self.lu[d].solve(b)


def apply_local_schur(lb: LocalBlocks, inner: ExactInterior, d, xd):
    """apply_local_schur, EllipticPdeDomainDecomposition.jl:639-654."""
    Sdxd = lb.A_GG[d] @ xd                                   # :646
    v = inner.solve(d, lb.A_IG[d] @ xd)                      # :648-650
    Sdxd -= lb.A_IG[d].T @ v                                 # :652
    return Sdxd


def apply_local_schurs(lb: LocalBlocks, inner: ExactInterior, sd: Subdomains, x):
    """Matrix-free apply_local_schurs, EllipticPdeDomainDecomposition.jl:711-747."""
    Sx = np.zeros(sd.n_gamma)                                # :721
    for d in range(sd.ndom):                                 # :723
        g = sd.gamma_l2g[d]
        xd = x[g]                                            # :728-730
        Sdxd = apply_local_schur(lb, inner, d, xd)
        Sx[g] += Sdxd                                        # :740-742 (each γ once per d)
    return Sx


def assemble_local_schurs(lb: LocalBlocks, inner: ExactInterior):
    """assemble_local_schurs, EllipticPdeDomainDecomposition.jl:667-695: dense S_d column by
    column, then Symmetric() (upper triangle mirrored, :692)."""
    out = []
    for d in range(len(lb.A_II)):
        n = lb.A_GG[d].shape[0]
        S = np.empty((n, n))
        for j in range(n):
            e = np.zeros(n)
            e[j] = 1.0
            S[:, j] = apply_local_schur(lb, inner, d, e)
        S = np.triu(S) + np.triu(S, 1).T
        out.append(S)
    return out


def apply_local_schurs_assembled(Sd, sd: Subdomains, x):
    """Assembled apply_local_schurs, EllipticPdeDomainDecomposition.jl:761-785."""
    Sx = np.zeros(sd.n_gamma)
    for d in range(sd.ndom):
        g = sd.gamma_l2g[d]
        Sx[g] += Sd[d] @ x[g]
    return Sx


def apply_global_schur(A_II, A_IGg, A_GG, inner: ExactInterior, x):
    """apply_global_schur, EllipticPdeDomainDecomposition.jl:596-625."""
    Sx = A_GG @ x                                            # :604
    for d in range(len(A_II)):
        v = inner.solve(d, A_IGg[d] @ x)                     # :617-619
        Sx -= A_IGg[d].T @ v                                 # :621
    return Sx


def get_schur_rhs(lb: LocalBlocks, inner: ExactInterior, sd: Subdomains):
    """get_schur_rhs (local-column variant), EllipticPdeDomainDecomposition.jl:835-864."""
    b_schur = lb.b_G.copy()                                  # :846-847
    for d in range(sd.ndom):
        v = inner.solve(d, lb.b_I[d])                        # :853-855
        w = lb.A_IG[d].T @ v                                 # :857
        b_schur[sd.gamma_l2g[d]] -= w                        # :858-860
    return b_schur


def get_subdomain_solutions(u_gamma, lb: LocalBlocks, inner: ExactInterior, sd: Subdomains):
    """get_subdomain_solutions, EllipticPdeDomainDecomposition.jl:1014-1025 (with local-column
    A_IΓ,d, i.e. A_IΓd * u_Γ == A_IΓdd * u_Γ[Γ_d])."""
    return [inner.solve(d, lb.b_I[d] - lb.A_IG[d] @ u_gamma[sd.gamma_l2g[d]])
            for d in range(sd.ndom)]


def merge_subdomain_solutions(u_gamma, u_I, sd: Subdomains, dirichlet_nodes, u_dir):
    """merge_subdomain_solutions, EllipticPdeDomainDecomposition.jl:1040-1070."""
    u = np.empty(sd.nnode)
    u[sd.node_gamma] = u_gamma                               # :1055-1057
    for d in range(sd.ndom):
        u[sd.node_Id[d]] = u_I[d]                            # :1059-1063
    u[dirichlet_nodes] = u_dir[dirichlet_nodes]              # :1065-1067
    return u


# ---------------------------------------------------------------------------------------------
# Neumann-Neumann preconditioner
# ---------------------------------------------------------------------------------------------
def prepare_neumann_neumann_schur_precond(Sd):
    """prepare_neumann_neumann_schur_precond (assembled variant),
    EllipticPdeDomainDecomposition.jl:1201-1220: pinv with rtol = sqrt(eps(Float64))."""
    rtol = np.sqrt(np.finfo(np.float64).eps)                 # :1209
    return [np.linalg.pinv(S, rcond=rtol) for S in Sd]      # :1211


def apply_neumann_neumann_schur(PiSd, sd: Subdomains, r):
    """apply_neumann_neumann_schur, EllipticPdeDomainDecomposition.jl:1361-1386."""
    z = np.zeros(sd.n_gamma)                                 # :1365
    cnt = sd.node_gamma_cnt
    for d in range(sd.ndom):                                 # :1367
        g = sd.gamma_l2g[d]
        rd = r[g] / cnt[g]                                   # :1373-1375
        t = PiSd[d] @ rd                                     # :1377
        z[g] += t / cnt[g]                                   # :1379-1381
    return z


# ---------------------------------------------------------------------------------------------
# Undecomposed system (cross-check of Example03:204)
# ---------------------------------------------------------------------------------------------
def do_isotropic_elliptic_assembly(cells, points, is_dirichlet, coeff, f_vals, u_dir):
    """Literal element loop of do_isotropic_elliptic_assembly, Fem/EllipticPde.jl:187-275 (0-based):
    global stiffness on the non-Dirichlet nodes, Dirichlet lift into the right-hand side (:253-254),
    load-vector rule (2 f_i + f_j + f_k) * Area / 12 (:259-271).  Pure Python, small meshes only.
    Returns A (csc), b, free-node list."""
    nnode = points.shape[1]
    nel = cells.shape[1]
    isd = np.asarray(is_dirichlet, bool)
    free = np.nonzero(~isd)[0]
    g2l = {int(nd): k for k, nd in enumerate(free)}                      # not_dirichlet_inds_g2l
    I, J, V = [], [], []
    b = np.zeros(free.size)
    x, y = np.zeros(3), np.zeros(3)
    dx, dy = np.zeros(3), np.zeros(3)
    for iel in range(nel):                                               # :209
        nodes = [int(v) for v in cells[:, iel]]
        da = 0.0
        for j, jn in enumerate(nodes):                                   # :214-218
            x[j], y[j] = points[0, jn], points[1, jn]
            da += coeff[jn]
        da /= 3.0
        dx[0], dx[1], dx[2] = x[2] - x[1], x[0] - x[2], x[1] - x[0]      # :221-226
        dy[0], dy[1], dy[2] = y[1] - y[2], y[2] - y[0], y[0] - y[1]
        area = (dx[2] * dy[1] - dx[1] * dy[2]) / 2.0                     # :229
        for i, inode in enumerate(nodes):                                # :232
            for j, jnode in enumerate(nodes):
                kij = da * (dy[i] * dy[j] + dx[i] * dx[j]) / 4 / area    # :241
                if (not isd[inode]) and (not isd[jnode]):                # :247
                    I.append(g2l[inode]); J.append(g2l[jnode]); V.append(kij)
                elif isd[inode] and not isd[jnode]:                      # :253
                    b[g2l[jnode]] -= u_dir[inode] * kij
        for i, inode in enumerate(nodes):                                # :259
            if not isd[inode]:
                j, k = (i + 1) % 3, (i + 2) % 3
                b[g2l[inode]] += (2 * f_vals[nodes[i]] + f_vals[nodes[j]] + f_vals[nodes[k]]) * area / 12
    A = sp.coo_matrix((V, (I, J)), shape=(free.size, free.size)).tocsc()  # sparse(I, J, V) sums duplicates
    A.sum_duplicates()
    return A, b, free

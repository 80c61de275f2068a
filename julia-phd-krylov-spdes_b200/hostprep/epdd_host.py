"""Host-side preparation of the Schur-complement solve: subdomain index maps and block assembly.

This is the part of the hot path that stays on the host (north_star: "Mesh generation, METIS and
FEM assembly stay on the host in Julia"); these numpy-vectorised functions mirror the reference's
Fem API names and outputs (Fem/EllipticPdeDomainDecomposition.jl:86-193 set_subdomains, :389-546
prepare_local_schurs) so the Python harness can feed libkrylov_b200.so exactly what the Julia
shim would.  They are checked bit-for-bit against the literal transcription in oracle/epdd.py by
tests/test_oracle_maps.py.  Everything is 0-based except ``node_owner`` (reference values).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np
import scipy.sparse as sp

@dataclass
class Subdomains:
    """Outputs of set_subdomains (EllipticPdeDomainDecomposition.jl:191-192), as arrays.

    gd_nodes[d][l]   global node of local interface index l of subdomain d
                     (inverse of the reference's Dict ind_Γd_g2l[d])
    gamma_l2g[d][l]  interface id (in Γ) of local interface index l
                     (inverse of ind_Γd_Γ2l[d]; cf. ind_Γd_l2Γ at :1090-1095)
    ind_gamma_g2l    nnode array, interface id of a node or -1 (Dict ind_Γ_g2l)
    node_owner       reference values: -1 interface, 0 Dirichlet, d+1 interior of d
    """
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

    def to_julia(self):
        """1-based values exactly as the reference would hold them."""
        return {
            "node_owner": self.node_owner.copy(),
            "node_Γ": self.node_gamma + 1,
            "node_Γ_cnt": self.node_gamma_cnt.copy(),
            "ind_Γd_l2Γ": [g + 1 for g in self.gamma_l2g],
            "ind_Γd_l2g": [g + 1 for g in self.gd_nodes],
            "node_Id": [g + 1 for g in self.node_Id],
        }


def set_subdomains(cells, cell_neighbors, epart, npart, is_dirichlet):
    """Vectorised set_subdomains with the reference's numbering
    (EllipticPdeDomainDecomposition.jl:86-193): interface nodes numbered by first
    encounter in the (element, edge, vertex) traversal (:123-163), local interface
    nodes by first encounter within the subdomain's own elements (:148-150),
    interior nodes by ascending global id within npart (:166-180)."""
    nel = cells.shape[1]
    nnode = int(cells.max()) + 1
    ndom = int(epart.max()) + 1
    is_dirichlet = np.asarray(is_dirichlet, dtype=bool)
    valid = cell_neighbors >= 0
    jel = np.where(valid, cell_neighbors, 0)
    diff = valid & (epart[jel] != epart[None, :])            # (3 j, nel)
    sel = np.nonzero(diff.any(axis=0))[0]                    # interface elements, ascending
    ic = cells[:, sel]                                       # (3 k, ns)
    jc = cells[:, jel[:, sel]]                               # (3 m, 3 j, ns)
    common = (ic[:, None, None, :] == jc[None, :, :, :]).any(axis=1)   # (3 k, 3 j, ns)
    hit = common & diff[None, :, sel] & ~is_dirichlet[ic][:, None, :]
    hit = np.transpose(hit, (2, 1, 0))                       # (ns, j, k): traversal order
    ev = np.nonzero(hit.ravel())[0]
    e_idx = ev // 9
    k_idx = ev % 3
    ev_node = ic[k_idx, e_idx]
    ev_dom = epart[sel[e_idx]]
    # Γ numbering: order of first occurrence
    uniq, first = np.unique(ev_node, return_index=True)
    node_gamma = uniq[np.argsort(first, kind="stable")]
    n_gamma = node_gamma.size
    ind_gamma = np.full(nnode, -1, dtype=np.int64)
    ind_gamma[node_gamma] = np.arange(n_gamma)
    node_owner = np.zeros(nnode, dtype=np.int64)
    node_owner[node_gamma] = -1
    # Γ_d numbering: order of first occurrence of (dom, node)
    key = ev_dom * np.int64(nnode) + ev_node
    ukey, kfirst = np.unique(key, return_index=True)
    kd = ukey // nnode
    kn = ukey % nnode
    order = np.lexsort((kfirst, kd))
    kd, kn = kd[order], kn[order]
    ptr = np.searchsorted(kd, np.arange(ndom + 1))
    gd_nodes = [kn[ptr[d]:ptr[d + 1]].copy() for d in range(ndom)]
    gamma_l2g = [ind_gamma[g] for g in gd_nodes]
    cnt = np.bincount(ind_gamma[kn], minlength=n_gamma).astype(np.int64)
    # interior nodes
    inter = np.nonzero(~is_dirichlet & (node_owner != -1))[0]
    node_owner[inter] = npart[inter] + 1
    o = np.argsort(npart[inter], kind="stable")
    inter_s = inter[o]
    iptr = np.searchsorted(npart[inter_s], np.arange(ndom + 1))
    node_Id = [inter_s[iptr[d]:iptr[d + 1]].copy() for d in range(ndom)]
    eo = np.argsort(epart, kind="stable")
    eptr = np.searchsorted(epart[eo], np.arange(ndom + 1))
    elemd = [eo[eptr[d]:eptr[d + 1]] for d in range(ndom)]
    return Subdomains(ndom, nnode, node_owner, node_gamma, cnt, ind_gamma,
                      gd_nodes, gamma_l2g, node_Id, elemd)



@dataclass
class LocalBlocks:
    """Outputs of prepare_local_schurs (:545): per-subdomain CSC blocks and right-hand sides."""
    A_II: List[sp.csc_matrix]
    A_IG: List[sp.csc_matrix]     # n_I,d x n_Γd, local interface columns
    A_GG: List[sp.csc_matrix]     # n_Γd x n_Γd, the subdomain's partial (Neumann) interface block
    b_I: List[np.ndarray]
    b_G: np.ndarray


def _element_geometry(cells, points, coeff):
    """Per-element P1 stiffness K[e,i,j] and area, EllipticPdeDomainDecomposition.jl:437-470."""
    x = points[0][cells]                                     # (3, nel)
    y = points[1][cells]
    da = (0.0 + coeff[cells[0]] + coeff[cells[1]] + coeff[cells[2]]) / 3.0     # :437-446
    dx = np.stack([x[2] - x[1], x[0] - x[2], x[1] - x[0]])   # :449-451
    dy = np.stack([y[1] - y[2], y[2] - y[0], y[0] - y[1]])   # :452-454
    area = (dx[2] * dy[1] - dx[1] * dy[2]) / 2.0             # :457
    K = (da[None, None, :] * (dy[:, None, :] * dy[None, :, :] + dx[:, None, :] * dx[None, :, :])
         / 4 / area[None, None, :])                          # :470  (3 i, 3 j, nel)
    return K, area


def _split_blocks(M, rptr, cptr):
    """Diagonal blocks of a block-diagonal CSC matrix given row / column offsets."""
    M = M.tocsc()
    M.sum_duplicates()
    M.sort_indices()
    out = []
    for d in range(len(rptr) - 1):
        c0, c1 = cptr[d], cptr[d + 1]
        p0, p1 = M.indptr[c0], M.indptr[c1]
        blk = sp.csc_matrix((M.data[p0:p1].copy(), M.indices[p0:p1] - rptr[d],
                             M.indptr[c0:c1 + 1] - p0),
                            shape=(rptr[d + 1] - rptr[d], c1 - c0))
        out.append(blk)
    return out


def prepare_local_schurs(cells, points, epart, sd: Subdomains, coeff, f_vals, u_dir):
    """Vectorised prepare_local_schurs (EllipticPdeDomainDecomposition.jl:389-546).

    Same routing of the element stiffness terms (:472-497), Dirichlet lift (:499-508)
    and load vector (:514-536); right-hand sides are accumulated in the reference's
    order (element-major; lift terms then load terms), so b_I / b_G are bitwise equal
    to prepare_local_schurs_ref."""
    ndom, nnode = sd.ndom, sd.nnode
    nel = cells.shape[1]
    K, area = _element_geometry(cells, points, coeff)
    owner = sd.node_owner
    n_Id = np.array([v.size for v in sd.node_Id])
    n_gd = np.array([v.size for v in sd.gd_nodes])
    iptr = np.concatenate([[0], np.cumsum(n_Id)])
    gptr = np.concatenate([[0], np.cumsum(n_gd)])
    # concatenated interior index of every interior node
    cat_I = np.full(nnode, -1, dtype=np.int64)
    for d in range(ndom):
        cat_I[sd.node_Id[d]] = iptr[d] + np.arange(n_Id[d])
    # concatenated local-interface index of (dom, node): sorted key table
    gk = np.concatenate([d * np.int64(nnode) + sd.gd_nodes[d] for d in range(ndom)]) \
        if ndom else np.zeros(0, np.int64)
    gv = np.arange(gk.size, dtype=np.int64)
    go = np.argsort(gk)
    gk_s, gv_s = gk[go], gv[go]

    def cat_G(dom, node):
        key = dom * np.int64(nnode) + node
        pos = np.searchsorted(gk_s, key)
        pos = np.minimum(pos, gk_s.size - 1)
        if not np.all(gk_s[pos] == key):
            # the reference would raise KeyError at :478/:495 for such a mesh/partition
            raise KeyError("interface node of an element is missing from its subdomain's Γ_d")
        return gv_s[pos]

    en = cells                                              # (3, nel)
    oi = owner[en]                                          # (3, nel)
    dom = np.broadcast_to(epart[None, None, :], (3, 3, nel))
    ni = np.broadcast_to(en[:, None, :], (3, 3, nel))
    nj = np.broadcast_to(en[None, :, :], (3, 3, nel))
    o_i = np.broadcast_to(oi[:, None, :], (3, 3, nel))
    o_j = np.broadcast_to(oi[None, :, :], (3, 3, nel))

    m_gg = (o_i == -1) & (o_j == -1)
    m_ii = (o_i > 0) & (o_j > 0)
    m_ig = (o_i > 0) & (o_j == -1)
    nI, nG = int(iptr[-1]), int(gptr[-1])
    A_II_all = sp.coo_matrix((K[m_ii], (cat_I[ni[m_ii]], cat_I[nj[m_ii]])), shape=(nI, nI))
    A_IG_all = sp.coo_matrix((K[m_ig], (cat_I[ni[m_ig]], cat_G(dom[m_ig], nj[m_ig]))), shape=(nI, nG))
    A_GG_all = sp.coo_matrix((K[m_gg], (cat_G(dom[m_gg], ni[m_gg]), cat_G(dom[m_gg], nj[m_gg]))),
                             shape=(nG, nG))
    A_II = _split_blocks(A_II_all, iptr, iptr)
    A_IG = _split_blocks(A_IG_all, iptr, gptr)
    A_GG = _split_blocks(A_GG_all, gptr, gptr)

    # right-hand sides: one accumulation stream in the reference's order
    # target index space: [0, n_gamma) for b_Γ, n_gamma + cat_I for b_I
    n_g = sd.n_gamma
    tgt_node = np.where(owner == -1, sd.ind_gamma_g2l, np.where(owner > 0, n_g + cat_I, -1))
    lift_mask = (o_i == 0) & (o_j != 0)                     # (i, j, nel)
    lift_val = -(K * u_dir[ni])
    lift_tgt = np.where(lift_mask, tgt_node[nj], -1)
    fi = f_vals[en]
    db = (2 * fi + fi[[1, 2, 0]] + fi[[2, 0, 1]]) * area[None, :] / 12     # (3, nel)
    load_tgt = np.where(oi != 0, tgt_node[en], -1)
    # per element: 9 lift terms in (i outer, j inner) order, then 3 load terms
    tgt = np.concatenate([np.transpose(lift_tgt, (2, 0, 1)).reshape(nel, 9),
                          load_tgt.T], axis=1).ravel()
    val = np.concatenate([np.transpose(lift_val, (2, 0, 1)).reshape(nel, 9),
                          db.T], axis=1).ravel()
    keep = tgt >= 0
    acc = np.bincount(tgt[keep], weights=val[keep], minlength=n_g + nI)
    b_G = acc[:n_g].copy()
    b_I = [acc[n_g + iptr[d]:n_g + iptr[d + 1]].copy() for d in range(ndom)]
    return LocalBlocks(A_II, A_IG, A_GG, b_I, b_G)



@dataclass
class AssemblyMaps:
    """Gather form of the element loop of prepare_local_schurs (:430-497) for kb200_upload_assembly_map: per subdomain
    and block (0: A_II, 1: A_IΓ, 2: A_ΓΓ) and per stored entry of the block (CSC order, rows sorted) the contributions
    src[ptr[p]:ptr[p+1]] (0-based here; the ABI takes both 1-based), each the position 9 e + 3 i + j of a contribution in
    the reference's loop nest (element e, vertex i outer, vertex j inner), ascending = the order in which sparse()
    (:540-543) adds the duplicates of an entry."""
    ptr: List[List[np.ndarray]]     # [d][which] -> int64 (nnz + 1)
    src: List[List[np.ndarray]]     # [d][which] -> int64 (number of contributions)


def assembly_maps(cells, epart, sd: Subdomains, dom_range=None) -> AssemblyMaps:
    """Contribution lists of every stored entry of A_II,d / A_IΓ,d / A_ΓΓ,d (same routing as prepare_local_schurs,
    :472-497; purely structural: mesh, partition and index maps, no coefficient)."""
    ndom, nnode = sd.ndom, sd.nnode
    nel = cells.shape[1]
    owner = sd.node_owner
    n_Id = np.array([v.size for v in sd.node_Id])
    n_gd = np.array([v.size for v in sd.gd_nodes])
    iptr = np.concatenate([[0], np.cumsum(n_Id)])
    gptr = np.concatenate([[0], np.cumsum(n_gd)])
    cat_I = np.full(nnode, -1, dtype=np.int64)
    for d in range(ndom):
        cat_I[sd.node_Id[d]] = iptr[d] + np.arange(n_Id[d])
    gk = np.concatenate([d * np.int64(nnode) + sd.gd_nodes[d] for d in range(ndom)]) if ndom else np.zeros(0, np.int64)
    go = np.argsort(gk)
    gk_s = gk[go]

    def cat_G(dom, node):
        key = dom * np.int64(nnode) + node
        pos = np.minimum(np.searchsorted(gk_s, key), gk_s.size - 1)
        if not np.all(gk_s[pos] == key):
            raise KeyError("interface node of an element is missing from its subdomain's Γ_d")
        return go[pos]

    e_ids = np.arange(nel, dtype=np.int64)
    if dom_range is not None:
        e_ids = e_ids[(epart >= dom_range[0]) & (epart < dom_range[1])]
    en = cells[:, e_ids]                                     # (3, m)
    oi = owner[en]
    m = e_ids.size
    ni = np.broadcast_to(en[:, None, :], (3, 3, m))
    nj = np.broadcast_to(en[None, :, :], (3, 3, m))
    o_i = np.broadcast_to(oi[:, None, :], (3, 3, m))
    o_j = np.broadcast_to(oi[None, :, :], (3, 3, m))
    dom = np.broadcast_to(epart[e_ids][None, None, :], (3, 3, m))
    cid = (9 * e_ids[None, None, :] + 3 * np.arange(3)[:, None, None] + np.arange(3)[None, :, None])   # (i, j, e)

    def lists(mask, rows, cols, rptr, cptr):
        c = cid[mask]
        o = np.lexsort((c, rows, cols))                      # column, then row, then contribution order
        r_s, c_s, id_s = rows[o], cols[o], c[o]
        new = np.ones(o.size, dtype=bool)
        new[1:] = (r_s[1:] != r_s[:-1]) | (c_s[1:] != c_s[:-1])
        start = np.flatnonzero(new)                          # first contribution of every stored entry
        ent_col = c_s[start]
        ptr_all = np.concatenate([start, [o.size]]).astype(np.int64)
        out_p, out_s = [], []
        e_lo = np.searchsorted(ent_col, cptr[:-1], side="left")
        e_hi = np.searchsorted(ent_col, cptr[1:], side="left")
        for d in range(ndom):
            a, b = e_lo[d], e_hi[d]
            pp = ptr_all[a:b + 1] - ptr_all[a]
            out_p.append(pp.copy())
            out_s.append(id_s[ptr_all[a]:ptr_all[b]].astype(np.int64))
        return out_p, out_s

    m_gg = (o_i == -1) & (o_j == -1)
    m_ii = (o_i > 0) & (o_j > 0)
    m_ig = (o_i > 0) & (o_j == -1)
    p0, s0 = lists(m_ii, cat_I[ni[m_ii]], cat_I[nj[m_ii]], iptr, iptr)
    p1, s1 = lists(m_ig, cat_I[ni[m_ig]], cat_G(dom[m_ig], nj[m_ig]), iptr, gptr)
    p2, s2 = lists(m_gg, cat_G(dom[m_gg], ni[m_gg]), cat_G(dom[m_gg], nj[m_gg]), gptr, gptr)
    return AssemblyMaps([[p0[d], p1[d], p2[d]] for d in range(ndom)], [[s0[d], s1[d], s2[d]] for d in range(ndom)])


def assemble_values_from_map(cells, points, coeff, ptr, src):
    """What kb200_assemble_realization computes for one block, on the host (numpy): ΔK of every contribution as :437-470,
    duplicates added in ascending contribution order.  Test twin of kernels.cuh::assemble_values_kernel."""
    e = src // 9
    i = (src % 9) // 3
    j = src % 3
    ue, inv = np.unique(e, return_inverse=True)
    K, _ = _element_geometry(cells[:, ue], points, coeff)
    v = K[i, j, inv]
    n = ptr.size - 1
    cnt = np.diff(ptr)
    acc = np.zeros(n)
    has = cnt > 0
    acc[has] = v[ptr[:-1][has]]
    for r in range(1, int(cnt.max()) if n else 0):
        mk = cnt > r
        acc[mk] = acc[mk] + v[ptr[:-1][mk] + r]
    return acc


@dataclass
class RhsMaps:
    """Arguments of kb200_upload_rhs_map (0-based here): the load vector (bI0 concatenated over the subdomains, bG0) and,
    for the entries that carry Dirichlet-lift terms (:499-508), their index `tgt` in [b_I ; b_Γ] and their terms
    src/val[ptr[t]:ptr[t+1]] in the order of the reference's loop nest -- src >= 0: `b -= ΔK_src * val` (val = uexact at
    the Dirichlet vertex), src = -1: `b += val` (val = Δb of :524-526)."""
    bI0: np.ndarray
    bG0: np.ndarray
    tgt: np.ndarray
    ptr: np.ndarray
    src: np.ndarray
    val: np.ndarray


def rhs_maps(cells, points, epart, sd: Subdomains, f_vals, u_dir) -> RhsMaps:
    """Right-hand-side part of prepare_local_schurs (:499-536) in gather form.  Only entries of nodes that share an
    element with a Dirichlet node depend on the coefficient; everything else is the load vector."""
    nnode, owner = sd.nnode, sd.node_owner
    is_dir = owner == 0
    el_dir = is_dir[cells].any(axis=0)
    tnode = np.zeros(nnode, dtype=bool)
    tnode[cells[:, el_dir].ravel()] = True
    tnode &= ~is_dir
    layer = np.flatnonzero(tnode[cells].any(axis=0))          # every element of such a node, ascending
    static = prepare_local_schurs(cells, points, epart, sd, np.ones(nnode), f_vals, u_dir)
    bI0 = np.concatenate(static.b_I) if sd.ndom else np.zeros(0)
    n_Id = np.array([v.size for v in sd.node_Id])
    iptr = np.concatenate([[0], np.cumsum(n_Id)])
    cat_I = np.full(nnode, -1, dtype=np.int64)
    for d in range(sd.ndom):
        cat_I[sd.node_Id[d]] = iptr[d] + np.arange(n_Id[d])
    tidx = np.where(owner == -1, iptr[-1] + sd.ind_gamma_g2l, np.where(owner > 0, cat_I, -1))
    en = cells[:, layer]                                       # (3, m)
    oi = owner[en]
    m = layer.size
    ii = np.arange(3)[:, None, None]
    jj = np.arange(3)[None, :, None]
    lift = (oi[:, None, :] == 0) & (oi[None, :, :] != 0) & tnode[en][None, :, :]          # (i, j, e): :499
    shape = (3, 3, m)
    l_tgt = np.broadcast_to(tidx[en][None, :, :], shape)[lift]
    l_src = np.broadcast_to(9 * layer[None, None, :] + 3 * ii + jj, shape)[lift]
    l_val = np.broadcast_to(u_dir[en][:, None, :], shape)[lift]
    l_seq = np.broadcast_to(12 * layer[None, None, :] + 3 * ii + jj, shape)[lift]
    _, area = _element_geometry(en, points, np.ones(nnode))
    fi = f_vals[en]
    db = (2 * fi + fi[[1, 2, 0]] + fi[[2, 0, 1]]) * area[None, :] / 12                   # :524-526
    load = (oi != 0) & tnode[en]
    d_tgt = tidx[en][load]
    d_val = db[load]
    d_seq = np.broadcast_to(12 * layer[None, :] + 9 + np.arange(3)[:, None], (3, m))[load]
    tg = np.concatenate([l_tgt, d_tgt])
    sr = np.concatenate([l_src, np.full(d_tgt.size, -1, dtype=np.int64)])
    vl = np.concatenate([l_val, d_val])
    sq = np.concatenate([l_seq, d_seq])
    o = np.lexsort((sq, tg))
    tg, sr, vl = tg[o], sr[o], vl[o]
    new = np.ones(tg.size, dtype=bool)
    new[1:] = tg[1:] != tg[:-1]
    start = np.flatnonzero(new)
    return RhsMaps(bI0, static.b_G, tg[start].astype(np.int64), np.concatenate([start, [tg.size]]).astype(np.int64),
                   sr.astype(np.int64), vl.astype(np.float64))


def assemble_rhs_from_map(cells, points, coeff, m: RhsMaps):
    """What kb200_assemble_realization computes for the right-hand sides, on the host (numpy): the load vector with the
    entries of `m.tgt` recomputed term by term in order.  Test twin of kernels.cuh::assemble_rhs_kernel."""
    out = np.concatenate([m.bI0, m.bG0])
    lift = m.src >= 0
    term = m.val.copy()
    if lift.any():
        c = m.src[lift]
        ue, inv = np.unique(c // 9, return_inverse=True)
        K, _ = _element_geometry(cells[:, ue], points, coeff)
        term[lift] = -(K[(c % 9) // 3, c % 3, inv] * m.val[lift])
    cnt = np.diff(m.ptr)
    acc = np.zeros(m.tgt.size)
    for r in range(int(cnt.max()) if cnt.size else 0):
        mk = cnt > r
        acc[mk] = acc[mk] + term[m.ptr[:-1][mk] + r]
    out[m.tgt] = acc
    return out[:m.bI0.size], out[m.bI0.size:]


def restriction(sd: Subdomains, d):
    """R_d: n_Γd x n_Γ boolean gather matrix (the Dict loops at :728-730)."""
    g = sd.gamma_l2g[d]
    return sp.csr_matrix((np.ones(g.size), (np.arange(g.size), g)), shape=(g.size, sd.n_gamma))


"""Index maps and block assembly: the vectorised host preparation must reproduce the literal
transcription of the reference (EllipticPdeDomainDecomposition.jl:86-193, :389-546) bit for bit."""
import numpy as np
import pytest

import oracle.epdd as E
from krylov_b200.hostprep import epdd_host as H
from krylov_b200.hostprep import mesh as M


def _case(nx, ndom, jitter):
    cells, pts, mk, nb = M.structured_mesh(nx, jitter=jitter)
    epart, npart = M.rcb_partition(cells, pts, ndom)
    return cells, pts, mk, nb, epart, npart


@pytest.mark.parametrize("nx,ndom,jitter", [(9, 2, 0.0), (17, 4, 0.0), (25, 5, 0.2), (33, 16, 0.2), (29, 7, 0.1)])
def test_set_subdomains_matches_literal_transcription(nx, ndom, jitter):
    cells, pts, mk, nb, epart, npart = _case(nx, ndom, jitter)
    a = E.set_subdomains_ref(cells, nb, epart, npart, mk == 1)
    b = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    assert np.array_equal(a.node_owner, b.node_owner)
    assert np.array_equal(a.node_gamma, b.node_gamma)
    assert np.array_equal(a.node_gamma_cnt, b.node_gamma_cnt)
    assert np.array_equal(a.ind_gamma_g2l, b.ind_gamma_g2l)
    for d in range(ndom):
        assert np.array_equal(a.gd_nodes[d], b.gd_nodes[d])
        assert np.array_equal(a.gamma_l2g[d], b.gamma_l2g[d])
        assert np.array_equal(a.node_Id[d], b.node_Id[d])
        assert np.array_equal(a.elemd[d], b.elemd[d])


def test_interface_properties():
    """Γ = non-Dirichlet nodes whose elements span several subdomains; node_Γ_cnt = how many."""
    cells, pts, mk, nb, epart, npart = _case(33, 16, 0.2)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    nnode = pts.shape[1]
    touch = [set() for _ in range(nnode)]
    for e in range(cells.shape[1]):
        for v in cells[:, e]:
            touch[v].add(int(epart[e]))
    expect = np.array([len(t) if (len(t) > 1 and mk[i] == 0) else 0 for i, t in enumerate(touch)])
    got = np.zeros(nnode, dtype=int)
    got[sd.node_gamma] = sd.node_gamma_cnt
    assert np.array_equal(expect, got)
    assert sd.node_gamma_cnt.max() <= 4
    # interior nodes are numbered by ascending global id inside their subdomain (:166-180)
    for d in range(sd.ndom):
        assert np.all(np.diff(sd.node_Id[d]) > 0)
    ju = sd.to_julia()
    assert ju["node_Γ"].min() >= 1 and all(g.min() >= 1 for g in ju["ind_Γd_l2Γ"])


@pytest.mark.parametrize("nx,ndom,jitter", [(17, 4, 0.0), (25, 5, 0.2)])
def test_prepare_local_schurs_matches_literal_transcription(nx, ndom, jitter):
    cells, pts, mk, nb, epart, npart = _case(nx, ndom, jitter)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    coeff = np.exp(M.lognormal_field(pts, 3))
    f = -np.ones(pts.shape[1])
    u = 0.734 * np.ones(pts.shape[1])
    la = E.prepare_local_schurs_ref(cells, pts, epart, sd, coeff, f, u)
    lb = H.prepare_local_schurs(cells, pts, epart, sd, coeff, f, u)
    for d in range(ndom):
        for x, y in ((la.A_II[d], lb.A_II[d]), (la.A_IG[d], lb.A_IG[d]), (la.A_GG[d], lb.A_GG[d])):
            assert x.shape == y.shape
            assert (x != y).nnz == 0 or abs(x - y).max() <= 4e-16 * abs(x).max()
        assert np.array_equal(la.b_I[d], lb.b_I[d])          # same accumulation order
    assert np.array_equal(la.b_G, lb.b_G)


def test_missing_gamma_d_node_raises_like_the_reference():
    """A partition where an element touches an interface node only through a vertex makes the
    reference fail with KeyError at :478/:495; the host preparation must refuse it too."""
    cells, pts, mk, nb, epart, npart = _case(9, 2, 0.0)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    # corrupt: drop one local interface node of subdomain 0
    sd.gd_nodes[0] = sd.gd_nodes[0][:-1]
    with pytest.raises(KeyError):
        H.prepare_local_schurs(cells, pts, epart, sd, np.ones(pts.shape[1]), -np.ones(pts.shape[1]),
                               np.zeros(pts.shape[1]))


def test_wire_formats_round_trip(tmp_path):
    """Mesh / partition files in the reference's NPZ layout (Fem/Mesh.jl:49-91, 221-258)."""
    cells, pts, mk, nb, epart, npart = _case(9, 3, 0.1)
    M.save_mesh(str(tmp_path), 81, cells, pts, mk, nb)
    M.save_partition(str(tmp_path), 81, 3, epart, npart)
    c2, p2, m2, n2 = M.load_mesh(str(tmp_path), 81)
    e2, np2 = M.load_partition(str(tmp_path), 81, 3)
    assert np.array_equal(c2, cells) and np.array_equal(p2, pts) and np.array_equal(m2, mk)
    assert np.array_equal(n2, nb) and np.array_equal(e2, epart) and np.array_equal(np2, npart)
    raw = np.load(open(tmp_path / "DoF81.cells.npz", "rb"))
    assert raw.shape == (cells.shape[1], 3) and raw.min() == 0       # nel x 3, 0-based (Mesh.jl:52)
    rawn = np.load(open(tmp_path / "DoF81.cell_neighbors.npz", "rb"))
    assert rawn.min() == -1 and rawn.max() == cells.shape[1]         # in-memory values (Mesh.jl:55)


def _region_growing_partition(cells, nb, ndom, seed, npart_rule="min"):
    """METIS-like stand-in: contiguous, jagged element parts grown breadth-first from random seed elements
    over the edge-adjacency graph (round-robin, so sizes stay comparable); npart = the part of one of the
    node's elements (lowest- or highest-numbered: mpmetis does not promise either)."""
    rng = np.random.default_rng(seed)
    nel = cells.shape[1]
    epart = -np.ones(nel, dtype=np.int64)
    fronts = []
    for d, e in enumerate(rng.choice(nel, size=ndom, replace=False)):
        epart[e] = d
        fronts.append([int(e)])
    left = nel - ndom
    while left:
        progressed = False
        for d in range(ndom):
            nxt = []
            grew = False
            for e in fronts[d]:
                for m in nb[:, e]:
                    if m >= 0 and epart[m] < 0 and not grew:
                        epart[m] = d
                        nxt.append(int(m))
                        left -= 1
                        grew = True
                nxt.append(e)
                if grew:
                    break
            fronts[d] = nxt + [e for e in fronts[d] if e not in nxt]
            progressed |= grew
        if not progressed:        # enclosed leftovers: attach to any assigned neighbour
            for e in np.flatnonzero(epart < 0):
                for m in nb[:, e]:
                    if m >= 0 and epart[m] >= 0:
                        epart[e] = epart[m]
                        left -= 1
                        break
    nnode = int(cells.max()) + 1
    npart = np.full(nnode, -1, dtype=np.int64)
    order = range(nel) if npart_rule == "max" else range(nel - 1, -1, -1)
    for e in order:
        npart[cells[:, e]] = epart[e]
    return epart, npart


@pytest.mark.parametrize("seed,ndom,rule", [(0, 5, "min"), (1, 9, "max"), (2, 12, "min")])
def test_maps_and_blocks_on_jagged_partitions(seed, ndom, rule):
    """Irregular (METIS-like) partitions: jagged interfaces, nodes shared by 3+ subdomains, npart not tied to
    the lowest element -- the vectorised host preparation still equals the literal transcription bit for
    bit, and the decomposed operator still reproduces the undecomposed solve (Example03:204)."""
    import scipy.sparse.linalg as spla
    cells, pts, mk, nb = M.structured_mesh(21, jitter=0.15)
    epart, npart = _region_growing_partition(cells, nb, ndom, seed, rule)
    assert epart.min() == 0 and epart.max() == ndom - 1 and (npart >= 0).all()
    a = E.set_subdomains_ref(cells, nb, epart, npart, mk == 1)
    b = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    assert np.array_equal(a.node_owner, b.node_owner) and np.array_equal(a.node_gamma, b.node_gamma)
    assert np.array_equal(a.node_gamma_cnt, b.node_gamma_cnt) and a.node_gamma_cnt.max() >= 3
    for d in range(ndom):
        assert np.array_equal(a.gamma_l2g[d], b.gamma_l2g[d]) and np.array_equal(a.node_Id[d], b.node_Id[d])
    coeff = np.exp(M.lognormal_field(pts, 5))
    f, u = -np.ones(pts.shape[1]), 0.734 * np.ones(pts.shape[1])
    la = E.prepare_local_schurs_ref(cells, pts, epart, b, coeff, f, u)
    lb = H.prepare_local_schurs(cells, pts, epart, b, coeff, f, u)
    for d in range(ndom):
        for x, y in ((la.A_II[d], lb.A_II[d]), (la.A_IG[d], lb.A_IG[d]), (la.A_GG[d], lb.A_GG[d])):
            assert x.shape == y.shape and ((x != y).nnz == 0 or abs(x - y).max() <= 4e-16 * abs(x).max())
    # DD solution == undecomposed solution on this partition
    inner = E.ExactInterior(lb.A_II)
    bs = E.get_schur_rhs(lb, inner, b)
    n_g = b.n_gamma
    Smat = np.column_stack([E.apply_local_schurs(lb, inner, b, e) for e in np.eye(n_g)])
    ug = np.linalg.solve(Smat, bs)
    uI = E.get_subdomain_solutions(ug, lb, inner, b)
    u_dd = E.merge_subdomain_solutions(ug, uI, b, np.flatnonzero(mk == 1), u)
    A, rhs, free = E.do_isotropic_elliptic_assembly(cells, pts, mk == 1, coeff, f, u)
    u_full = u.copy()
    u_full[free] = spla.spsolve(A.tocsc(), rhs)
    assert np.abs(u_dd - u_full).max() <= 1e-10 * np.abs(u_full).max()

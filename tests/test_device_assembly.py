"""Re-assembly of the local blocks on the device (row f1: kb200_upload_assembly_mesh / kb200_upload_assembly_map /
kb200_assemble_realization; the element loop of prepare_local_schurs, EllipticPdeDomainDecomposition.jl:430-497).

CPU part: the contribution lists the host hands to the library (hostprep.epdd_host.assembly_maps) reproduce the literal
transcription of the reference -- triplets pushed element by element, duplicates combined by sparse() in push order --
BIT FOR BIT through the numpy twin of the device kernel, and the right-hand sides of a new coefficient come out of the
boundary layer alone.  GPU part: the kernel against that twin, the oracle's operator, and a whole realization loop."""
import numpy as np
import pytest

import oracle.epdd as E
import oracle.solvers as OS
from krylov_b200.hostprep import epdd_host as H
from krylov_b200.hostprep import mesh as M
from krylov_b200.hostprep import problem as PB


def _case(nx, ndom, jitter):
    cells, pts, mk, nb = M.structured_mesh(nx, jitter=jitter)
    epart, npart = M.rcb_partition(cells, pts, ndom)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    return cells, pts, epart, sd


@pytest.mark.parametrize("nx,ndom,jitter", [(17, 4, 0.0), (25, 5, 0.2), (33, 9, 0.2)])
def test_assembly_maps_reproduce_the_literal_transcription_bit_for_bit(nx, ndom, jitter):
    cells, pts, epart, sd = _case(nx, ndom, jitter)
    coeff = np.exp(M.lognormal_field(pts, 3))
    f, u = -np.ones(pts.shape[1]), 0.734 * np.ones(pts.shape[1])
    ref = E.prepare_local_schurs_ref(cells, pts, epart, sd, coeff, f, u)
    maps = H.assembly_maps(cells, epart, sd)
    for d in range(ndom):
        for which, A in enumerate((ref.A_II[d], ref.A_IG[d], ref.A_GG[d])):
            ptr, src = maps.ptr[d][which], maps.src[d][which]
            assert ptr.size == A.nnz + 1 and ptr[0] == 0 and ptr[-1] == src.size
            assert np.all(np.diff(ptr) >= 1)                                   # every stored entry has a contribution
            assert all(np.all(np.diff(src[ptr[p]:ptr[p + 1]]) > 0) for p in range(A.nnz))   # ascending = push order
            v = H.assemble_values_from_map(cells, pts, coeff, ptr, src)
            assert np.array_equal(v, A.data)                                   # the BITS of sparse(I, J, V)
    # every contribution of the loop nest that is routed to a block (:472-497) appears exactly once
    owner = sd.node_owner[cells]
    routed = int(((owner[:, None, :] != 0) & (owner[None, :, :] != 0)
                  & ~((owner[:, None, :] == -1) & (owner[None, :, :] > 0))).sum())
    assert sum(maps.src[d][w].size for d in range(ndom) for w in range(3)) == routed


def test_assembly_maps_of_a_shard_equal_the_full_maps():
    cells, pts, epart, sd = _case(33, 8, 0.2)
    full = H.assembly_maps(cells, epart, sd)
    part = H.assembly_maps(cells, epart, sd, dom_range=(2, 5))
    for d in range(2, 5):
        for w in range(3):
            assert np.array_equal(full.ptr[d][w], part.ptr[d][w]) and np.array_equal(full.src[d][w], part.src[d][w])


def test_host_blocks_agree_with_the_ordered_sums_to_one_ulp():
    """The vectorised host stand-in adds duplicates in scipy's order: same values to the last bit or the one before."""
    cells, pts, epart, sd = _case(33, 4, 0.2)
    coeff = np.exp(M.lognormal_field(pts, 1))
    lb = H.prepare_local_schurs(cells, pts, epart, sd, coeff, -np.ones(pts.shape[1]), np.zeros(pts.shape[1]))
    maps = H.assembly_maps(cells, epart, sd)
    for d in range(4):
        for w, A in enumerate((lb.A_II[d], lb.A_IG[d], lb.A_GG[d])):
            v = H.assemble_values_from_map(cells, pts, coeff, maps.ptr[d][w], maps.src[d][w])
            assert np.abs(v - A.data).max() <= 4e-16 * np.abs(A.data).max()


@pytest.mark.parametrize("nx,ndom", [(33, 4), (65, 16)])
def test_right_hand_sides_of_a_new_coefficient_from_the_boundary_layer(nx, ndom):
    base = PB.build_mesh_and_partition(nx, ndom)
    pr = PB.realization(base, ireal=2, factorize=False)
    up = PB.RhsUpdater(base, pr.f_vals, pr.u_dir)
    assert up.layer.sum() < 0.5 * base[0].shape[1]
    b_I, b_G = up(pr.coeff)
    assert np.array_equal(b_G, pr.blocks.b_G)
    assert all(np.array_equal(a, b) for a, b in zip(b_I, pr.blocks.b_I))


@pytest.mark.parametrize("nx,ndom,jitter", [(17, 4, 0.0), (25, 5, 0.2)])
def test_rhs_maps_reproduce_the_literal_transcription_bit_for_bit(nx, ndom, jitter):
    """Right-hand sides in gather form (kb200_upload_rhs_map): load vector + ordered terms of the boundary-layer entries,
    with a non-constant f and uexact so that every term is distinguishable."""
    cells, pts, epart, sd = _case(nx, ndom, jitter)
    rng = np.random.default_rng(7)
    coeff = np.exp(M.lognormal_field(pts, 3))
    f = -1.0 + 0.1 * rng.standard_normal(pts.shape[1])
    u = 0.734 + 0.1 * rng.standard_normal(pts.shape[1])
    ref = E.prepare_local_schurs_ref(cells, pts, epart, sd, coeff, f, u)
    m = H.rhs_maps(cells, pts, epart, sd, f, u)
    assert np.all(np.diff(m.tgt) > 0) and m.ptr[0] == 0 and m.ptr[-1] == m.src.size and np.all(np.diff(m.ptr) >= 1)
    assert m.tgt.size < 0.7 * (m.bI0.size + m.bG0.size)
    bI, bG = H.assemble_rhs_from_map(cells, pts, coeff, m)
    assert np.array_equal(bI, np.concatenate(ref.b_I)) and np.array_equal(bG, ref.b_G)
    # another coefficient through the same maps
    coeff2 = np.exp(M.lognormal_field(pts, 4))
    ref2 = E.prepare_local_schurs_ref(cells, pts, epart, sd, coeff2, f, u)
    bI2, bG2 = H.assemble_rhs_from_map(cells, pts, coeff2, m)
    assert np.array_equal(bI2, np.concatenate(ref2.b_I)) and np.array_equal(bG2, ref2.b_G)


@pytest.mark.parametrize("seed,ndom,rule", [(0, 5, "min"), (1, 9, "max"), (2, 12, "min")])
def test_assembly_and_rhs_maps_on_jagged_partitions(seed, ndom, rule):
    """Irregular (METIS-like) partitions -- jagged interfaces, nodes shared by three and more subdomains, interface nodes
    on the Dirichlet layer: blocks and right-hand sides in gather form are still the bits of the literal transcription."""
    from test_oracle_maps import _region_growing_partition
    cells, pts, mk, nb = M.structured_mesh(21, jitter=0.15)
    epart, npart = _region_growing_partition(cells, nb, ndom, seed, rule)
    sd = H.set_subdomains(cells, nb, epart, npart, mk == 1)
    rng = np.random.default_rng(seed)
    coeff = np.exp(M.lognormal_field(pts, 5))
    f = -1.0 + 0.1 * rng.standard_normal(pts.shape[1])
    u = 0.734 + 0.1 * rng.standard_normal(pts.shape[1])
    ref = E.prepare_local_schurs_ref(cells, pts, epart, sd, coeff, f, u)
    maps = H.assembly_maps(cells, epart, sd)
    for d in range(ndom):
        for w, A in enumerate((ref.A_II[d], ref.A_IG[d], ref.A_GG[d])):
            assert maps.ptr[d][w].size == A.nnz + 1
            assert np.array_equal(H.assemble_values_from_map(cells, pts, coeff, maps.ptr[d][w], maps.src[d][w]), A.data)
    bI, bG = H.assemble_rhs_from_map(cells, pts, coeff, H.rhs_maps(cells, pts, epart, sd, f, u))
    assert np.array_equal(bI, np.concatenate(ref.b_I)) and np.array_equal(bG, ref.b_G)


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ndom,csize", [(65, 16, 4), (97, 9, 4)])
def test_device_assembly_of_a_new_realization(nx, ndom, csize):
    from krylov_b200 import fem, solvers
    base = PB.build_mesh_and_partition(nx, ndom)
    cells, points, _, _, epart, _, sd = base
    pr1 = PB.realization(base, ireal=1)
    pr2 = PB.realization(base, ireal=2, factorize=False)
    S = fem.B200SchurOperator(sd, pr1.blocks, pr1.factors, csize=csize)
    x = np.random.default_rng(5).standard_normal(sd.n_gamma)
    y1 = S * x
    S.enable_device_assembly(cells, points, epart, pr1.blocks.A_II)
    S.assemble_realization(pr2.coeff)
    # the values the kernel wrote: bit for bit those of its numpy twin (= the literal transcription, see above)
    maps = H.assembly_maps(cells, epart, sd)
    for d in (0, ndom // 2, ndom - 1):
        blocks = (pr2.blocks.A_II[d], pr2.blocks.A_IG[d], pr2.blocks.A_GG[d])
        got = S.h.get_local_block_values(d, *[A.nnz for A in blocks])
        for w, A in enumerate(blocks):
            twin = H.assemble_values_from_map(cells, points, pr2.coeff, maps.ptr[d][w], maps.src[d][w])
            assert np.abs(got[w] - twin).max() <= 4e-16 * np.abs(twin).max()
            assert np.array_equal(got[w], twin)
    # operator, right-hand side and back-substitution of realization 2 against the oracle
    y2 = S * x
    inner2 = E.ExactInterior(pr2.blocks.A_II)
    assert rel(y2, E.apply_local_schurs(pr2.blocks, inner2, sd, x)) <= 1e-12
    st = S.h.stats()
    assert st["host_analyses"] == 1 and st["device_assemblies"] == 1 and st["device_refactorizations"] == 1
    b_I, b_G = PB.RhsUpdater(base, pr2.f_vals, pr2.u_dir)(pr2.coeff)
    bs = S.schur_rhs(b_I, b_G)
    assert rel(bs, E.get_schur_rhs(pr2.blocks, inner2, sd)) <= 1e-12
    # a whole solve of the new realization (no preconditioner): converged for the ORACLE's operator, in about the
    # iterations the oracle's CG takes (a few hundred unpreconditioned iterations: the stopping index has a rounding floor)
    xg, it, _ = solvers.cg(S, bs, np.zeros(sd.n_gamma))
    A2 = lambda v: E.apply_local_schurs(pr2.blocks, inner2, sd, v)   # noqa: E731
    xo, ito, _ = OS.pcg(A2, bs, np.zeros(sd.n_gamma), lambda r: r.copy())
    assert np.linalg.norm(A2(xg) - bs) <= 1.5e-7 * np.linalg.norm(bs)
    assert abs(it - ito) <= max(2, ito // 20)
    # back to realization 1 from its coefficient alone: the operator of the first stream, to rounding
    S.assemble_realization(pr1.coeff)
    assert rel(S * x, y1) <= 1e-12 and S.h.stats()["device_assemblies"] == 2
    # right-hand sides on the device as well: bit for bit the host's, and the same b_schur / back-substitution
    with pytest.raises(Exception):
        S.schur_rhs()                                         # no map yet
    S.enable_device_rhs(cells, points, epart, pr1.f_vals, pr1.u_dir)
    for pr in (pr2, pr1):
        S.assemble_realization(pr.coeff)
        bI, bG = S.h.get_rhs()
        assert np.array_equal(bI, np.concatenate(pr.blocks.b_I)) and np.array_equal(bG, pr.blocks.b_G)
        assert np.array_equal(S.schur_rhs(), S.schur_rhs(pr.blocks.b_I, pr.blocks.b_G))
        ua, ub = S.subdomain_solutions(x), S.subdomain_solutions(x, pr.blocks.b_I)
        assert all(np.array_equal(a, b) for a, b in zip(ua, ub))

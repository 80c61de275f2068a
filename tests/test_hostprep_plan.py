"""Host-side Cholesky and the host analysis (plan) of the stream kernel, checked on the CPU with an
emulator that interprets the device program exactly as csrc/kernels.cuh does."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import ROOT, get_problem
from krylov_b200.hostprep import chol

EMU_SRC = os.path.join(ROOT, "tests", "planemu", "planemu.cpp")
EMU_SO = os.path.join(ROOT, "tests", "planemu", "libplanemu.so")
PLAN_SRC = os.path.join(ROOT, "julia-phd-krylov-spdes_b200", "csrc", "plan.cpp")


@pytest.fixture(scope="module")
def emu():
    from planemu.build import build as build_emu
    lib = C.CDLL(build_emu())
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    lib.planemu_run.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, f64p, i32p, i32p, f64p, i32p, i32p, f64p,
                                C.c_int, C.c_int, f64p, f64p, f64p, f64p, f64p, C.c_char_p, C.c_int]
    return lib


def run_emu(lib, p, L, G, GG, csize, mode, xd, bI, min_block=None, env=None):
    nI, ng = G.shape
    out, uI, st = np.zeros(ng), np.zeros(nI), np.zeros(16)
    err = C.create_string_buffer(256)
    G, GG, L = sp.csc_matrix(G), sp.csc_matrix(GG), sp.csc_matrix(L)
    for m in (G, GG, L):
        m.sort_indices()
    if min_block is not None:
        os.environ["KB200_MIN_BLOCK"] = str(min_block)
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    try:
        rc = lib.planemu_run(nI, ng, p.astype(np.int32), L.indptr.astype(np.int32), L.indices.astype(np.int32),
                             np.ascontiguousarray(L.data), G.indptr.astype(np.int32), G.indices.astype(np.int32),
                             np.ascontiguousarray(G.data), GG.indptr.astype(np.int32), GG.indices.astype(np.int32),
                             np.ascontiguousarray(GG.data), csize, mode, np.ascontiguousarray(xd),
                             np.ascontiguousarray(bI), out, uI, st, err, 256)
    finally:
        os.environ.pop("KB200_MIN_BLOCK", None)
        for k in (env or {}):
            os.environ.pop(k, None)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return out, uI, st


def test_cholesky_factor_and_ordering():
    pr = get_problem(65, 16)
    for d in (0, 5, 15):
        p, L = pr.factors[d]
        A = pr.blocks.A_II[d].tocsc()
        assert sorted(p.tolist()) == list(range(A.shape[0]))
        Ap = A[p][:, p]
        assert abs(L @ L.T - Ap).max() <= 1e-13 * abs(Ap).max()
        assert (sp.triu(L, 1)).nnz == 0
        first = L.indices[L.indptr[:-1]]
        assert np.array_equal(first, np.arange(A.shape[0]))          # diagonal first in each column
    with pytest.raises(np.linalg.LinAlgError):
        chol.cholesky(sp.csc_matrix(np.array([[1.0, 2.0], [2.0, 1.0]])), np.array([0, 1]))


@pytest.mark.parametrize("csize", [1, 2, 4, 8])
@pytest.mark.parametrize("min_block", [0, 4])
def test_plan_program_matches_exact_solves(emu, csize, min_block):
    pr = get_problem(65, 4)
    d = 1
    p, L = pr.factors[d]
    A, G, GG = pr.blocks.A_II[d].tocsc(), pr.blocks.A_IG[d], pr.blocks.A_GG[d]
    lu = spla.splu(A)
    rng = np.random.default_rng(csize)
    xd, bI = rng.standard_normal(G.shape[1]), rng.standard_normal(G.shape[0])
    ref_apply = GG @ xd - G.T @ lu.solve(G @ xd)
    ref_rhs = -(G.T @ lu.solve(bI))
    ref_bs = lu.solve(bI - G @ xd)
    o, _, st = run_emu(emu, p, L, G, GG, csize, 0, xd, bI, min_block)
    assert abs(o - ref_apply).max() <= 1e-12 * abs(ref_apply).max()
    o, _, _ = run_emu(emu, p, L, G, GG, csize, 1, xd, bI, min_block)
    assert abs(o - ref_rhs).max() <= 1e-12 * abs(ref_rhs).max()
    _, u, _ = run_emu(emu, p, L, G, GG, csize, 2, xd, bI, min_block)
    assert abs(u - ref_bs).max() <= 1e-12 * abs(ref_bs).max()
    if min_block:
        assert st[2] < 80          # block inversion collapses the supernode chains' levels


@pytest.mark.parametrize("env", [{}, {"KB200_WARPS": 16}, {"KB200_WARP_PRIVATE": 1},
                                 {"KB200_WARP_PRIVATE": 1, "KB200_PRIVATE_SPLIT": 4},
                                 {"KB200_MERGE_DEPTH": 2}, {"KB200_MERGE_DEPTH": 3, "KB200_WARPS": 16},
                                 {"KB200_MERGE_DEPTH": 4, "KB200_WARP_PRIVATE": 1}])
@pytest.mark.parametrize("csize", [1, 4])
def test_plan_variants_on_a_large_subdomain(emu, env, csize):
    """~4k interior nodes: levels hold more slices than consumer warps (balanced slots with empty
    entries), 16-warp layout, warp-private bottom subtrees -- all against exact solves."""
    pr = get_problem(129, 4)
    d = 2
    p, L = pr.factors[d]
    A, G, GG = pr.blocks.A_II[d].tocsc(), pr.blocks.A_IG[d], pr.blocks.A_GG[d]
    lu = spla.splu(A)
    rng = np.random.default_rng(7)
    xd, bI = rng.standard_normal(G.shape[1]), rng.standard_normal(G.shape[0])
    o, _, st = run_emu(emu, p, L, G, GG, csize, 0, xd, bI, env=env)
    ref = GG @ xd - G.T @ lu.solve(G @ xd)
    assert abs(o - ref).max() <= 1e-12 * abs(ref).max()
    _, u, _ = run_emu(emu, p, L, G, GG, csize, 2, xd, bI, env=env)
    ref = lu.solve(bI - G @ xd)
    assert abs(u - ref).max() <= 1e-12 * abs(ref).max()
    if env.get("KB200_WARP_PRIVATE"):
        assert st[12] > 0       # rows were handled warp-privately


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_plan_on_irregular_factors(emu, seed):
    """Random sparse SPD matrices under random and banded orderings (elimination trees that look nothing
    like a nested-dissection tree: long chains, bushy forests, several roots): every cluster size,
    supernode grouping depth and block threshold against a dense solve."""
    rng = np.random.default_rng(seed)
    n, ng = 260 + 40 * seed, 23
    B = sp.random(n, n, density=0.012, random_state=seed, format="csr")
    A = (B + B.T + sp.diags(np.full(n, 1.0))).tocsr()
    A = (A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel())).tocsc()      # strictly diagonally dominant
    if seed == 2:   # a disconnected matrix: forest with several roots
        A = sp.block_diag([A[:n // 2, :n // 2], A[n // 2:, n // 2:]]).tocsc()
    perm = rng.permutation(n).astype(np.int32) if seed != 1 else np.arange(n, dtype=np.int32)
    L, _ = chol.cholesky(A, perm)
    G = sp.random(n, ng, density=0.05, random_state=seed + 10, format="csc")
    GG = (sp.random(ng, ng, density=0.3, random_state=seed + 20) + sp.eye(ng) * 5.0).tocsc()
    xd, bI = rng.standard_normal(ng), rng.standard_normal(n)
    Ad = A.toarray()
    ref_apply = GG @ xd - G.T @ np.linalg.solve(Ad, G @ xd)
    ref_bs = np.linalg.solve(Ad, bI - G @ xd)
    for csize in (1, 2, 4):
        for md in (1, 2, 3):
            for mb in (0, 2, 4):
                env = {"KB200_MERGE_DEPTH": md}
                o, _, _ = run_emu(emu, perm, L, G, GG, csize, 0, xd, bI, min_block=mb, env=env)
                assert abs(o - ref_apply).max() <= 1e-11 * abs(ref_apply).max(), (csize, md, mb)
                _, u, _ = run_emu(emu, perm, L, G, GG, csize, 2, xd, bI, min_block=mb, env=env)
                assert abs(u - ref_bs).max() <= 1e-11 * abs(ref_bs).max(), (csize, md, mb)


def test_plan_edge_cases(emu):
    # natural ordering (banded factor: one long dependency chain), tiny and 1x1 subdomains
    n = 40
    A = sp.diags([-1.0, 2.5, -1.0], [-1, 0, 1], shape=(n, n)).tocsc()
    p = np.arange(n, dtype=np.int32)
    L, _ = chol.cholesky(A, p)
    G = sp.random(n, 7, density=0.2, random_state=1, format="csc")
    GG = sp.eye(7, format="csc") * 3.0
    rng = np.random.default_rng(0)
    xd, bI = rng.standard_normal(7), rng.standard_normal(n)
    ref = GG @ xd - G.T @ spla.spsolve(A, G @ xd)
    for cs in (1, 4):
        o, _, _ = run_emu(emu, p, L, G, GG, cs, 0, xd, bI)
        assert abs(o - ref).max() <= 1e-12 * abs(ref).max()
    A1 = sp.csc_matrix(np.array([[4.0]]))
    L1, _ = chol.cholesky(A1, np.array([0]))
    G1 = sp.csc_matrix(np.array([[1.0, 0.0, 2.0]]))
    GG1 = sp.eye(3, format="csc")
    x3 = np.array([1.0, 2.0, 3.0])
    o, _, _ = run_emu(emu, np.array([0]), L1, G1, GG1, 2, 0, x3, np.zeros(1))
    assert np.allclose(o, GG1 @ x3 - G1.T @ (G1 @ x3 / 4.0))
    # malformed factor: missing diagonal
    Lbad = sp.csc_matrix(np.array([[0.0, 0.0], [1.0, 1.0]]))
    with pytest.raises(RuntimeError):
        run_emu(emu, np.array([0, 1]), Lbad, sp.csc_matrix((2, 1)), sp.csc_matrix((1, 1)), 1, 0, np.zeros(1), np.zeros(2))

"""Fuzzing of the host analysis (csrc/plan.cpp) through the CPU emulator of the device program (tests/planemu): random
sparse / grid / banded / block-diagonal SPD matrices under random, natural and minimum-degree orderings, random cluster
sizes, supernode grouping depths, block thresholds, warp layouts; the three modes of the stream kernel against exact
solves.  No GPU needed.   usage: fuzz_plan.py [seconds=600] [first_seed=1000]"""
import sys, os, time, ctypes as C, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import krylov_b200
from krylov_b200.hostprep import chol
import test_hostprep_plan as T
from planemu.build import build as build_emu
lib = C.CDLL(build_emu())
i32p = np.ctypeslib.ndpointer(np.int32, flags="C"); f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
lib.planemu_run.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, f64p, i32p, i32p, f64p, i32p, i32p, f64p, C.c_int, C.c_int, f64p, f64p, f64p, f64p, f64p, C.c_char_p, C.c_int]
t_end = time.time() + float(sys.argv[1]) if len(sys.argv) > 1 else time.time() + 600
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ncase = nrun = 0; worst = 0.0
while time.time() < t_end:
    rng = np.random.default_rng(seed)
    n = int(rng.integers(30, 1200)); ng = int(rng.integers(1, 60))
    kind = int(rng.integers(0, 4))
    if kind == 0:      # random sparse
        B = sp.random(n, n, density=float(rng.uniform(0.002, 0.03)), random_state=seed, format="csr")
        A = (B + B.T).tocsr()
    elif kind == 1:    # 2D grid Laplacian-like
        m = int(np.sqrt(n)); n = m * m
        I = sp.eye(m); D = sp.diags([-1.0, -1.0], [-1, 1], shape=(m, m))
        A = (sp.kron(I, D) + sp.kron(D, I)).tocsr()
    elif kind == 2:    # banded
        bw = int(rng.integers(1, 12)); A = sp.diags([-1.0] * (2 * bw), [k for k in range(-bw, bw + 1) if k], shape=(n, n)).tocsr()
    else:              # block diagonal forest of random blocks
        blocks = []; left = n
        while left > 0:
            b = int(min(left, rng.integers(1, 200))); Bk = sp.random(b, b, density=0.05, random_state=seed + left, format="csr"); blocks.append(Bk + Bk.T); left -= b
        A = sp.block_diag(blocks).tocsr()
    A = (A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)).tocsc()
    pk = int(rng.integers(0, 3))
    if pk == 0: perm = rng.permutation(n).astype(np.int32)
    elif pk == 1: perm = np.arange(n, dtype=np.int32)
    else:
        lu = spla.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        perm = np.argsort(np.asarray(lu.perm_c)).astype(np.int32)
    L, _ = chol.cholesky(A, perm)
    G = sp.random(n, ng, density=float(rng.uniform(0.01, 0.3)), random_state=seed + 10, format="csc")
    GG = (sp.random(ng, ng, density=0.3, random_state=seed + 20) + sp.eye(ng) * 5.0).tocsc()
    xd, bI = rng.standard_normal(ng), rng.standard_normal(n)
    lus = spla.splu(A)
    ref_apply = GG @ xd - G.T @ lus.solve(G @ xd); ref_bs = lus.solve(bI - G @ xd); ref_rhs = -(G.T @ lus.solve(bI))
    for _ in range(4):
        csize = int(rng.choice([1, 2, 4, 8])); md = int(rng.choice([1, 2, 3, 4])); mb = int(rng.choice([0, 2, 4, 8]))
        env = {"KB200_MERGE_DEPTH": md}
        if rng.random() < 0.3: env["KB200_WARPS"] = 16
        if rng.random() < 0.2: env["KB200_WARP_PRIVATE"] = 1
        try:
            o, _, _ = T.run_emu(lib, perm, L, G, GG, csize, 0, xd, bI, min_block=mb, env=env)
            e1 = abs(o - ref_apply).max() / max(abs(ref_apply).max(), 1e-300)
            _, u, _ = T.run_emu(lib, perm, L, G, GG, csize, 2, xd, bI, min_block=mb, env=env)
            e2 = abs(u - ref_bs).max() / max(abs(ref_bs).max(), 1e-300)
            o, _, _ = T.run_emu(lib, perm, L, G, GG, csize, 1, xd, bI, min_block=mb, env=env)
            e3 = abs(o - ref_rhs).max() / max(abs(ref_rhs).max(), 1e-300)
        except Exception as ex:
            print("EXC", seed, n, ng, kind, pk, csize, md, mb, env, str(ex)[:200], flush=True); continue
        nrun += 1; worst = max(worst, e1, e2, e3)
        if max(e1, e2, e3) > 1e-10:
            print("FAIL", seed, n, ng, kind, pk, csize, md, mb, env, e1, e2, e3, flush=True)
    ncase += 1; seed += 1
print("done: matrices", ncase, "plan runs", nrun, "worst rel err", worst, "last seed", seed, flush=True)

"""Probe: the stream kernel on factors from a fill-reducing ordering the library did not choose.  A Julia user uploads
CHOLMOD's (`F.p`, `sparse(F.L)`), i.e. an AMD-type ordering; here: SuperLU's minimum-degree ordering on A' + A
(scipy `permc_spec="MMD_AT_PLUS_A"`) next to the repo's own geometric nested dissection, same subdomains.
usage: probe_ordering.py [nx ndom]  -> JSON lines (levels, stream size, kernel time, GB/s, parity of the apply)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import krylov_b200  # noqa: F401
from krylov_b200 import fem
from krylov_b200.hostprep import chol, problem as PB

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 64
pr = PB.build_problem(nx, ndom, ireal=1)
B, parts = PB.algorithmic_bytes_per_iteration(pr)


def mmd_factor(d):
    A = sp.csc_matrix(pr.blocks.A_II[d])
    lu = spla.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    p = np.asarray(lu.perm_c, dtype=np.int32)          # column permutation of the ordering step
    perm = np.argsort(p).astype(np.int32)              # perm[new] = old
    L, _ = chol.cholesky(A, perm)
    return perm, L


with ThreadPoolExecutor(min(32, os.cpu_count() or 1)) as ex:
    t0 = time.time()
    mmd = list(ex.map(mmd_factor, range(ndom)))
    t_mmd = time.time() - t0
x = np.random.default_rng(0).standard_normal(pr.sd.n_gamma)
ref = None
for name, factors in (("nested dissection (repo)", pr.factors), ("minimum degree (SuperLU MMD on A'+A)", mmd)):
    nnzL = sum(f[1].nnz for f in factors)
    alg = sum(2 * (12 * factors[d][1].nnz + 4 * (pr.blocks.A_II[d].shape[0] + 1) + 24 * pr.blocks.A_II[d].shape[0]) for d in range(ndom)) + parts["spmv"]
    t0 = time.time()
    S = fem.B200SchurOperator(pr.sd, pr.blocks, factors, profile=True)
    S.h.finalize()
    tf = time.time() - t0
    st = S.h.stats()
    ms = S.h.time_kernel(0, 20)
    y = S * x
    if ref is None:
        ref = y
    print(json.dumps({"ordering": name, "nx": nx, "ndom": ndom, "nnz_L": int(nnzL), "levels_per_sweep": int(st["max_levels"]),
                      "top_nodes": int(st["max_top"]), "stream_GB": st["stream_bytes"] / 1e9, "smem_bytes": int(st["smem_bytes"]),
                      "consumer_warps": int(st["consumer_warps"]), "merge_depth": int(st["merge_depth"]), "host_analysis_s": tf,
                      "kernel_ms": ms, "stream_GBs": st["stream_bytes"] / ms / 1e6, "algorithmic_GBs": alg / ms / 1e6,
                      "apply_rel_diff_vs_first": float(np.abs(y - ref).max() / np.abs(ref).max())}), flush=True)
    del S

"""Profiling target (run under ncu on the GPU box): a short stretch of the benchmark's PCG iterations.
usage: [OPTS="shard_world=8+shard_rank=3"] ncu_target.py [nx ndom iters mode]
mode: pcg (default; kernels per iteration: stream, fused_xr, dense_gemv, fused_p) | dmma (the DMMA kernels:
batched dense products with 64 right-hand sides, tall-skinny V'AV and V (Q Z) at spdim = 3 ndom, and the multi-vector
stream kernel).  The preconditioner blocks are identity-like: the cost of the kernels does not depend on their values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch  # noqa: F401
import krylov_b200  # noqa: F401
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
mode = sys.argv[4] if len(sys.argv) > 4 else "pcg"
opts = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("OPTS", "").split("+") if kv)}
base = PB.build_mesh_and_partition(nx, ndom)
sd = base[-1]
d0, d1 = 0, sd.ndom
if "shard_world" in opts:
    d0, d1 = (sd.ndom * opts["shard_rank"]) // opts["shard_world"], (sd.ndom * (opts["shard_rank"] + 1)) // opts["shard_world"]
pr = PB.realization(base, ireal=1, dom_range=(d0, d1))
S = fem.B200SchurOperator(sd, pr.blocks, pr.factors, options=opts)
M = fem.B200NeumannNeumann(S, {d: np.eye(sd.gamma_l2g[d].size) * 0.5 for d in range(d0, d1)})
n = sd.n_gamma
rng = np.random.default_rng(0)
if mode == "pcg":
    b = rng.standard_normal(n)
    x, it, res = solvers.pcg(S, b, np.zeros(n), M, maxit=iters)
    print("pcg it", it, "stats", {k: S.h.stats()[k] for k in ("n_cta", "consumer_warps", "stream_bytes", "vec_blocks")})
else:
    X = rng.standard_normal((n, 64))
    Z = S.h.apply_nn(X)                                   # dense_gemm_kernel<64>
    Y = S.h.apply_schur(X[:, :4].copy())                  # multi-vector stream kernel
    h = S.h
    spdim = min(3 * ndom, 768)
    V, AV = h.mat_alloc(spdim), h.mat_alloc(spdim)
    h.mat_set(V, 0, rng.standard_normal((n, spdim)))
    h.mat_set(AV, 0, rng.standard_normal((n, spdim)))
    C = h.op_gemm_tn(V, 0, spdim, AV, 0, spdim)           # gemm_tn_dmma_kernel
    h.op_gemm_nn(V, 0, spdim, rng.standard_normal((spdim, int(2.5 * ndom))), V, 0)   # gemm_nn_dmma_kernel
    print("dmma target done", Z.shape, Y.shape, C.shape)

"""Development probe (GPU): the multi-vector stream kernel and the DMMA dense products at a given size.
usage: probe_multi.py [nx ndom]; prints one JSON line per measurement."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_b200  # noqa: F401
from krylov_b200 import fem
from krylov_b200.hostprep import problem as PB

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 256
opts = dict(kv.split("=") for kv in os.environ.get("OPTS", "").split("+") if kv)
t = time.time()
pr = PB.build_problem(nx, ndom, ireal=1)
print("host prep %.1fs" % (time.time() - t), flush=True)
S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors, options={k: int(v) for k, v in opts.items()})
S.h.finalize()
st = S.h.stats()
print(json.dumps({k: st[k] for k in ("stream_bytes", "smem_bytes", "consumer_warps", "ring_stages", "merge_depth", "max_levels", "n_cta")}), flush=True)
t1 = S.h.time_kernel(0, 20)
out = {"nx": nx, "ndom": ndom, "opts": opts, "stream_R1_ms": t1}
for which, R in ((4, 2), (5, 4)):
    try:
        ms = S.h.time_kernel(which, 20)
        out[f"stream_R{R}_ms"] = ms
        out[f"stream_R{R}_ms_per_rhs"] = ms / R
    except Exception as e:
        out[f"stream_R{R}"] = str(e)[:120]
print(json.dumps(out), flush=True)
t = time.time()
Sd = S.assemble_local_schurs()
t_asm = time.time() - t
Pi = {d: PB.pinv_sym(Sd[d]) for d in range(ndom)}
M = fem.B200NeumannNeumann(S, Pi)
o2 = {"assemble_local_schurs_s": t_asm, "nn_gemv_ms": S.h.time_kernel(1, 20), "dense_bytes": st["dense_bytes"]}
for which, k in ((6, 8), (7, 16), (8, 32), (9, 64)):
    ms = S.h.time_kernel(which, 10)
    o2[f"nn_dmma_k{k}_ms"] = ms
    o2[f"nn_dmma_k{k}_tflops"] = 2 * st["dense_bytes"] / 8 * k / ms / 1e9
    o2[f"nn_dmma_k{k}_GBs"] = st["dense_bytes"] / ms / 1e6
print(json.dumps(o2), flush=True)
# tall-skinny products at the restart sizes of eigpcg (spdim = 3 ndom, nvec = 1.25 ndom)
h, n = S.h, pr.sd.n_gamma
spdim, nvec = min(3 * ndom, 768), int(1.25 * ndom)
rng = np.random.default_rng(0)
V = h.mat_alloc(spdim); AV = h.mat_alloc(spdim)
h.mat_set(V, 0, rng.standard_normal((n, spdim)))
h.mat_set(AV, 0, rng.standard_normal((n, spdim)))
import torch
torch.cuda.synchronize()
for rep in range(2):
    t = time.perf_counter(); C = h.op_gemm_tn(V, 0, spdim, AV, 0, spdim); t_tn = time.perf_counter() - t
Q = rng.standard_normal((spdim, 2 * nvec))
for rep in range(2):
    t = time.perf_counter(); h.op_gemm_nn(V, 0, spdim, Q, V, 0); t_nn = time.perf_counter() - t
t = time.perf_counter(); h.op_apply(0, V, 0, AV, 0, 16); t_ap = time.perf_counter() - t
print(json.dumps({"n_gamma": n, "spdim": spdim, "two_nvec": 2 * nvec, "gemm_tn_wall_ms": t_tn * 1e3,
                  "gemm_tn_tflops": 2 * n * spdim * spdim / t_tn / 1e12, "gemm_nn_wall_ms": t_nn * 1e3,
                  "gemm_nn_tflops": 2 * n * spdim * 2 * nvec / t_nn / 1e12, "op_apply_16cols_wall_ms": t_ap * 1e3}), flush=True)

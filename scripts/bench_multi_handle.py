"""One process, several GPUs (kb200_create_multi): the benchmark solve of bench.py driven through a single handle.
usage: bench_multi_handle.py [nx ndom ngpus]   -> one JSON line"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import krylov_b200  # noqa: F401
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ngpus = int(sys.argv[3]) if len(sys.argv) > 3 else torch.cuda.device_count()
t0 = time.time()
base = PB.build_mesh_and_partition(nx, ndom)
sd = base[-1]
pr = PB.realization(base, ireal=1)
pr0 = PB.realization(base, ireal=0, median=True)
print(f"host preparation {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
devices = list(range(ngpus))
t0 = time.time()
S0 = fem.B200SchurOperator(sd, pr0.blocks, pr0.factors, devices=devices)
Sd0 = S0.assemble_local_schurs()
PiS = {d: PB.pinv_sym(Sd0[d]) for d in range(ndom)}
del S0, Sd0
S = fem.B200SchurOperator(sd, pr.blocks, pr.factors, devices=devices)
M = fem.B200NeumannNeumann(S, PiS)
b = S.schur_rhs(pr.blocks.b_I, pr.blocks.b_G)
print(f"device setup {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
n = sd.n_gamma
out = []
for rep in range(4):
    x = np.zeros(n)
    t0 = time.perf_counter()
    _, it, res, tm = solvers.pcg(S, b, x, M, maxit=5000, return_timings=True)
    wall = time.perf_counter() - t0
    out.append((tm.solve_ms, wall, it))
dev_ms = min(o[0] for o in out[1:])
wall = min(o[1] for o in out[1:])
print(json.dumps({"metric": "schur_pcg_solves_per_sec", "handle": "kb200_create_multi (one process)", "n_gpus": ngpus, "nx": nx, "ndom": ndom,
                  "value": 1e3 / dev_ms, "e2e_value": 1.0 / wall, "it": int(out[-1][2]), "ms_per_iteration": dev_ms / max(out[-1][2] - 1, 1),
                  "relres": float(res[-1] / np.linalg.norm(b)), "peer_exchange": S.h.stats()["peer_exchange"]}), flush=True)

"""Development probe: build the per-GPU shard of the benchmark workload (as rank r of w) and launch the
stream kernel a few times (for ncu captures of the multi-GPU per-rank configuration on one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krylov_b200  # noqa: F401
from krylov_b200 import fem
from krylov_b200.hostprep import problem as PB

nx, ndom, w, r = (int(v) for v in (sys.argv[1:5] + ["2001", "256", "8", "3"][len(sys.argv) - 1:]))
base = PB.build_mesh_and_partition(nx, ndom)
sd = base[-1]
d0, d1 = (sd.ndom * r) // w, (sd.ndom * (r + 1)) // w
pr = PB.realization(base, ireal=1, dom_range=(d0, d1))
opts = {"shard_world": w, "shard_rank": r}
opts.update({k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("OPTS", "").split("+") if kv)})
S = fem.B200SchurOperator(sd, pr.blocks, pr.factors, options=opts)
S.h.finalize()
print(S.h.stats(), S.h.time_kernel(0, 10), flush=True)

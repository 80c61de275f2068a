"""Development probe: time the PCG solve of the benchmark workload under several option sets inside ONE
job (the host preparation dominates the cost of a job).  Works under torchrun like bench.py.
usage: VARIANTS="pdl=0,pdl=1,pdl=1+fuse_pcg=0" [torchrun ...] python scripts/variants.py [nx ndom]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import krylov_b200  # noqa: F401
from krylov_b200 import capi, fem, solvers
from krylov_b200.hostprep import problem as PB

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(lr)
nccl_id = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    obj = [capi.Handle.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    nccl_id = obj[0]
base = PB.build_mesh_and_partition(nx, ndom)
sd = base[-1]
d0, d1 = (sd.ndom * rank) // world, (sd.ndom * (rank + 1)) // world
pr = PB.realization(base, ireal=1, dom_range=(d0, d1))
if world > 1:
    pr.blocks.b_G = PB.interface_rhs(base, pr.coeff, pr.f_vals, pr.u_dir)
# OPTS="shard_world=8+shard_rank=3": emulate the per-GPU shard of a multi-GPU run on one GPU
opts = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("OPTS", "").split("+") if kv)}
if "shard_world" in opts:
    d0, d1 = (sd.ndom * opts["shard_rank"]) // opts["shard_world"], (sd.ndom * (opts["shard_rank"] + 1)) // opts["shard_world"]
    pr = PB.realization(base, ireal=1, dom_range=(d0, d1))
S = fem.B200SchurOperator(sd, pr.blocks, pr.factors, device=lr, rank=rank, world=world, nccl_id=nccl_id, profile=True,
                          options=opts)
# identity-like preconditioner blocks are enough to time the kernels: the cost of M does not depend on its values
rng = np.random.default_rng(0)
PiS = {d: np.eye(sd.gamma_l2g[d].size) * 0.5 for d in range(d0, d1)}
M = fem.B200NeumannNeumann(S, PiS)
b = rng.standard_normal(sd.n_gamma)
maxit = int(os.environ.get("MAXIT", "400"))
for var in os.environ.get("VARIANTS", "pdl=1").split(","):
    for kv in var.split("+"):
        k, v = kv.split("=")
        S.h.set_option(k, int(v))
    ts = []
    for rep in range(3):
        x = np.zeros(sd.n_gamma)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        _, it, res, tm = solvers.pcg(S, b, x, M, maxit=maxit, return_timings=True)
        ts.append(tm.solve_ms / max(it - 1, 1))
    t = torch.tensor([min(ts)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{var:40s} it {it} ms/iteration {float(t[0]):.4f}  (schur {tm.schur_ms / max(tm.schur_launches, 1):.4f}, nn {tm.nn_ms / max(tm.nn_launches, 1):.4f}) res {res[-1] / res[0]:.2e}", flush=True)
if os.environ.get("XCH"):   # the fused assembly (+ exchange) / update kernel alone, back to back
    ms = S.h.time_kernel(3, 200)
    if rank == 0:
        print(f"fused assembly{' + peer exchange' if S.h.stats()['peer_exchange'] else ''} + update kernel alone: {ms * 1e3:.1f} us per launch", flush=True)
if world > 1:
    dist.destroy_process_group()

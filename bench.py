#!/usr/bin/env python
"""Benchmark of the Schur-complement PCG hot path (BASELINE.json metric).

A step = one whole Neumann-Neumann PCG solve of S u_Γ = b_schur (Example07's inner loop body,
Example07_PcgSchurStochasticEllipticPde.jl:273) on the synthetic 4M-DOF / 256-subdomain workload of
SURVEY.md 8(d).  `value` is solves/s from the device-side solve time (inputs resident in HBM);
`e2e` is the same through kb200_pcg with host buffers (H2D of b and x0, D2H of x / res_norm inside
the timed region).  With --impl reference the same metric is measured for the CPU port of the
reference algorithm (oracle/kb_oracle.c, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def schur_kernel_algorithmic_bytes(pr, d0, d1):
    """Algorithmic bytes of ONE launch of the local-Schur stream kernel:
    forward + backward triangular solves, the two A_IΓ products, A_ΓΓ,d x_d and the gather maps."""
    b = 0.0
    for d in range(d0, d1):
        nI = pr.blocks.A_II[d].shape[0]
        ng = pr.blocks.A_GG[d].shape[0]
        b += 2 * (12 * pr.factors[d][1].nnz + 4 * (nI + 1) + 8 * 3 * nI)
        b += 2 * (12 * pr.blocks.A_IG[d].nnz + 4 * (nI + 1)) + 12 * pr.blocks.A_GG[d].nnz + 4 * (ng + 1)
        b += 4 * 2 * ng + 8 * 2 * ng
    return b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=2001, help="nodes per side (2001 -> 4.0M DOF)")
    ap.add_argument("--ndom", type=int, default=256)
    ap.add_argument("--csize", type=int, default=0, help="CTAs per subdomain cluster (0 = automatic)")
    ap.add_argument("--maxit", type=int, default=5000)
    ap.add_argument("--cpu-sample-iters", type=int, default=6)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: WORLD_SIZE={world} != --gpus {args.gpus}")
    workload = f"synthetic {args.nx}x{args.nx}-node unit-square P1 mesh, {args.ndom} RCB subdomains, lognormal a (sig2=1, L=0.1, realization 1), median NN preconditioner"
    config = {"workload": workload, "n_dof": (args.nx - 2) ** 2, "ndom": args.ndom, "tol": 1e-7,
              "x0": "zeros", "l2": "inputs_larger_than_L2 (3.5 GB streamed per Schur apply at N=1; >= 0.44 GB per GPU at N=8)",
              "parallelism": f"subdomains sharded over {max(world, 1)} GPU(s), NCCL all-reduce of the interface vector"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, config)

    import torch
    import torch.distributed as dist
    import krylov_b200  # noqa: F401
    from krylov_b200 import capi, fem, solvers
    from krylov_b200.hostprep import problem as PB

    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        obj = [capi.Handle.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]

    t0 = time.time()
    base = PB.build_mesh_and_partition(args.nx, args.ndom)
    sd = base[-1]
    d0, d1 = (sd.ndom * rank) // world, (sd.ndom * (rank + 1)) // world
    pr = PB.realization(base, ireal=1, dom_range=(d0, d1))
    pr0 = PB.realization(base, ireal=0, median=True, dom_range=(d0, d1))       # "median" operator (Example07:89)
    if world > 1:   # every rank assembled only its own elements: complete b_Γ by a sum over ranks
        bg = torch.from_numpy(pr.blocks.b_G).cuda()
        dist.all_reduce(bg)
        pr.blocks.b_G = bg.cpu().numpy()
    log(f"[rank {rank}] host preparation {time.time() - t0:.1f}s (n_gamma={sd.n_gamma}, local subdomains {d0}:{d1})")

    t0 = time.time()
    # constant preconditioner from the median realization (Example07:152-154): S_d assembled on the
    # device, pseudo-inverse on the host (LAPACK), uploaded once
    S0 = fem.B200SchurOperator(sd, pr0.blocks, pr0.factors, device=local_rank, csize=args.csize,
                               options={"shard_world": world, "shard_rank": rank})
    Sd0 = S0.assemble_local_schurs()
    PiS = {d: PB.pinv_sym(Sd0[d]) for d in range(d0, d1)}
    del S0, Sd0
    S = fem.B200SchurOperator(sd, pr.blocks, pr.factors, device=local_rank, csize=args.csize, rank=rank,
                              world=world, nccl_id=nccl_id, profile=True)
    M = fem.B200NeumannNeumann(S, PiS)
    b_I_cat = np.concatenate([pr.blocks.b_I[d] if d0 <= d < d1 else np.zeros(sd.node_Id[d].size)
                              for d in range(sd.ndom)])
    b_schur = S.h.schur_rhs(b_I_cat, pr.blocks.b_G)
    stats = S.h.stats()
    if world > 1:
        config["parallelism"] = (f"subdomains sharded over {world} GPUs; interface vectors combined by "
                                 + ("the fused assemble + peer-memory exchange kernel (CUDA IPC over NVLink, no NCCL call per iteration)"
                                    if stats["peer_exchange"] else "an NCCL all-reduce per operator application"))
    else:
        config["parallelism"] = "1 GPU, all subdomains local"
    config["kernel"] = {"cluster_size": int(stats["csize"]), "consumer_warps": int(stats["consumer_warps"]),
                        "ring_stages": int(stats["ring_stages"]), "ctas": int(stats["n_cta"]),
                        "supernode_merge_depth": int(stats["merge_depth"]), "min_block": int(stats["min_block"]),
                        "levels_per_sweep": int(stats["max_levels"])}
    log(f"[rank {rank}] device setup {time.time() - t0:.1f}s; stream {stats['stream_bytes'] / 1e9:.2f} GB, "
        f"levels {stats['max_levels']:.0f}, smem/CTA {stats['smem_bytes']:.0f} B, CTAs {stats['n_cta']:.0f}")

    n = sd.n_gamma
    x = np.zeros(n)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    its = []
    for _ in range(args.warmup):
        x[:] = 0.0
        _, it, res, tm = solvers.pcg(S, b_schur, x, M, maxit=args.maxit, return_timings=True)
        its.append(it)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms = schur_ms = nn_ms = 0.0
    schur_launches = nn_launches = launches = h2d = d2h = 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        x[:] = 0.0
        _, it, res, tm = solvers.pcg(S, b_schur, x, M, maxit=args.maxit, return_timings=True)
        dev_ms += tm.solve_ms
        schur_ms += tm.schur_ms
        nn_ms += tm.nn_ms
        schur_launches += tm.schur_launches
        nn_launches += tm.nn_launches
        launches += tm.kernel_launches
        h2d += tm.h2d_bytes
        d2h += tm.d2h_bytes
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    per_rank = None
    if world > 1:
        t = torch.tensor([dev_ms, wall_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_s = float(t[0]), float(t[1])
        # per-rank kernel times (the ranks meet at every interface exchange: the slowest one sets the pace)
        mine = torch.tensor([schur_ms / max(schur_launches, 1), nn_ms / max(nn_launches, 1)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"schur_kernel_ms": [round(float(a[0]), 4) for a in allr], "nn_kernel_ms": [round(float(a[1]), 4) for a in allr]}
    iters = it - 1
    relres = float(res[-1] / np.linalg.norm(b_schur))
    peak, peak_src = measured_peak()
    alg = schur_kernel_algorithmic_bytes(pr, d0, d1)
    avg_ms = schur_ms / max(schur_launches, 1)
    achieved = alg / (avg_ms * 1e-3) / 1e9
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("nx") == args.nx and tj.get("ndom") == args.ndom and world == 1:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = tj.get("source")
    B_iter, parts = PB.algorithmic_bytes_per_iteration(pr, (d0, d1))
    out = {
        "metric": "schur_pcg_solves_per_sec", "value": args.steps / (dev_ms * 1e-3), "unit": "solves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "iterations_per_solve": iters, "ms_per_iteration": dev_ms / args.steps / max(iters, 1), "relres": relres,
        "e2e": {"value": args.steps / wall_s, "unit": "solves/s", "h2d_bytes_per_step": h2d // args.steps,
                "d2h_bytes_per_step": d2h // args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "schur_stream_kernel<MODE_APPLY>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "stream_bytes_per_launch": float(stats["stream_bytes"]), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_ms,
                     # launch durations are sampled (every 16th iteration is bracketed by events; an event between
                     # two kernels would serialise the programmatic dependent launch chain of the others)
                     "timed_launches": int(schur_launches),
                     "kernel_share_of_step": avg_ms * (iters + 1) * args.steps / dev_ms,
                     "per_iteration": {"B_iter_bytes": B_iter, "achieved_GBs": B_iter * iters * args.steps / (dev_ms * 1e-3) / 1e9,
                                       "frac": B_iter * iters * args.steps / (dev_ms * 1e-3) / 1e9 / peak}},
        "clocks": clocks,
    }
    if per_rank is not None:
        out["per_rank"] = per_rank
    if world == 1 and rank == 0:
        out["cpu_baseline"], out["parity_at_full_size"] = cpu_baseline(args, pr, sd, PiS, b_schur, iters, S=S, M=M)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args, pr, sd, PiS, b_schur, iters, threads=None, S=None, M=None):
    """The oracle's C port of the same algorithm (exact Cholesky interior solves) timed on this
    box's host cores on a bounded sample: `cpu_sample_iters` PCG iterations, extrapolated to the
    iteration count of the full solve."""
    from oracle.cpu_model import CpuModel
    cores = threads or (os.cpu_count() or 1)
    cm = CpuModel(sd, pr.blocks, pr.factors, [PiS[d] for d in range(sd.ndom)])
    k = args.cpu_sample_iters
    t0 = time.perf_counter()
    cm.pcg(b_schur, np.zeros(sd.n_gamma), maxit=k + 1, nthreads=cores)
    t_all = (time.perf_counter() - t0) / (k + 1)      # k iterations + the initial residual applies
    t0 = time.perf_counter()
    cm.pcg(b_schur, np.zeros(sd.n_gamma), maxit=3, nthreads=1)
    t_one = (time.perf_counter() - t0) / 3
    parity = None
    if S is not None:   # outside the timed region: the CUDA operators against the CPU port at the benchmark's own size
        rng = np.random.default_rng(0)
        v = rng.standard_normal(sd.n_gamma)
        ys, yc = S * v, cm.apply_schur(v, nthreads=cores)
        zs, zc = M.solve(v), cm.apply_nn(v, nthreads=cores)
        parity = {"apply_schur_max_rel_err": float(np.abs(ys - yc).max() / np.abs(yc).max()),
                  "apply_nn_max_rel_err": float(np.abs(zs - zc).max() / np.abs(zc).max())}
    return {"value": 1.0 / (t_all * (iters + 1)), "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{k} PCG iterations (+ initial residual) of the same workload with {cores} OpenMP threads over "
                      f"subdomains, {t_all * 1e3:.1f} ms/iteration, extrapolated to {iters} iterations",
            "single_core": {"value": 1.0 / (t_one * (iters + 1)), "ms_per_iteration": t_one * 1e3,
                            "sample": "2 PCG iterations, sequential for-idom loop as in the reference"},
            "ms_per_iteration": t_all * 1e3}, parity


def run_reference(args, config):
    """CPU arm: the reference algorithm's C port (oracle/kb_oracle.c) with all host threads."""
    import krylov_b200  # noqa: F401
    from krylov_b200.hostprep import problem as PB
    import oracle.epdd as E  # noqa: F401
    from oracle.cpu_model import CpuModel
    cores = os.cpu_count() or 1
    config = dict(config, parallelism=f"CPU port of the reference algorithm, {cores} OpenMP threads over subdomains")
    base = PB.build_mesh_and_partition(args.nx, args.ndom)
    sd = base[-1]
    pr = PB.realization(base, ireal=1)
    # median preconditioner on the CPU: dense S_d by exact solves would take minutes at 4M DOF, so the
    # sample uses the realization's own dense blocks for a handful of subdomains' worth of time:
    # the per-iteration cost of M is identical whichever S_d the pseudo-inverse came from.
    PiS = [np.eye(g.size) for g in sd.gamma_l2g]
    cm = CpuModel(sd, pr.blocks, pr.factors, PiS)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(sd.n_gamma)
    k = args.cpu_sample_iters
    iters_full = int(os.environ.get("KB200_REF_ITERS", "0")) or None

    def step():
        t0 = time.perf_counter()
        cm.pcg(b, np.zeros(sd.n_gamma), maxit=k + 1, nthreads=cores)
        return (time.perf_counter() - t0) / (k + 1)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    per_it = [step() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    t_it = float(np.mean(per_it))
    # iterations of a full solve: taken from the last B200 bench line if the driver exported it,
    # else from the documented count for this workload
    iters = iters_full or int(json.load(open(os.path.join(ROOT, "profiles", "workload_iters.json"))).get(
        f"{args.nx}x{args.ndom}", 300)) if os.path.exists(os.path.join(ROOT, "profiles", "workload_iters.json")) else (iters_full or 300)
    v = 1.0 / (t_it * (iters + 1))
    out = {"impl": "reference", "metric": "schur_pcg_solves_per_sec", "value": v, "unit": "solves/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                            "sample": f"each step = {k} PCG iterations (+ initial residual) with {cores} OpenMP threads, "
                                      f"{t_it * 1e3:.1f} ms/iteration, extrapolated to {iters} iterations per solve; "
                                      f"wall {wall:.1f}s"},
           "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "ms_per_iteration": t_it * 1e3, "iterations_per_solve": iters}
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())

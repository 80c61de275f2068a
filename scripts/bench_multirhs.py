"""BASELINE.json config 4: Example11-style multiple-right-hand-side solve with a constant operator
(Example11_EigInitPcgEllipticPdeMultipleRhs_Functions.jl:140-179): ~1M DOF, 128 subdomains, 64 right-hand sides
b = rand(n) (seed 481456, :145), median Neumann-Neumann preconditioner.  One JSON line per variant:
  sequential    64 x kb200_pcg, one after the other (the reference's loop)
  block_mf      kb200_pcg_multi, matrix-free operator (multi-vector stream kernel) + DMMA preconditioner products
  block_dense   kb200_pcg_multi, assembled operator: S_d X_d and Pi S_d (D R)_d as batched FP64 DMMA GEMMs
usage: bench_multirhs.py [nx ndom nrhs]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import krylov_b200  # noqa: F401
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nrhs = int(sys.argv[3]) if len(sys.argv) > 3 else 64
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
base = PB.build_mesh_and_partition(nx, ndom)
sd = base[-1]
pr0 = PB.realization(base, ireal=0, median=True)
S0 = fem.B200SchurOperator(sd, pr0.blocks, pr0.factors)
Sd0 = S0.assemble_local_schurs()
PiS = {d: PB.pinv_sym(Sd0[d]) for d in range(ndom)}
del S0, Sd0
pr = PB.realization(base, ireal=1)
S = fem.B200SchurOperator(sd, pr.blocks, pr.factors, profile=True)
M = fem.B200NeumannNeumann(S, PiS)
n = sd.n_gamma
B = np.random.default_rng(481456).random((n, nrhs))
st = S.h.stats()
t0 = time.perf_counter(); S.assemble_local_schurs(); t_asm = time.perf_counter() - t0
common = {"metric": "rhs_solves_per_sec", "unit": "rhs/s", "workload": f"{nx}x{nx} nodes, {ndom} subdomains, {nrhs} right-hand sides rand(n) seed 481456, "
          "median NN preconditioner, realization 1", "n_dof": (nx - 2) ** 2, "n_gamma": n, "dtype": "f64",
          "stream_bytes": st["stream_bytes"], "dense_bytes": st["dense_bytes"], "hbm_peak_GBs": peak}


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def seq():
    its = []
    X = np.zeros((n, nrhs))
    for j in range(nrhs):
        x, it, res = solvers.pcg(S, B[:, j].copy(), np.zeros(n), M)
        X[:, j] = x
        its.append(it)
    return X, np.array(its)


t_seq, (Xs, its_s) = timed(seq, 1)
iters = float((its_s - 1).sum())
print(json.dumps(dict(common, variant="sequential", value=nrhs / t_seq, wall_s=t_seq, iterations_total=iters, its_min=int(its_s.min()), its_max=int(its_s.max()),
                      ms_per_rhs_iteration=1e3 * t_seq / iters,
                      bytes_per_rhs_iteration=st["stream_bytes"] + st["dense_bytes"],
                      achieved_GBs=(st["stream_bytes"] + st["dense_bytes"]) * iters / t_seq / 1e9)), flush=True)
# RB="0,2,4": right-hand sides per pass of the multi-vector stream kernel to try (0: the library's choice)
variants = [(f"block_mf_rb{rb}" if rb else "block_mf", False, rb) for rb in (int(v) for v in os.environ.get("RB", "0").split(","))]
for variant, assembled, rb in variants + [("block_dense", True, 0)]:
    S.h.set_option("rhs_block", rb)
    t_b, (X, its, _) = timed(lambda: S.h.pcg_multi(B, np.zeros((n, nrhs)), assembled=assembled), 2)
    iters_b = float((its - 1).sum())
    sweeps = float((its - 1).max())                      # block iterations: every launch serves all still-running columns
    if assembled:
        bytes_per_sweep = 2 * st["dense_bytes"]          # S_d and Pi S_d blocks read once per block iteration (<= 64 columns)
        flops_per_sweep = 2 * 2 * st["dense_bytes"] / 8 * nrhs
    else:
        R = rb if rb else 4                                # upper bound of the library's own choice
        bytes_per_sweep = st["stream_bytes"] * nrhs / R + st["dense_bytes"]
        flops_per_sweep = 2 * st["dense_bytes"] / 8 * nrhs
    print(json.dumps(dict(common, variant=variant, value=nrhs / t_b, wall_s=t_b, speedup_vs_sequential=t_seq / t_b, iterations_total=iters_b,
                          block_iterations=sweeps, its_max_abs_diff_vs_sequential=int(np.abs(its - its_s).max()),
                          x_rel_diff_vs_sequential=float(np.linalg.norm(X - Xs) / np.linalg.norm(Xs)),
                          ms_per_block_iteration=1e3 * t_b / sweeps, ms_per_rhs_iteration=1e3 * t_b / iters_b,
                          bytes_per_rhs_iteration=bytes_per_sweep / nrhs, achieved_GBs=bytes_per_sweep * sweeps / t_b / 1e9,
                          dmma_tflops=flops_per_sweep * sweeps / t_b / 1e12,
                          assemble_local_schurs_s=t_asm if assembled else None)), flush=True)

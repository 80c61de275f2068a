"""Example07-style loop (Example07_PcgSchurStochasticEllipticPde.jl:160-204, 273) on the GPU: Example07-style loop at the benchmark size -- new realization -> value-only update of the
resident stream -> solve; reports the update time next to the solve time (f1 entry)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import krylov_b200  # noqa: F401
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 256
base = PB.build_mesh_and_partition(nx, ndom)
sd = base[-1]
pr0 = PB.realization(base, ireal=0, median=True)
S0 = fem.B200SchurOperator(sd, pr0.blocks, pr0.factors)
Sd0 = S0.assemble_local_schurs()
PiS = {d: PB.pinv_sym(Sd0[d]) for d in range(ndom)}
del S0, Sd0
prs = [PB.realization(base, ireal=i) for i in (1, 2)]
S = fem.B200SchurOperator(sd, prs[0].blocks, prs[0].factors)
M = fem.B200NeumannNeumann(S, PiS)
out = []
for rep in range(4):
    pr = prs[(rep + 1) % 2]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S.update_realization(pr.blocks, pr.factors)
    t_up = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = S.schur_rhs(pr.blocks.b_I, pr.blocks.b_G)          # first operation: runs the device refresh
    t_rhs = time.perf_counter() - t0
    t0 = time.perf_counter()
    x, it, res, tm = solvers.pcg(S, b, np.zeros(sd.n_gamma), M, maxit=5000, return_timings=True)
    t_solve = time.perf_counter() - t0
    st = S.h.stats()
    out.append({"rep": rep, "update_call_s": t_up, "first_op_incl_refresh_s": t_rhs, "solve_s": t_solve, "it": int(it),
                "solve_device_ms": tm.solve_ms, "host_analyses": st["host_analyses"], "value_refreshes": st["value_refreshes"]})
    print(json.dumps(out[-1]), flush=True)
# The same loop with the numeric Cholesky on the device (kb200_refactor_interior): the host sends the values of A_II,d,
# A_IΓ,d, A_ΓΓ,d -- no host factorisation, no factor upload.  For comparison: the host-side factorisation of all
# subdomains (what a caller of the value-only path pays per realization before it can call update_realization).
from krylov_b200.hostprep import chol
t0 = time.perf_counter()
chol.factor_subdomains(prs[1].blocks.A_II, base[1], sd.node_Id)
t_host_chol = time.perf_counter() - t0
S.enable_device_refactorization(prs[0].blocks.A_II)
for rep in range(4):
    pr = prs[(rep + 1) % 2]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S.refactor_realization(pr.blocks)
    t_up = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = S.schur_rhs(pr.blocks.b_I, pr.blocks.b_G)          # first operation: device Cholesky + refresh
    t_rhs = time.perf_counter() - t0
    t0 = time.perf_counter()
    x, it, res, tm = solvers.pcg(S, b, np.zeros(sd.n_gamma), M, maxit=5000, return_timings=True)
    t_solve = time.perf_counter() - t0
    st = S.h.stats()
    out.append({"rep": rep, "mode": "device_refactorization", "update_call_s": t_up, "first_op_incl_cholesky_and_refresh_s": t_rhs,
                "solve_s": t_solve, "it": int(it), "host_analyses": st["host_analyses"], "device_refactorizations": st["device_refactorizations"],
                "host_cholesky_all_subdomains_s": t_host_chol})
    print(json.dumps(out[-1]), flush=True)
# The same loop with the ASSEMBLY on the device as well (kb200_assemble_realization): the host sends one vector, the nodal
# coefficient; the right-hand sides come from the boundary layer (hostprep.problem.RhsUpdater).  For comparison: the host
# assembly of all blocks (prepare_local_schurs, numpy-vectorised stand-in of the reference's element loop) that the two
# loops above need per realization before their update calls -- not counted in their numbers.
cells, points, epart = base[0], base[1], base[4]
t0 = time.perf_counter()
PB.realization(base, ireal=2, factorize=False)
t_host_asm = time.perf_counter() - t0
t0 = time.perf_counter()
S.enable_device_assembly(cells, points, epart, prs[0].blocks.A_II)
t_enable = time.perf_counter() - t0
rhs = PB.RhsUpdater(base, prs[0].f_vals, prs[0].u_dir)
for rep in range(4):
    pr = prs[(rep + 1) % 2]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S.assemble_realization(pr.coeff)
    t_up = time.perf_counter() - t0
    t0 = time.perf_counter()
    b_I, b_G = rhs(pr.coeff)
    t_b = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = S.schur_rhs(b_I, b_G)                               # first operation: device Cholesky + refresh
    t_rhs = time.perf_counter() - t0
    t0 = time.perf_counter()
    x, it, res, tm = solvers.pcg(S, b, np.zeros(sd.n_gamma), M, maxit=5000, return_timings=True)
    t_solve = time.perf_counter() - t0
    st = S.h.stats()
    out.append({"rep": rep, "mode": "device_assembly", "assemble_call_s": t_up, "host_rhs_s": t_b,
                "first_op_incl_cholesky_and_refresh_s": t_rhs, "solve_s": t_solve, "it": int(it),
                "host_analyses": st["host_analyses"], "device_assemblies": st["device_assemblies"],
                "bytes_uploaded": 8 * pr.coeff.size, "host_assembly_all_blocks_s": t_host_asm, "enable_once_s": t_enable})
    print(json.dumps(out[-1]), flush=True)
nnzL = sum(f[1].nnz for f in prs[0].factors)
print(json.dumps({"nx": nx, "ndom": ndom, "nnz_L": nnzL, "bytes_uploaded_per_update": 8 * (nnzL + sum(a.nnz for a in prs[0].blocks.A_IG) + sum(a.nnz for a in prs[0].blocks.A_GG)),
                  "value_refs": st["value_refs"], "update_over_solve": out[3]["update_call_s"] / out[3]["solve_s"],
                  "update_plus_refresh_over_solve": (out[3]["update_call_s"] + out[3]["first_op_incl_refresh_s"]) / out[3]["solve_s"],
                  "device_refactorization_update_plus_first_op_over_solve": (out[7]["update_call_s"] + out[7]["first_op_incl_cholesky_and_refresh_s"]) / out[7]["solve_s"],
                  "device_assembly_all_steps_over_solve": (out[-1]["assemble_call_s"] + out[-1]["host_rhs_s"] + out[-1]["first_op_incl_cholesky_and_refresh_s"]) / out[-1]["solve_s"]}), flush=True)

"""Realization loop with the assembly on the device at the benchmark size (row f1): a new realization is the nodal
coefficient alone -- kb200_assemble_realization (element loop of prepare_local_schurs, :430-497), then the numeric
Cholesky and the stream refresh in front of the next operation.  Prints one JSON line per step.

    python scripts/bench_device_assembly.py [nx=2001] [ndom=256]

Without a CUDA device only the host-side steps are timed (contribution lists, right-hand-side updater)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_b200  # noqa: F401
from krylov_b200.hostprep import epdd_host, mesh
from krylov_b200.hostprep import problem as PB


def say(**kw):
    print(json.dumps(kw), flush=True)


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
    ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    t0 = time.perf_counter()
    base = PB.build_mesh_and_partition(nx, ndom)
    cells, points, epart, sd = base[0], base[1], base[4], base[-1]
    say(step="mesh + partition + set_subdomains", s=time.perf_counter() - t0, n_gamma=sd.n_gamma, nel=int(cells.shape[1]))
    t0 = time.perf_counter()
    pr1 = PB.realization(base, ireal=1, factorize=gpu)
    say(step="host: coefficient + prepare_local_schurs" + (" + Cholesky" if gpu else "") + " of realization 1 (what a realization costs without device assembly)",
        s=time.perf_counter() - t0)
    t0 = time.perf_counter()
    coeff2 = np.exp(mesh.lognormal_field(points, 2))
    say(step="host: coefficient of realization 2", s=time.perf_counter() - t0)
    if not gpu:
        t0 = time.perf_counter()
        maps = epdd_host.assembly_maps(cells, epart, sd)
        say(step="host: assembly_maps (once per mesh)", s=time.perf_counter() - t0,
            contributions=int(sum(maps.src[d][w].size for d in range(ndom) for w in range(3))))
        t0 = time.perf_counter()
        up = PB.RhsUpdater(base, pr1.f_vals, pr1.u_dir)
        say(step="host: RhsUpdater set-up (once per mesh)", s=time.perf_counter() - t0, layer_elements=int(up.layer.sum()))
        t0 = time.perf_counter()
        up(coeff2)
        say(step="host: right-hand sides of a new coefficient", s=time.perf_counter() - t0)
        return 0
    from krylov_b200 import fem
    t0 = time.perf_counter()
    S = fem.B200SchurOperator(sd, pr1.blocks, pr1.factors)
    x = np.random.default_rng(0).standard_normal(sd.n_gamma)
    y1 = S * x
    say(step="device set-up (host analysis + upload) + first apply", s=time.perf_counter() - t0)
    t0 = time.perf_counter()
    S.enable_device_assembly(cells, points, epart, pr1.blocks.A_II)
    say(step="enable_device_assembly (patterns, contribution lists, mesh; once per mesh)", s=time.perf_counter() - t0)
    for rep in range(4):
        coeff = coeff2 if rep % 2 == 0 else pr1.coeff
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        S.assemble_realization(coeff)
        t_asm = time.perf_counter() - t0
        t0 = time.perf_counter()
        y = S * x                                            # first operation: device Cholesky + refresh + one apply
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        y = S * x
        t_apply = time.perf_counter() - t0
        st = S.h.stats()
        rec = dict(step="realization", rep=rep, assemble_call_s=t_asm, first_apply_incl_cholesky_and_refresh_s=t_first,
                   plain_apply_s=t_apply, bytes_uploaded=int(8 * coeff.size), host_analyses=st["host_analyses"],
                   device_assemblies=st["device_assemblies"], device_refactorizations=st["device_refactorizations"])
        if rep % 2 == 1:   # back at realization 1: the operator built from the host's blocks and factors, to rounding
            rec["rel_diff_to_host_assembled_operator"] = float(np.abs(y - y1).max() / np.abs(y1).max())
        say(**rec)
    return 0


if __name__ == "__main__":
    sys.exit(main())

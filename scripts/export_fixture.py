"""Replay kit, part 1: write a workload in the REFERENCE's own on-disk formats (Fem/Mesh.jl:49-56,
221-229: `data/DoF{N}.cells.npz` ... `data/DoF{N}-ndom{d}.npart.npz`, bare .npy payloads) together with
the coefficient vector and the oracle's answers, so that a machine WITH Julia can run the reference's
own functions on identical inputs (`julia-phd-krylov-spdes_b200/julia/replay_fixture.jl`) and pin the
oracle -- the reference ships no golden data and cannot run in this image.

usage: python scripts/export_fixture.py --nx 65 --ndom 16 --out /tmp/replay65
       (then, in a checkout of the reference:  julia replay_fixture.jl /tmp/replay65 4225 16)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylov_b200  # noqa: E402,F401
from krylov_b200.hostprep import mesh as M  # noqa: E402
from krylov_b200.hostprep import problem as PB  # noqa: E402
import oracle.epdd as E  # noqa: E402
import oracle.solvers as OS  # noqa: E402


def export(nx, ndom, out, ireal=1):
    data = os.path.join(out, "data")
    os.makedirs(data, exist_ok=True)
    pr = PB.build_problem(nx, ndom, ireal=ireal)
    N = pr.points.shape[1]                       # the reference's `tentative_nnode` is only a file label
    M.save_mesh(data, N, pr.cells, pr.points, pr.markers, pr.neighbors)
    M.save_partition(data, N, ndom, pr.epart, pr.npart)

    def w(name, arr):
        with open(os.path.join(data, f"DoF{N}-ndom{ndom}.{name}.npz"), "wb") as f:
            np.save(f, np.asarray(arr))
    w("coeff", pr.coeff)                          # a at the nodes (coeff::Array{Float64,1}, EPDD.jl:389-398)
    inner = E.ExactInterior(pr.blocks.A_II)
    Sd = E.assemble_local_schurs(pr.blocks, inner)
    Pi = E.prepare_neumann_neumann_schur_precond(Sd)
    bs = E.get_schur_rhs(pr.blocks, inner, pr.sd)
    x = np.random.default_rng(481456).standard_normal(pr.sd.n_gamma)
    ug, it, res = OS.pcg(lambda v: E.apply_local_schurs(pr.blocks, inner, pr.sd, v), bs, np.zeros_like(bs),
                         lambda r: E.apply_neumann_neumann_schur(Pi, pr.sd, r))
    # expected answers (1-based where the reference is 1-based)
    w("expected.node_gamma", pr.sd.node_gamma + 1)
    w("expected.node_gamma_cnt", pr.sd.node_gamma_cnt)
    w("expected.b_schur", bs)
    w("expected.x_probe", x)
    w("expected.S_x", E.apply_local_schurs(pr.blocks, inner, pr.sd, x))
    w("expected.M_x", E.apply_neumann_neumann_schur(Pi, pr.sd, x))
    w("expected.pcg_it", np.array([it], dtype=np.int64))
    w("expected.pcg_res", res)
    w("expected.pcg_x", ug)
    return N, it


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=65)
    ap.add_argument("--ndom", type=int, default=16)
    ap.add_argument("--out", default="replay_fixture")
    ap.add_argument("--ireal", type=int, default=1)
    a = ap.parse_args()
    N, it = export(a.nx, a.ndom, a.out, a.ireal)
    print(f"written {a.out}/data/DoF{N}*  (oracle NN-PCG it = {it});  replay:  julia replay_fixture.jl {a.out} {N} {a.ndom}")

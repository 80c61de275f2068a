"""Single-process multi-GPU handle (kb200_create_multi; SURVEY 8(b): `kb200_create(ngpus, device_ids)`): one call of
the C ABI drives every member GPU.  Checked against the single-GPU handle of the same library and the CPU oracle:

* one member: the plain handle behind the forwarding layer (any box);
* two members on the SAME device: the multi-member machinery (sharding, owner routing of the uploads, in-process
  rendezvous, fused assembly + peer exchange kernels) on a one-GPU box, for the calls that allocate nothing while
  kernels are in flight (operators, right-hand side, PCG, back-substitution, dense assembly).  Two members sharing a
  device is a test configuration, not a deployment: an implicit device-wide synchronisation (cudaFree / first use of a
  kernel under lazy module loading) in one member's thread then waits for the other member's exchange kernel, which
  waits for the first member's next launch -- on separate devices the two are independent;
* two members on two devices when the box has them.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_SCRIPT = r"""
import sys, json
sys.path.insert(0, {root!r})
import numpy as np
import krylov_b200  # noqa
from krylov_b200 import fem, solvers
from krylov_b200.hostprep import problem as PB
import oracle.epdd as E
import oracle.solvers as OS
devices = {devices!r}
full = {full!r}
pr = PB.build_problem(65, 16, ireal=1)
sd, lb = pr.sd, pr.blocks
inner = E.ExactInterior(lb.A_II)
Sd = E.assemble_local_schurs(lb, inner)
Pi = E.prepare_neumann_neumann_schur_precond(Sd)
n = sd.n_gamma
x = np.random.default_rng(3).standard_normal(n)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
S1 = fem.B200SchurOperator(sd, lb, pr.factors, device=0)
M1 = fem.B200NeumannNeumann(S1, Pi)
S = fem.B200SchurOperator(sd, lb, pr.factors, devices=devices)
M = fem.B200NeumannNeumann(S, Pi)
out = {{}}
assert S.local_range == (0, sd.ndom)
y, y1, yo = S * x, S1 * x, E.apply_local_schurs(lb, inner, sd, x)
out["apply_vs_single"] = rel(y, y1); out["apply_vs_oracle"] = rel(y, yo)
out["apply_bitwise"] = bool(np.array_equal(y, y1))
z, z1 = M.solve(x), M1.solve(x)
out["nn_vs_single"] = rel(z, z1)
bs, bs1 = S.schur_rhs(lb.b_I, lb.b_G), S1.schur_rhs(lb.b_I, lb.b_G)
out["rhs_vs_single"] = rel(bs, bs1)
xs, it, res = solvers.pcg(S, bs1, np.zeros(n), M)
x1, it1, res1 = solvers.pcg(S1, bs1, np.zeros(n), M1)
out["pcg_it"] = [int(it), int(it1)]; out["pcg_x"] = rel(xs, x1); out["pcg_res_ok"] = bool(res[-1] <= 1e-7 * np.linalg.norm(bs1))
uI, uI1 = S.subdomain_solutions(xs, lb.b_I), S1.subdomain_solutions(xs, lb.b_I)
out["backsub"] = max(rel(a, b) for a, b in zip(uI, uI1))
if not full:
    Sa = S.assemble_local_schurs(); Sa1 = S1.assemble_local_schurs()
    out["dense_sd"] = max(rel(Sa[d], Sa1[d]) for d in range(sd.ndom))
    out["assembled_apply"] = rel(S.h.apply_schur_assembled(x), y1)
    out["peer"] = S.h.stats()["peer_exchange"]
    print("RESULT " + json.dumps(out)); sys.exit(0)
X, its, _ = S.h.pcg_multi(np.column_stack([bs1, x]), np.zeros((n, 2)))
X1, its1, _ = S1.h.pcg_multi(np.column_stack([bs1, x]), np.zeros((n, 2)))
out["multi_it"] = [its.tolist(), its1.tolist()]; out["multi_x"] = rel(X, X1)
W = np.linalg.qr(np.random.default_rng(5).standard_normal((n, 6)))[0]
xd, itd, _ = solvers.defpcg(S, bs1, np.zeros(n), W, M)
xd1, itd1, _ = solvers.defpcg(S1, bs1, np.zeros(n), W, M1)
out["defpcg_it"] = [int(itd), int(itd1)]; out["defpcg_x"] = rel(xd, xd1)
xe, ite, _, V = solvers.eigpcg(S, bs1, np.zeros(n), M, 4, 12)
xe1, ite1, _, V1 = solvers.eigpcg(S1, bs1, np.zeros(n), M1, 4, 12)
out["eigpcg_it"] = [int(ite), int(ite1)]; out["eigpcg_x"] = rel(xe, xe1); out["eigpcg_V"] = list(V.shape)
Sa = S.assemble_local_schurs(); Sa1 = S1.assemble_local_schurs()
out["dense_sd"] = max(rel(Sa[d], Sa1[d]) for d in range(sd.ndom))
ya = S.h.apply_schur_assembled(x)
out["assembled_apply"] = rel(ya, y1)
st = S.h.stats()
out["peer"] = st["peer_exchange"]; out["ctas"] = st.get("n_cta", 0)
try:
    S.h.mat_alloc(2)
    out["block_api"] = "allowed"
except Exception as e:
    out["block_api"] = "refused" if "multi-GPU handle" in str(e) else str(e)
# realization 2 from its nodal coefficient alone (device assembly + numeric Cholesky, every member its own subdomains)
# against a single-GPU handle fed with the host's blocks and factors of that realization
pr2 = PB.realization((pr.cells, pr.points, pr.markers, pr.neighbors, pr.epart, pr.npart, sd), ireal=2)
S.enable_device_assembly(pr.cells, pr.points, pr.epart, lb.A_II)
S.assemble_realization(pr2.coeff)
S2 = fem.B200SchurOperator(sd, pr2.blocks, pr2.factors, device=0)
out["asm_vs_single"] = rel(S * x, S2 * x)
dl = sd.ndom - 1
blk = [pr2.blocks.A_II[dl].tocsc(), pr2.blocks.A_IG[dl].tocsc(), pr2.blocks.A_GG[dl].tocsc()]
for A in blk:
    A.sort_indices()
got = S.h.get_local_block_values(dl, *[A.nnz for A in blk])
out["asm_values"] = max(rel(g, A.data) for g, A in zip(got, blk))
out["asm_count"] = S.h.stats()["device_assemblies"]
# ... and the right-hand sides of that realization (every member its own slice of b_I, the first one b_Γ)
S.enable_device_rhs(pr.cells, pr.points, pr.epart, pr.f_vals, pr.u_dir)
S.assemble_realization(pr2.coeff)
bI, bG = S.h.get_rhs()
out["rhs_dev_bits"] = bool(np.array_equal(bI, np.concatenate(pr2.blocks.b_I)) and np.array_equal(bG, pr2.blocks.b_G))
out["rhs_dev_schur"] = rel(S.schur_rhs(), S2.schur_rhs(pr2.blocks.b_I, pr2.blocks.b_G))
print("RESULT " + json.dumps(out))
"""


def _run(devices, timeout=420, full=True):
    import json
    code = _SCRIPT.format(root=ROOT, devices=devices, full=full)
    env = dict(os.environ)
    env["CUDA_MODULE_LOADING"] = "EAGER"
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "tests")] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, env=env)
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert p.returncode == 0 and lines, f"multi-GPU handle run failed (devices {devices}):\n{p.stdout[-2000:]}\n{p.stderr[-4000:]}"
    return json.loads(lines[-1][7:])


def _check(o, members, full=True):
    assert o["apply_vs_single"] <= 1e-13 and o["apply_vs_oracle"] <= 1e-12
    assert o["nn_vs_single"] <= 1e-13 and o["rhs_vs_single"] <= 1e-12 and o["backsub"] <= 1e-12
    assert abs(o["pcg_it"][0] - o["pcg_it"][1]) <= 1 and o["pcg_res_ok"] and o["pcg_x"] <= 1e-6
    if not full:
        assert o["dense_sd"] <= 1e-11 and o["assembled_apply"] <= 1e-11 and o["peer"] == 1.0
        return
    assert np.abs(np.array(o["multi_it"][0]) - np.array(o["multi_it"][1])).max() <= 1 and o["multi_x"] <= 1e-6
    assert abs(o["defpcg_it"][0] - o["defpcg_it"][1]) <= 1 and o["defpcg_x"] <= 1e-6
    assert abs(o["eigpcg_it"][0] - o["eigpcg_it"][1]) <= 1 and o["eigpcg_x"] <= 1e-6 and o["eigpcg_V"][1] == 4
    assert o["dense_sd"] <= 1e-11 and o["assembled_apply"] <= 1e-11
    assert o["block_api"] == "refused"
    assert o["asm_vs_single"] <= 1e-12 and o["asm_values"] <= 4e-16 and o["asm_count"] == 1.0
    assert o["rhs_dev_bits"] and o["rhs_dev_schur"] <= 1e-12
    if members > 1:
        assert o["peer"] == 1.0, "the members must exchange through peer memory"
    else:
        assert o["apply_bitwise"]


def test_multi_handle_one_member():
    _check(_run([0]), 1)


def test_multi_handle_two_members_on_one_device():
    _check(_run([0, 0], full=False), 2, full=False)


def test_multi_handle_two_devices():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _check(_run([0, 1]), 2)

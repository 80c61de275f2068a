"""pcg / defpcg / eigpcg / eigdefpcg restatements (cg.jl:67-109, defcg.jl:242-308, 337-473,
eigcg.jl:143-267) on small SPD systems."""
import numpy as np
import pytest

import oracle.solvers as OS


def _spd(n, cond, seed, nsmall=0):
    """SPD matrix with `nsmall` isolated small eigenvalues below a cluster in [cond/2, cond]."""
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    if nsmall:
        lam = np.concatenate([np.linspace(1.0, 3.0, nsmall), np.linspace(cond / 2, cond, n - nsmall)])
    else:
        lam = np.geomspace(1.0, cond, n)
    return (Q * lam) @ Q.T, Q, lam


def test_pcg_conventions():
    A, Q, lam = _spd(60, 50.0, 0)
    b = np.random.default_rng(1).standard_normal(60)
    x0 = np.zeros(60)
    x, it, res = OS.pcg(lambda v: A @ v, b, x0, lambda r: r / np.diag(A))
    assert x is x0                                   # updated in place (cg.jl:97)
    assert res.size == it and np.isclose(res[0], np.linalg.norm(b), rtol=1e-14)
    assert res[-1] <= 1e-7 * np.linalg.norm(b) < res[-2]
    assert np.linalg.norm(A @ x - b) <= 2e-7 * np.linalg.norm(b)
    # maxit caps `it` (cg.jl:92)
    _, it2, res2 = OS.pcg(lambda v: A @ v, b, np.zeros(60), lambda r: r.copy(), maxit=5)
    assert it2 == 5 and res2.size == 5


def test_defpcg_with_exact_eigenvectors_needs_fewer_iterations():
    A, Q, lam = _spd(80, 400.0, 2, nsmall=10)
    b = np.random.default_rng(3).standard_normal(80)
    _, it0, _ = OS.pcg(lambda v: A @ v, b, np.zeros(80), lambda r: r.copy())
    W = Q[:, :10]                                    # least dominant eigenvectors
    x, it1, res = OS.defpcg(lambda v: A @ v, b, np.zeros(80), W, lambda r: r.copy())
    assert it1 < it0
    assert np.linalg.norm(A @ x - b) <= 2e-7 * np.linalg.norm(b)


def test_eigpcg_recovers_least_dominant_eigenvectors_and_eigdefpcg_uses_them():
    n, nvec, spdim = 120, 6, 18
    A, Q, lam = _spd(n, 200.0, 4, nsmall=6)
    b = np.random.default_rng(5).standard_normal(n)
    M = lambda r: r.copy()
    x, it, res, V = OS.eigpcg(lambda v: A @ v, b, np.zeros(n), M, nvec, spdim)
    assert V.shape == (n, nvec) and np.linalg.norm(A @ x - b) <= 2e-7 * np.linalg.norm(b)
    # Ritz vectors: the first few align with the smallest eigenvectors of A
    for k in range(3):
        v = V[:, k] / np.linalg.norm(V[:, k])
        assert np.linalg.norm(Q[:, :nvec].T @ v) > 0.98
    b2 = np.random.default_rng(6).standard_normal(n)
    _, it_plain, _ = OS.pcg(lambda v: A @ v, b2, np.zeros(n), M)
    x2, it_def, res2, V2 = OS.eigdefpcg(lambda v: A @ v, b2, np.zeros(n), M, V, spdim)
    assert it_def < it_plain
    assert np.linalg.norm(A @ x2 - b2) <= 5e-7 * np.linalg.norm(b2)
    assert V2.shape == (n, nvec)


@pytest.mark.parametrize("variant", ["fma", "ldot", "tdot", "bdot"])
def test_cpu_port_arithmetic_variants_agree_on_a_small_solve(variant):
    """The builds of oracle/kb_oracle.c used for the rounding-floor probes of DESIGN.md section 2 (fused multiply-adds,
    extended-precision / tree-shaped inner products) are the same algorithm: same stopping index and solution as the
    default build where the recurrence is short enough for rounding not to matter."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from conftest import get_problem, get_oracle_ops\n"
        "from oracle.cpu_model import CpuModel\n"
        "pr = get_problem(33, 4); inner, Sd, Pi, bs = get_oracle_ops(pr)\n"
        "cm = CpuModel(pr.sd, pr.blocks, pr.factors, Pi)\n"
        "x, it, res = cm.pcg(bs, np.zeros(pr.sd.n_gamma))\n"
        "print('RESULT ' + json.dumps({'it': int(it), 'x': x.tolist()}))\n") % (root, os.path.join(root, "tests"))
    outs = []
    for v in ("", variant):
        env = dict(os.environ, KB_ORACLE_VARIANT=v)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][7:]))
    assert abs(outs[0]["it"] - outs[1]["it"]) <= 1
    a, b = np.array(outs[0]["x"]), np.array(outs[1]["x"])
    assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(a)

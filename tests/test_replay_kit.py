"""The replay kit (scripts/export_fixture.py + julia/replay_fixture.jl) writes a workload in the
reference's own on-disk formats (Fem/Mesh.jl:49-56, 221-229) plus the oracle's answers, so the oracle can
be pinned against the real Julia code on a machine that has it.  Here: the export round-trips through the
loaders that mirror the reference's, and the expected answers are self-consistent."""
import importlib.util
import os

import numpy as np

from conftest import ROOT
from krylov_b200.hostprep import epdd_host, mesh as M


def _load_export():
    spec = importlib.util.spec_from_file_location("export_fixture", os.path.join(ROOT, "scripts", "export_fixture.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_export_round_trip(tmp_path):
    ex = _load_export()
    N, it = ex.export(33, 4, str(tmp_path))
    data = os.path.join(str(tmp_path), "data")
    cells, points, markers, nb = M.load_mesh(data, N)
    epart, npart = M.load_partition(data, N, 4)
    assert cells.shape[0] == 3 and points.shape == (2, N) and markers.sum() == 4 * 33 - 4
    # reference conventions on disk: 0-based cells / parts, 1-based neighbours with -1 on the boundary
    raw = np.load(open(os.path.join(data, f"DoF{N}.cell_neighbors.npz"), "rb"))
    assert raw.shape == (cells.shape[1], 3) and raw.min() == -1 and 0 not in raw
    assert np.load(open(os.path.join(data, f"DoF{N}.cells.npz"), "rb")).min() == 0

    def rd(name):
        return np.load(open(os.path.join(data, f"DoF{N}-ndom4.{name}.npz"), "rb"))
    sd = epdd_host.set_subdomains(cells, nb, epart, npart, markers == 1)
    assert np.array_equal(rd("expected.node_gamma"), sd.node_gamma + 1)          # 1-based, like node_Γ
    assert np.array_equal(rd("expected.node_gamma_cnt"), sd.node_gamma_cnt)
    assert rd("coeff").shape == (N,) and rd("coeff").min() > 0
    res = rd("expected.pcg_res")
    assert int(rd("expected.pcg_it")[0]) == it == res.size
    assert res[-1] <= 1e-7 * np.linalg.norm(rd("expected.b_schur")) < res[-2]    # stopped by the reference's test
    assert os.path.exists(os.path.join(ROOT, "julia-phd-krylov-spdes_b200", "julia", "replay_fixture.jl"))

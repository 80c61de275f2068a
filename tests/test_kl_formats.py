"""KL wire formats and draws (KarhunenLoeveDomainDecompositionHelper.jl:72-79,
KarhunenLoevePllDomainDecomposition.jl:613-614, KarhunenLoeveDomainDecomposition.jl:985-1043)."""
import numpy as np
import pytest

from krylov_b200.hostprep import kl, mesh, samplers


def test_root_filename_matches_the_reference_string():
    # Example07:17-19, 73: model "SExp", sig2 = 1., L = .1
    assert kl.get_root_filename("SExp", 1.0, 0.1, 4000) == "SExp_sig21.0_L0.1_DoF4000"
    assert kl.get_root_filename("SExp", 1, 0.05, 16000) == "SExp_sig21.0_L0.05_DoF16000"


def test_kl_round_trip_and_draw(tmp_path):
    cells, points, markers, nb = mesh.structured_mesh(17)
    Lam, Psi = samplers.rff_modes(points, m=16)
    root = kl.get_root_filename("SExp", 1.0, 0.1, points.shape[1])
    kl.save_kl(str(tmp_path), root, Lam, Psi)
    # bare .npy payloads under the .npz name, as NPZ.jl's npzwrite of an array produces
    raw = np.load(open(tmp_path / f"{root}.kl-eigvecs.npz", "rb"))
    assert raw.shape == Psi.shape and raw.dtype == np.float64
    L2, P2 = kl.load_kl(str(tmp_path), root)
    assert np.array_equal(L2, Lam) and np.array_equal(P2, Psi)
    # a column-major payload (what Julia writes) reads back identically
    with open(tmp_path / f"{root}.kl-eigvecs.npz", "wb") as f:
        np.save(f, np.asfortranarray(Psi))
    assert np.array_equal(kl.load_kl(str(tmp_path), root)[1], Psi)
    xi, g = kl.draw(Lam, Psi, np.random.default_rng(0))
    assert xi.shape == Lam.shape and np.allclose(g, Psi @ (np.sqrt(Lam) * xi), rtol=0, atol=1e-13 * np.abs(g).max())
    xi2, g2 = np.empty_like(xi), np.empty_like(g)
    kl.draw_into(Lam, Psi, xi2, g2, np.random.default_rng(0))
    assert np.array_equal(xi, xi2) and np.array_equal(g, g2)
    # the sampler accepts pairs loaded from disk (Example09:124-137)
    smp = samplers.McmcSampler(L2, P2, seed=1)
    assert smp.draw() >= 1 and smp.g.shape == (points.shape[1],)
    with pytest.raises(ValueError):
        kl.save_kl(str(tmp_path), root, Lam[:-1], Psi)

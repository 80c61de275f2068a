"""Karhunen-Loeve wire formats and draws of the reference (host side).

The examples read the global KL pairs from ``data/{root}.kl-eigvals.npz`` / ``.kl-eigvecs.npz`` (bare .npy payloads
under an .npz name, written by Fem/KarhunenLoevePllDomainDecomposition.jl:613-614, read at e.g.
Example09_DefPcgMcmcStochasticEllipticPde.jl:124-125, Example07_PcgSchurStochasticEllipticPde.jl:74-75) with
``root = get_root_filename(model, sig2, L, nnode)`` (Fem/KarhunenLoeveDomainDecompositionHelper.jl:72-79) and turn
them into coefficient realizations with ``draw`` / ``draw!`` (Fem/KarhunenLoeveDomainDecomposition.jl:985-999,
:1026-1043): g = sum_a sqrt(Lambda_a) xi_a Psi[:, a], xi ~ N(0, I); the PDE coefficient is exp(g)."""
import os

import numpy as np


def _julia_float(v):
    """Julia prints Float64 with the shortest round-trip digits and always a decimal point ("1.0", "0.1")."""
    s = repr(float(v))
    return s


def get_root_filename(model, sig2, L, nnode):
    """KarhunenLoeveDomainDecompositionHelper.jl:72-79: "SExp_sig21.0_L0.1_DoF4000"."""
    return f"{model}_sig2{_julia_float(sig2)}_L{_julia_float(L)}_DoF{int(nnode)}"


def save_kl(prefix, root, Lam, Psi):
    """KarhunenLoevePllDomainDecomposition.jl:613-614: Λ (nmode,), Ψ (nnode x nmode), Float64."""
    Lam = np.ascontiguousarray(Lam, dtype=np.float64)
    Psi = np.asarray(Psi, dtype=np.float64)
    if Psi.ndim != 2 or Psi.shape[1] != Lam.size:
        raise ValueError("Psi must be nnode x nmode with one column per eigenvalue")
    for name, arr in (("kl-eigvals", Lam), ("kl-eigvecs", Psi)):
        with open(os.path.join(prefix, f"{root}.{name}.npz"), "wb") as f:
            np.save(f, arr)


def load_kl(prefix, root):
    """npzread of the two files (Example09:124-125); accepts Fortran- or C-ordered payloads (NPZ.jl writes column-major)."""
    out = []
    for name in ("kl-eigvals", "kl-eigvecs"):
        with open(os.path.join(prefix, f"{root}.{name}.npz"), "rb") as f:
            out.append(np.load(f))
    Lam, Psi = np.asarray(out[0], dtype=np.float64).ravel(), np.asarray(out[1], dtype=np.float64)
    if Psi.ndim != 2 or Psi.shape[1] != Lam.size:
        raise ValueError(f"{root}: eigvecs {Psi.shape} do not match {Lam.size} eigenvalues")
    return Lam, Psi


def draw_into(Lam, Psi, xi, g, rng):
    """draw! (:1026-1043): xi .= randn(nmode); g = sum_a sqrt(Lam_a) xi_a Psi[:, a], accumulated mode by mode (axpy
    order of the reference)."""
    xi[...] = rng.standard_normal(Lam.size)
    g[...] = 0.0
    for a in range(Lam.size):
        g += (np.sqrt(Lam[a]) * xi[a]) * Psi[:, a]
    return xi, g


def draw(Lam, Psi, rng):
    """draw (:985-999): returns (xi, g)."""
    return draw_into(Lam, Psi, np.empty(Lam.size), np.empty(Psi.shape[0]), rng)

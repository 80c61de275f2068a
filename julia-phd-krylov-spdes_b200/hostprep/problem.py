"""Workload builder: the host-side steps of Example03 / Example07 up to the point where the
reference would call ``pcg`` (Example03_EllipticPdeDomainDecomposition.jl:34-150,
Example07_PcgSchurStochasticEllipticPde.jl:34-204), on the synthetic inputs of SURVEY.md 8(d)."""
from dataclasses import dataclass
from typing import List

import numpy as np

from . import chol, epdd_host, mesh


@dataclass
class Problem:
    nx: int
    ndom: int
    cells: np.ndarray
    points: np.ndarray
    markers: np.ndarray
    neighbors: np.ndarray
    epart: np.ndarray
    npart: np.ndarray
    sd: epdd_host.Subdomains
    blocks: epdd_host.LocalBlocks
    factors: List[tuple]
    f_vals: np.ndarray
    u_dir: np.ndarray
    coeff: np.ndarray

    @property
    def n_dof(self):
        return int((self.markers == 0).sum())


def build_mesh_and_partition(nx, ndom, jitter=0.2):
    cells, points, markers, neighbors = mesh.structured_mesh(nx, jitter=jitter)
    epart, npart = mesh.rcb_partition(cells, points, ndom)
    sd = epdd_host.set_subdomains(cells, neighbors, epart, npart, markers == 1)
    return cells, points, markers, neighbors, epart, npart, sd


def realization(base, ireal=1, median=False, factorize=True, dom_range=None):
    """Blocks + factors for one coefficient realization on a prepared mesh/partition tuple.
    dom_range=(d0, d1) assembles / factorizes only those subdomains' elements (multi-rank runs):
    b_G then holds this rank's partial sum and the other subdomains' blocks are empty."""
    cells, points, markers, neighbors, epart, npart, sd = base
    nnode = points.shape[1]
    g = mesh.lognormal_field(points, ireal, median=median)
    coeff = np.exp(g)
    f_vals = -np.ones(nnode)            # f = -1     (Example07:65-67)
    u_dir = 0.734 * np.ones(nnode)      # uexact     (Example07:69-71)
    if dom_range is not None and (dom_range[0] > 0 or dom_range[1] < sd.ndom):
        keep = (epart >= dom_range[0]) & (epart < dom_range[1])
        blocks = epdd_host.prepare_local_schurs(cells[:, keep], points, epart[keep], sd, coeff, f_vals, u_dir)
    else:
        blocks = epdd_host.prepare_local_schurs(cells, points, epart, sd, coeff, f_vals, u_dir)
    factors = chol.factor_subdomains(blocks.A_II, points, sd.node_Id) if factorize else []
    nx = int(round(np.sqrt(nnode)))
    return Problem(nx, sd.ndom, cells, points, markers, neighbors, epart, npart, sd, blocks, factors,
                   f_vals, u_dir, coeff)


def interface_rhs(base, coeff, f_vals, u_dir):
    """b_Γ of the WHOLE mesh from the elements that touch an interface node only (a few percent of the mesh).  The
    accumulation of prepare_local_schurs is element-major, and elements without an interface node add nothing to b_Γ, so
    the result has the bits of the full assembly -- every rank of a sharded run can compute it by itself instead of
    summing per-rank partials in a rank-dependent order."""
    cells, points, markers, neighbors, epart, npart, sd = base
    touch = (sd.ind_gamma_g2l[cells] >= 0).any(axis=0)
    return epdd_host.prepare_local_schurs(cells[:, touch], points, epart[touch], sd, coeff, f_vals, u_dir).b_G


class RhsUpdater:
    """b_I,d / b_Γ of prepare_local_schurs (:499-536) for a new coefficient WITHOUT assembling the blocks (the companion of
    kb200_assemble_realization, which leaves the right-hand sides with the caller).  Only the Dirichlet lift (:499-508)
    depends on the coefficient, and only in the entries of nodes that share an element with a Dirichlet node; the
    accumulation is element-major, so re-running the element loop over ALL elements of those nodes (a boundary layer,
    O(sqrt(nel)) elements) reproduces their entries with the bits of the full assembly, and every other entry is the
    load vector, computed once."""

    def __init__(self, base, f_vals, u_dir):
        cells, points, markers, neighbors, epart, npart, sd = base
        self.base, self.f_vals, self.u_dir = base, f_vals, u_dir
        is_dir = sd.node_owner == 0
        el_dir = is_dir[cells].any(axis=0)
        tgt = np.zeros(sd.nnode, dtype=bool)
        tgt[cells[:, el_dir].ravel()] = True
        tgt &= ~is_dir                                       # nodes whose entry carries lift terms
        self.layer = tgt[cells].any(axis=0)                  # every element of such a node
        static = epdd_host.prepare_local_schurs(cells, points, epart, sd, np.ones(sd.nnode), f_vals, u_dir)
        self.b_G0, self.b_I0 = static.b_G, static.b_I
        self.mask_G = tgt[sd.node_gamma]
        self.mask_I = [tgt[sd.node_Id[d]] for d in range(sd.ndom)]

    def __call__(self, coeff):
        cells, points, markers, neighbors, epart, npart, sd = self.base
        part = epdd_host.prepare_local_schurs(cells[:, self.layer], points, epart[self.layer], sd, coeff, self.f_vals, self.u_dir)
        b_G = np.where(self.mask_G, part.b_G, self.b_G0)
        b_I = [np.where(self.mask_I[d], part.b_I[d], self.b_I0[d]) for d in range(sd.ndom)]
        return b_I, b_G


def build_problem(nx, ndom, ireal=1, jitter=0.2, median=False):
    return realization(build_mesh_and_partition(nx, ndom, jitter), ireal, median)


def pinv_sym(S, rtol=None):
    """pinv of a symmetric positive semi-definite block through its eigen-decomposition, with the
    reference's cut-off rtol = sqrt(eps) relative to the largest singular value
    (EllipticPdeDomainDecomposition.jl:1209-1213).  3x cheaper than the SVD route."""
    rtol = np.sqrt(np.finfo(np.float64).eps) if rtol is None else rtol
    w, V = np.linalg.eigh((S + S.T) * 0.5)
    keep = np.abs(w) > rtol * np.abs(w).max()
    Vk = V[:, keep]
    return (Vk / w[keep]) @ Vk.T


def algorithmic_bytes_per_iteration(prob: Problem, dom_range=None):
    """B_iter of SURVEY.md 8(d), evaluated from the actual nnz counts (int32 indices, FP64 values)."""
    b = 0.0
    parts = {"trisolve": 0.0, "spmv": 0.0, "nn": 0.0, "maps": 0.0, "vectors": 0.0}
    d0, d1 = dom_range if dom_range is not None else (0, prob.ndom)
    for d in range(d0, d1):
        nI = prob.blocks.A_II[d].shape[0]
        ng = prob.blocks.A_GG[d].shape[0]
        nnzL = prob.factors[d][1].nnz
        parts["trisolve"] += 2 * (12 * nnzL + 4 * (nI + 1) + 8 * 3 * nI)
        parts["spmv"] += 2 * (12 * prob.blocks.A_IG[d].nnz + 4 * (nI + 1)) + 12 * prob.blocks.A_GG[d].nnz + 4 * (ng + 1)
        parts["nn"] += 8 * ng * ng
        parts["maps"] += 4 * 4 * ng + 8 * 4 * ng
    parts["vectors"] = 8 * prob.sd.n_gamma * 12
    b = sum(parts.values())
    return b, parts

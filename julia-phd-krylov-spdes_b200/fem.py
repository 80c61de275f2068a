"""Host-side mirror of the reference's Fem API for the hot path: the two new types the Julia shim
introduces (julia/KrylovB200.jl) with the same call shapes --

  S = B200SchurOperator(...)      usable wherever the reference uses its LinearMaps.FunctionMap
                                  (``S * x``, ``mul!(y, S, x)``, ``S.N``: cg.jl:83, 93; defcg.jl:263)
  M = B200NeumannNeumann(S)       ``M \\ r``, ``ldiv!(z, M, r)``, ``ldiv!(M, r)``
                                  (EllipticPdeDomainDecomposition.jl:1389-1403)

plus get_schur_rhs / get_subdomain_solutions / assemble_local_schurs equivalents.  Everything
numerical happens in libkrylov_b200.so."""
import numpy as np

from . import capi


class B200SchurOperator:
    def __init__(self, sd, blocks, factors, device=0, csize=None, rank=0, world=1, nccl_id=None,
                 profile=False, options=None, devices=None):
        """sd: Subdomains (set_subdomains); blocks: LocalBlocks (prepare_local_schurs);
        factors: [(perm, L)] Cholesky factors of A_II,d (F.p, sparse(F.L)).
        devices=[0, 1, ...]: one process driving several GPUs (kb200_create_multi)."""
        self.sd = sd
        self.h = capi.Handle(device, rank, world, nccl_id, devices=devices)
        if csize is not None:
            self.h.set_option("csize", csize)
        if profile:
            self.h.set_option("profile", 1)
        for k, v in (options or {}).items():   # "shard_world" before "shard_rank"
            self.h.set_option(k, v)
        n_I = [v.size for v in sd.node_Id]
        n_gd = [g.size for g in sd.gamma_l2g]
        self.h.set_topology(n_I, n_gd, [g + 1 for g in sd.gamma_l2g], sd.node_gamma_cnt)
        d0, d1 = self.h.local_range()
        for d in range(d0, d1):
            self.h.upload_local_blocks(d, blocks.A_IG[d], blocks.A_GG[d])
            self.h.upload_cholesky(d, factors[d][0], factors[d][1])
        self.local_range = (d0, d1)
        self.N = sd.n_gamma
        self.assembled = False

    def update_realization(self, blocks, factors):
        """New coefficient realization on the same mesh / partition / ordering (the loop of
        Example07_PcgSchurStochasticEllipticPde.jl:160-204): values only."""
        d0, d1 = self.local_range
        for d in range(d0, d1):
            self.h.update_local_block_values(d, blocks.A_IG[d], blocks.A_GG[d])
            self.h.update_cholesky_values(d, factors[d][1])
        # the dense S_d of the previous realization are dropped by the library (the reference re-assembles per
        # realization, EllipticPdeDomainDecomposition.jl:667-695): back to the matrix-free operator
        self.assembled = False

    def enable_device_refactorization(self, A_II):
        """Pattern of every local A_II,d (kb200_upload_interior_pattern): from now on a new realization needs no host
        Cholesky -- ``refactor_realization`` recomputes the factors' values on the GPU."""
        d0, d1 = self.local_range
        for d in range(d0, d1):
            self.h.upload_interior_pattern(d, A_II[d])

    def refactor_realization(self, blocks):
        """New coefficient realization on the same mesh / partition / ordering with the numeric Cholesky on the device
        (kb200_refactor_interior; the reference calls cholesky(A_IIdd[d]) per realization, Example07:160-204 /
        EllipticPdeDomainDecomposition.jl:1536)."""
        d0, d1 = self.local_range
        for d in range(d0, d1):
            self.h.update_local_block_values(d, blocks.A_IG[d], blocks.A_GG[d])
            self.h.refactor_interior(d, blocks.A_II[d])
        self.assembled = False

    def enable_device_assembly(self, cells, points, epart, A_II):
        """Mesh + contribution lists of every local block (kb200_upload_assembly_mesh / kb200_upload_assembly_map) on top
        of ``enable_device_refactorization``: from now on a new realization is ONE vector -- ``assemble_realization(coeff)``
        runs the element loop of prepare_local_schurs (:430-497), the numeric Cholesky and the stream refresh on the GPU."""
        from .hostprep import epdd_host
        self.enable_device_refactorization(A_II)
        d0, d1 = self.local_range
        maps = epdd_host.assembly_maps(cells, epart, self.sd, dom_range=(d0, d1))
        self.h.upload_assembly_mesh(cells, points)
        for d in range(d0, d1):
            for which in range(3):
                self.h.upload_assembly_map(d, which, maps.ptr[d][which], maps.src[d][which])

    def enable_device_rhs(self, cells, points, epart, f_vals, u_dir):
        """Right-hand sides on the device as well (kb200_upload_rhs_map): after this ``assemble_realization`` also assembles
        b_I,d / b_Γ (:499-536), and ``schur_rhs()`` / ``subdomain_solutions(u_gamma)`` without right-hand-side arguments use
        them."""
        from .hostprep import epdd_host
        m = epdd_host.rhs_maps(cells, points, epart, self.sd, f_vals, u_dir)
        self.h.upload_rhs_map(m.bI0, m.bG0, m.tgt, m.ptr, m.src, m.val)

    def assemble_realization(self, coeff):
        """New coefficient realization from the nodal coefficient alone (kb200_assemble_realization)."""
        self.h.assemble_realization(coeff)
        self.assembled = False

    # LinearMaps.FunctionMap surface
    def __matmul__(self, x):
        return self.mul(x)

    def __mul__(self, x):
        return self.mul(x)

    def mul(self, x):
        return self.h.apply_schur_assembled(x) if self.assembled else self.h.apply_schur(x)

    def mul_into(self, y, x):  # mul!(y, S, x)
        y[...] = self.mul(x)
        return y

    # get_schur_rhs / get_subdomain_solutions (EllipticPdeDomainDecomposition.jl:835-864, 1014-1025)
    def schur_rhs(self, b_I=None, b_gamma=None):
        if b_I is None and b_gamma is None:            # the right-hand sides of the last assemble_realization (device)
            return self.h.schur_rhs()
        return self.h.schur_rhs(np.concatenate(b_I) if len(b_I) else np.zeros(0), b_gamma)

    def subdomain_solutions(self, u_gamma, b_I=None):
        u = self.h.back_substitute(u_gamma, None if b_I is None else np.concatenate(b_I))
        ptr = np.concatenate([[0], np.cumsum([v.size for v in self.sd.node_Id])])
        return [u[ptr[d]:ptr[d + 1]] for d in range(self.sd.ndom)]

    # assemble_local_schurs (:667-695): dense S_d computed on the device
    def assemble_local_schurs(self):
        self.h.assemble_local_schurs()
        d0, d1 = self.local_range
        return {d: self.h.get_dense_local_schur(d) for d in range(d0, d1)}


class B200GlobalSchurOperator(B200SchurOperator):
    """The operator of apply_global_schur (EllipticPdeDomainDecomposition.jl:596-625) from the outputs of
    prepare_global_schur (:212-369): A_IΓ,d with global interface columns and the assembled A_ΓΓ.
    ``S * x`` = A_ΓΓ x - Σ_d A_IΓd' A_IId^{-1} A_IΓd x  (the identity Example03:175 checks)."""

    def __init__(self, sd, A_IG_global, A_GG, factors, device=0, csize=None, options=None):
        self.sd = sd
        self.h = capi.Handle(device)
        if csize is not None:
            self.h.set_option("csize", csize)
        for k, v in (options or {}).items():
            self.h.set_option(k, v)
        self.h.set_topology([v.size for v in sd.node_Id], [g.size for g in sd.gamma_l2g],
                            [g + 1 for g in sd.gamma_l2g], sd.node_gamma_cnt)
        d0, d1 = self.h.local_range()
        for d in range(d0, d1):
            self.h.upload_global_blocks(d, A_IG_global[d])
            self.h.upload_cholesky(d, factors[d][0], factors[d][1])
        self.h.upload_global_interface_block(A_GG)
        self.local_range = (d0, d1)
        self.N = sd.n_gamma
        self.assembled = False

    def mul(self, x):
        return self.h.apply_global_schur(x)


class B200NeumannNeumann:
    """NeumannNeumannSchurPreconditioner (EllipticPdeDomainDecomposition.jl:1111-1137) resident on
    the GPU of ``S``; ``PiSd[d]`` as produced by prepare_neumann_neumann_schur_precond (:1201-1220)."""

    def __init__(self, S, PiSd):
        self.S = S
        d0, d1 = S.local_range
        for d in range(d0, d1):
            S.h.upload_nn_pinv(d, PiSd[d])

    def solve(self, r):        # M \ r
        return self.S.h.apply_nn(r)

    def ldiv_into(self, z, r):  # ldiv!(z, M, r)
        z[...] = self.solve(r)
        return z

    def ldiv_inplace(self, r):  # ldiv!(M, r)
        r[...] = self.solve(r.copy())
        return r


class B200Lorasc:
    """LorascPreconditioner (EllipticPdeDomainDecomposition.jl:1406-1428) for the full system, applied on the GPU(s) of
    ``S`` (kb200_apply_lorasc): interior factors and A_IΓ,d are the ones ``S`` already holds; ``chol_A_GG`` = (F.p,
    sparse(F.L)) of cholesky(A_ΓΓ) (:1525) and ``E`` (n_Γ x nev) the correction vectors of prepare_lorasc_precond
    (:1502-1676), both prepared on the host as in the reference.  ``free`` = the non-Dirichlet nodes in the order of
    not_dirichlet_inds_g2l: the vectors ``solve`` takes and returns are indexed like the reference's (:1921-1933)."""

    def __init__(self, S, chol_A_GG, E, free):
        self.S = S
        sd = S.sd
        S.h.upload_lorasc(chol_A_GG[0], chol_A_GG[1], E)
        g2l = np.full(sd.nnode, -1, dtype=np.int64)
        g2l[np.asarray(free)] = np.arange(len(free))
        self.idx_I = g2l[np.concatenate(sd.node_Id)] if sd.ndom else np.zeros(0, np.int64)   # x_Id[idom], concatenated
        self.idx_G = g2l[sd.node_gamma]
        self.n = len(free)

    def solve(self, x):        # Πlorasc \ x  (:1980-1982)
        u_I, u_G = self.S.h.apply_lorasc(x[self.idx_I], x[self.idx_G])
        u = np.empty(self.n)
        u[self.idx_I] = u_I
        u[self.idx_G] = u_G
        return u

    def ldiv_into(self, z, r):  # ldiv!(z, Πlorasc, r)  (:1984-1988)
        z[...] = self.solve(r)
        return z

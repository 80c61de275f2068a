# Replay kit, part 2 (to be run INSIDE a checkout of venkovic/julia-phd-krylov-spdes, Julia >= 1.5):
#
#     julia replay_fixture.jl <fixture dir> <N> <ndom>
#
# Runs the reference's OWN functions (Fem, RecyclingKrylovSolvers) on the inputs written by
# scripts/export_fixture.py and compares with the answers of this repository's oracle:
#   * set_subdomains (EllipticPdeDomainDecomposition.jl:86-193): node_Γ and node_Γ_cnt bit for bit;
#   * prepare_local_schurs (:389-546) + get_schur_rhs (:835-864) with exact interior solves
#     (Cholesky factors as `preconds`, so the inner IterativeSolvers.cg converges at once);
#   * apply_local_schurs (:711-747), the Neumann-Neumann preconditioner (:1201-1220, :1361-1403);
#   * pcg (cg.jl:67-109): iteration count, residual history, solution.
# Not executable in this repository's build image (no Julia): it follows Example03 / Example07 of the
# reference line for line (Example03_EllipticPdeDomainDecomposition.jl:34-150, :193).
push!(LOAD_PATH, "./Utils/")
push!(LOAD_PATH, "./Fem/")
push!(LOAD_PATH, "./RecyclingKrylovSolvers/")
import Pkg
Pkg.activate(".")
using Fem
using RecyclingKrylovSolvers
using LinearMaps
using LinearAlgebra
using SparseArrays
using NPZ

dir, N, ndom = ARGS[1], parse(Int, ARGS[2]), parse(Int, ARGS[3])
rd(name) = npzread(joinpath(dir, "data", "DoF$N-ndom$ndom.$name.npz"))

# the reference's loaders read ./data/DoF{N}.*: run them from the fixture directory
cells, points, point_markers, cell_neighbors, epart, npart = cd(dir) do
  c, p, m, nb = load_mesh(N)
  e, n = load_partition(N, ndom)
  c, p, m, nb, e, n
end

dirichlet_inds_g2l, not_dirichlet_inds_g2l,
dirichlet_inds_l2g, not_dirichlet_inds_l2g = get_dirichlet_inds(points, point_markers)

ind_Id_g2l, ind_Γd_g2l, ind_Γ_g2l, ind_Γd_Γ2l, node_owner,
elemd, node_Γ, node_Γ_cnt, node_Id, nnode_Id = set_subdomains(cells, cell_neighbors, epart, npart,
                                                              dirichlet_inds_g2l)

report(name, ok, detail="") = println(rpad(name, 44), ok ? "OK   " : "FAIL ", detail)
relerr(a, b) = norm(a - b) / norm(b)

report("node_Γ (bit for bit)", node_Γ == vec(rd("expected.node_gamma")))
report("node_Γ_cnt (bit for bit)", node_Γ_cnt == vec(rd("expected.node_gamma_cnt")))

coeff = vec(rd("coeff"))
f(x::Float64, y::Float64) = -1.                  # Example07:65-67
uexact(x::Float64, y::Float64) = 0.734           # Example07:69-71

A_IIdd, A_IΓdd, A_ΓΓdd, b_Idd, b_Γ = prepare_local_schurs(cells, points, epart, ind_Id_g2l, ind_Γd_g2l,
                                                          ind_Γ_g2l, node_owner, coeff, f, uexact)
Π_IId = [cholesky(A_IIdd[idom]) for idom in 1:ndom]          # exact interior solves (cf. EPDD.jl:1536)

b_schur = get_schur_rhs(b_Idd, A_IIdd, A_IΓdd, b_Γ, ind_Γd_Γ2l, preconds=Π_IId)
e = relerr(b_schur, vec(rd("expected.b_schur")))
report("get_schur_rhs", e <= 1e-10, "rel err $e")

n_Γ = length(b_Γ)
S = LinearMap(x -> apply_local_schurs(A_IIdd, A_IΓdd, A_ΓΓdd, ind_Γd_Γ2l, node_Γ_cnt, x, preconds=Π_IId),
              nothing, n_Γ, issymmetric=true)
x_probe = vec(rd("expected.x_probe"))
e = relerr(S * x_probe, vec(rd("expected.S_x")))
report("apply_local_schurs", e <= 1e-8, "rel err $e   (inner cg reltol 1e-9)")

# assemble_local_schurs has its `preconds` branches swapped (EPDD.jl:676-691, SURVEY.md section 3.5): call it
# the way Example03:129 does; the dense S_d then come from plain inner CG at reltol 1e-9
Sd_local_mat = assemble_local_schurs(A_IIdd, A_IΓdd, A_ΓΓdd, preconds=Π_IId)
ΠSnn = prepare_neumann_neumann_schur_precond(Sd_local_mat, ind_Γd_Γ2l, node_Γ_cnt)
e = relerr(ΠSnn \ x_probe, vec(rd("expected.M_x")))
report("NeumannNeumannSchurPreconditioner \\", e <= 1e-5, "rel err $e   (pinv of S_d known to ~1e-9)")

S_mat = LinearMap(x -> apply_local_schurs(Sd_local_mat, ind_Γd_Γ2l, node_Γ_cnt, x), nothing, n_Γ, issymmetric=true)
u_Γ, it, res = pcg(S_mat, b_schur, zeros(n_Γ), ΠSnn)
it_o = Int(rd("expected.pcg_it")[1])
res_o = vec(rd("expected.pcg_res"))
k = min(it, it_o)
report("pcg iteration count (±1)", abs(it - it_o) <= 1, "reference $it, oracle $it_o")
report("pcg residual history", maximum(abs.(res[1:k] - res_o[1:k])) <= 1e-6 * maximum(res_o), "")
e = relerr(u_Γ, vec(rd("expected.pcg_x")))
report("pcg solution", e <= 1e-6, "rel err $e   (both stop at a 1e-7 residual)")

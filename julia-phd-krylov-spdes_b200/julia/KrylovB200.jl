# KrylovB200.jl -- ccall shim that puts libkrylov_b200.so behind the reference's own API.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI (the build image has no Julia).  Every ccall below has a
# one-to-one, tested twin in ../capi.py (ctypes); the argument order, types and index conventions
# are the ones declared in include/krylov_b200.h.
#
# Usage inside the reference (e.g. Example07_PcgSchurStochasticEllipticPde.jl, after
# prepare_local_schurs / set_subdomains):
#
#   push!(LOAD_PATH, "path/to/julia-phd-krylov-spdes_b200/julia"); using KrylovB200
#   F  = [cholesky(A_IIdd[d]) for d in 1:ndom]                       # host, CHOLMOD (as :1536)
#   S  = B200SchurOperator(A_IΓdd, A_ΓΓdd, F, ind_Γd_Γ2l, node_Γ_cnt)  # replaces the LinearMap at Example07:128-136
#   M  = B200NeumannNeumann(S, ΠSnn_0.ΠSd)                             # replaces ΠSnn_0 (Example07:152-154)
#   u_Γ, it, res = pcg(S, b_schur, zeros(S.N), M)                      # same call as Example07:273
#
# New types are introduced (no method piracy on LinearMaps.FunctionMap); the solver methods dispatch
# on them and run the whole solve on the GPU.
module KrylovB200

using SparseArrays, LinearAlgebra, SuiteSparse
import Base: *, \, size
import LinearAlgebra: mul!, ldiv!

export B200SchurOperator, B200NeumannNeumann, pcg, cg, get_schur_rhs, get_subdomain_solutions,
       assemble_local_schurs!

const LIB = get(ENV, "KRYLOV_B200_LIB", joinpath(@__DIR__, "..", "libkrylov_b200.so"))

struct KB200Error <: Exception
  code::Cint
  msg::String
end
function chk(rc::Cint)
  rc == 0 && return
  throw(KB200Error(rc, unsafe_string(ccall((:kb200_last_error, LIB), Cstring, ()))))
end

mutable struct Timings            # kb200_timings
  solve_ms::Cdouble; schur_ms::Cdouble; nn_ms::Cdouble
  schur_launches::Int64; nn_launches::Int64; kernel_launches::Int64; h2d_bytes::Int64; d2h_bytes::Int64
  Timings() = new(0, 0, 0, 0, 0, 0, 0, 0)
end

mutable struct B200SchurOperator
  h::Ptr{Cvoid}
  N::Int                          # = n_Γ, as S.N of the reference's FunctionMap (cg.jl:76)
  n_I::Vector{Int64}
  assembled::Bool
end
size(S::B200SchurOperator) = (S.N, S.N)

# Flatten ind_Γd_Γ2l[d] (Dict Γ-index => local index) exactly like ind_Γd_l2Γ at
# Fem/EllipticPdeDomainDecomposition.jl:1090-1095.
function flatten_l2Γ(ind_Γd_Γ2l::Dict{Int,Int})
  out = Vector{Int64}(undef, length(ind_Γd_Γ2l))
  for (lnode_in_Γ, lnode_in_Γd) in ind_Γd_Γ2l
    out[lnode_in_Γd] = lnode_in_Γ
  end
  return out
end

"""
    B200SchurOperator(A_IΓdd, A_ΓΓdd, F, ind_Γd_Γ2l, node_Γ_cnt; device=0, csize=4)

`F[d]` is the CHOLMOD factor of `A_IIdd[d]` (`cholesky(A_IIdd[d])`): `F[d].p` and `sparse(F[d].L)`
are uploaded once; the device applies `A_II,d⁻¹` with its own level-scheduled triangular solves.
`devices = [0, 1, ..., 7]` spreads the subdomains over the GPUs of the box from this one process.
"""
function B200SchurOperator(A_IΓdd::Vector{SparseMatrixCSC{Float64,Int}},
                           A_ΓΓdd::Vector{SparseMatrixCSC{Float64,Int}},
                           F::Vector, ind_Γd_Γ2l::Vector{Dict{Int,Int}}, node_Γ_cnt::Vector{Int};
                           device::Int=0, csize::Int=4, devices::Union{Vector{Int},Nothing}=nothing)
  ndom = length(A_IΓdd)
  hr = Ref{Ptr{Cvoid}}(C_NULL)
  if devices === nothing
    chk(ccall((:kb200_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, hr))
  else
    # ONE Julia process, several GPUs: the library shards the subdomains over the devices and drives each with a
    # host thread of its own; every call below then acts on all of them (kb200_create_multi)
    ids = Vector{Cint}(devices)
    chk(ccall((:kb200_create_multi, LIB), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), length(ids), ids, hr))
  end
  h = hr[]
  chk(ccall((:kb200_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), h, "csize", csize))
  n_I = Int64[A_IΓdd[d].m for d in 1:ndom]
  n_gd = Int64[A_IΓdd[d].n for d in 1:ndom]
  maps = [flatten_l2Γ(ind_Γd_Γ2l[d]) for d in 1:ndom]          # 1-based, as the ABI expects
  ptrs = [pointer(m) for m in maps]
  cnt = Vector{Int64}(node_Γ_cnt)
  GC.@preserve maps begin
    chk(ccall((:kb200_set_topology, LIB), Cint,
              (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Ptr{Int64}}, Ptr{Int64}),
              h, ndom, length(cnt), n_I, n_gd, ptrs, cnt))
  end
  for d in 1:ndom
    G, H = A_IΓdd[d], A_ΓΓdd[d]
    chk(ccall((:kb200_upload_local_blocks, LIB), Cint,
              (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
              h, d - 1, G.colptr, G.rowval, G.nzval, H.colptr, H.rowval, H.nzval))
    L = sparse(F[d].L)                                        # simplicial lower factor, L*L' = A[p,p]
    p = Vector{Int64}(F[d].p)
    chk(ccall((:kb200_upload_cholesky, LIB), Cint,
              (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
              h, d - 1, p, L.colptr, L.rowval, L.nzval))
  end
  chk(ccall((:kb200_finalize, LIB), Cint, (Ptr{Cvoid},), h))
  S = B200SchurOperator(h, length(cnt), n_I, false)
  finalizer(s -> ccall((:kb200_destroy, LIB), Cint, (Ptr{Cvoid},), s.h), S)
  return S
end

# S * x and mul!(y, S, x): what cg.jl:83 and cg.jl:93 call on the reference's FunctionMap
# (apply_local_schurs, EllipticPdeDomainDecomposition.jl:711-747 / :761-785).
function mul!(y::Vector{Float64}, S::B200SchurOperator, x::Vector{Float64})
  if S.assembled
    chk(ccall((:kb200_apply_schur_assembled, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), S.h, x, y, 1))
  else
    chk(ccall((:kb200_apply_schur, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), S.h, x, y, 1))
  end
  return y
end
*(S::B200SchurOperator, x::Vector{Float64}) = mul!(similar(x), S, x)

"""assemble_local_schurs (:667-695) on the device; afterwards `S.assembled = true` selects the dense path."""
function assemble_local_schurs!(S::B200SchurOperator)
  chk(ccall((:kb200_assemble_local_schurs, LIB), Cint, (Ptr{Cvoid},), S.h))
  return S
end

struct B200NeumannNeumann
  S::B200SchurOperator
end
"""NeumannNeumannSchurPreconditioner (:1111-1137): uploads ΠSd[d] (column-major Matrix{Float64})."""
function B200NeumannNeumann(S::B200SchurOperator, ΠSd::Vector{Matrix{Float64}})
  for d in 1:length(ΠSd)
    chk(ccall((:kb200_upload_nn_pinv, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), S.h, d - 1, ΠSd[d]))
  end
  return B200NeumannNeumann(S)
end
# `\`, ldiv!(z, M, r), ldiv!(M, r): EllipticPdeDomainDecomposition.jl:1389-1403
function ldiv!(z::Vector{Float64}, M::B200NeumannNeumann, r::Vector{Float64})
  chk(ccall((:kb200_apply_nn, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), M.S.h, r, z, 1))
  return z
end
\(M::B200NeumannNeumann, r::Vector{Float64}) = ldiv!(similar(r), M, r)
ldiv!(M::B200NeumannNeumann, r::Vector{Float64}) = (r .= M \ copy(r))

# get_schur_rhs (:835-864) and get_subdomain_solutions (:1014-1025)
function get_schur_rhs(S::B200SchurOperator, b_Id::Vector{Vector{Float64}}, b_Γ::Vector{Float64})
  out = similar(b_Γ)
  chk(ccall((:kb200_schur_rhs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            S.h, vcat(b_Id...), b_Γ, out))
  return out
end
function get_subdomain_solutions(S::B200SchurOperator, u_Γ::Vector{Float64}, b_Id::Vector{Vector{Float64}})
  u = Vector{Float64}(undef, sum(S.n_I))
  chk(ccall((:kb200_back_substitute, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            S.h, u_Γ, vcat(b_Id...), u))
  off = cumsum(vcat(0, S.n_I))
  return [u[off[d]+1:off[d+1]] for d in 1:length(S.n_I)]
end

const eps_rtol = 1e-7   # RecyclingKrylovSolvers.jl:21

"""
    pcg(S, b, x, M; maxit=0)   ->  (x, it, res_norm)      # cg.jl:67-109, whole solve on the GPU

Same positional signature and return tuple as RecyclingKrylovSolvers.pcg; `x` is updated in place.
"""
function pcg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64}, M::Union{B200NeumannNeumann,Nothing};
             maxit::Int=0, timings::Union{Timings,Nothing}=nothing)
  cap = (maxit == 0 ? S.N : maxit) + 1
  res = Vector{Float64}(undef, cap)
  it = Ref{Int64}(0)
  tm = timings === nothing ? Timings() : timings
  chk(ccall((:kb200_pcg, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Cint, Cint, Ref{Int64}, Ptr{Float64}, Ref{Timings}),
            S.h, b, x, maxit, eps_rtol, S.assembled ? 1 : 0, M === nothing ? 0 : 1, it, res, tm))
  return x, Int(it[]), res[1:it[]]
end
cg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64}; maxit::Int=0) = pcg(S, b, x, nothing; maxit=maxit)


# ---------------------------------------------------------------------------------------------------
# Deflated / recycling solvers (defcg.jl:242-308, eigcg.jl:143-267, defcg.jl:337-473): ONE ccall each.
# The whole solve runs inside libkrylov_b200.so -- device bursts between the restarts, AV = A V with the
# multi-vector stream kernel, V'AV and V (Q Z) on FP64 tensor cores, the spdim-sized eigen / svd / rank algebra
# of eigcg.jl:222-227 with cuSOLVER.  Signatures and return tuples are RecyclingKrylovSolvers' own.
# ---------------------------------------------------------------------------------------------------
"""
    defpcg(S, b, x, W, M; maxit=0) -> (x, it, res_norm)          # defcg.jl:242-308
"""
function defpcg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64}, W::Matrix{Float64},
                M::Union{B200NeumannNeumann,Nothing}; maxit::Int=0)
  cap = (maxit == 0 ? S.N : maxit) + 1
  res = Vector{Float64}(undef, cap); it = Ref{Int64}(0); tm = Timings()
  chk(ccall((:kb200_defpcg, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Float64, Cint, Cint, Ref{Int64}, Ptr{Float64}, Ref{Timings}),
            S.h, b, x, W, size(W, 2), maxit, eps_rtol, S.assembled ? 1 : 0, M === nothing ? 0 : 1, it, res, tm))
  return x, Int(it[]), res[1:it[]]
end

"""
    eigpcg(S, b, x, M, nvec, spdim; maxit=0) -> (x, it, res_norm, V[:, 1:nvec])      # eigcg.jl:143-267
"""
function eigpcg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64}, M::Union{B200NeumannNeumann,Nothing},
                nvec::Int, spdim::Int; maxit::Int=0)
  cap = (maxit == 0 ? S.N : maxit) + 1
  res = Vector{Float64}(undef, cap); it = Ref{Int64}(0); tm = Timings()
  V = Matrix{Float64}(undef, S.N, nvec)
  chk(ccall((:kb200_eigpcg, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Float64, Cint, Cint, Ref{Int64}, Ptr{Float64}, Ptr{Float64}, Ref{Timings}),
            S.h, b, x, nvec, spdim, maxit, eps_rtol, S.assembled ? 1 : 0, M === nothing ? 0 : 1, it, res, V, tm))
  return x, Int(it[]), res[1:it[]], V
end

"""
    eigdefpcg(S, b, x, M, W, spdim; maxit=0) -> (x, it, res_norm, V[:, 1:nvec])      # defcg.jl:337-473
"""
function eigdefpcg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64}, M::Union{B200NeumannNeumann,Nothing},
                   W::Matrix{Float64}, spdim::Int; maxit::Int=0)
  nvec = size(W, 2)
  cap = (maxit == 0 ? S.N : maxit) + 1
  res = Vector{Float64}(undef, cap); it = Ref{Int64}(0); tm = Timings()
  V = Matrix{Float64}(undef, S.N, nvec)
  chk(ccall((:kb200_eigdefpcg, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Float64, Cint, Cint, Ref{Int64}, Ptr{Float64}, Ptr{Float64}, Ref{Timings}),
            S.h, b, x, W, nvec, spdim, maxit, eps_rtol, S.assembled ? 1 : 0, M === nothing ? 0 : 1, it, res, V, tm))
  return x, Int(it[]), res[1:it[]], V
end

"""
    pcg(S, B::Matrix, X::Matrix, M; maxit=0) -> (X, its, res_norms)

Independent recurrences for the columns of B (the loop over right-hand sides of
Example11_EigInitPcgEllipticPdeMultipleRhs_Functions.jl:140-179) sharing the operator launches.
"""
function pcg(S::B200SchurOperator, B::Matrix{Float64}, X::Matrix{Float64}, M::Union{B200NeumannNeumann,Nothing}; maxit::Int=0)
  nrhs = size(B, 2)
  stride = (maxit == 0 ? S.N : maxit) + 1
  its = zeros(Int64, nrhs); res = zeros(Float64, stride, nrhs)
  chk(ccall((:kb200_pcg_multi, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Float64, Cint, Cint, Ptr{Int64}, Ptr{Float64}, Int64),
            S.h, B, X, nrhs, maxit, eps_rtol, S.assembled ? 1 : 0, M === nothing ? 0 : 1, its, res, stride))
  return X, its, [res[1:its[k], k] for k in 1:nrhs]
end

"""
    update_realization!(S, A_IΓdd, A_ΓΓdd, F)

New coefficient realization on the same mesh / partition / ordering (the loop of Example07:160-204): values only.
The symbolic plan on the GPU is kept; the resident stream is rewritten on the device before the next operation.
`F[d]` must be a factorization with the SAME permutation and pattern (e.g. `cholesky!(F[d], A_IIdd[d])`).
"""
function update_realization!(S::B200SchurOperator, A_IΓdd, A_ΓΓdd, F)
  for d in 1:length(F)
    L = sparse(F[d].L)
    chk(ccall((:kb200_update_local_block_values, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
              S.h, d - 1, A_IΓdd[d].nzval, A_ΓΓdd[d].nzval))
    chk(ccall((:kb200_update_cholesky_values, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), S.h, d - 1, L.nzval))
  end
  S.assembled = false     # dense S_d of the previous realization are dropped by the library
  return S
end

"""
    enable_device_refactorization!(S, A_IIdd)
    refactor_realization!(S, A_IIdd, A_IΓdd, A_ΓΓdd)

Numeric re-factorisation on the GPU: the permutation and pattern of the factors uploaded at construction are the symbolic
factorisation; a new realization sends the values of `A_IIdd[d]` (7 per row) instead of a host `cholesky()` and the values
of `sparse(F.L)` (33 per row).
"""
function enable_device_refactorization!(S::B200SchurOperator, A_IIdd)
  for d in 1:length(A_IIdd)
    A = A_IIdd[d]
    chk(ccall((:kb200_upload_interior_pattern, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}),
              S.h, d - 1, Vector{Int64}(A.colptr), Vector{Int64}(A.rowval)))
  end
  return S
end
function refactor_realization!(S::B200SchurOperator, A_IIdd, A_IΓdd, A_ΓΓdd)
  for d in 1:length(A_IIdd)
    chk(ccall((:kb200_update_local_block_values, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
              S.h, d - 1, A_IΓdd[d].nzval, A_ΓΓdd[d].nzval))
    chk(ccall((:kb200_refactor_interior, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), S.h, d - 1, A_IIdd[d].nzval))
  end
  S.assembled = false
  return S
end

"""
    enable_device_assembly!(S, cells, points, epart, node_owner, ind_Id_g2l, ind_Γd_g2l, A_IIdd, A_IΓdd, A_ΓΓdd)
    assemble_realization!(S, coeff)

The element loop of prepare_local_schurs (EllipticPdeDomainDecomposition.jl:430-497) on the GPU: after the one-time
set-up a new realization is ONE vector, the nodal coefficient `coeff` (the `Array{Float64,1}` branch of :440-441); the
blocks are re-assembled with the bits of `sparse(I, J, V)` (:540-543), the numeric Cholesky and the stream refresh follow
on the device.  The set-up replays the routing of the loop nest (:459-497) once to tell the library, for every stored
entry of A_IIdd[d] / A_IΓdd[d] / A_ΓΓdd[d], which contributions `9 (iel - 1) + 3 (i - 1) + j` it is the sum of.
The right-hand sides (b_Id, b_Γ) remain the caller's (only their Dirichlet-lift terms depend on the coefficient).
"""
function enable_device_assembly!(S::B200SchurOperator, cells::Array{Int,2}, points::Array{Float64,2}, epart::Array{Int,1},
                                 node_owner::Array{Int,1}, ind_Id_g2l::Vector{Dict{Int,Int}}, ind_Γd_g2l::Vector{Dict{Int,Int}},
                                 A_IIdd, A_IΓdd, A_ΓΓdd)
  enable_device_refactorization!(S, A_IIdd)
  ndom = length(A_IIdd)
  nel = size(cells, 2)
  chk(ccall((:kb200_upload_assembly_mesh, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Float64}),
            S.h, size(points, 2), nel, cells, points))
  # position of entry (r, c) in the nzval array of a SparseMatrixCSC (rows of a column are sorted)
  pos(A, r, c) = (rg = A.colptr[c]:A.colptr[c+1]-1; rg[searchsortedfirst(view(A.rowval, rg), r)])
  lists = [[[Int64[] for _ in 1:length(A[d].nzval)] for A in (A_IIdd, A_IΓdd, A_ΓΓdd)] for d in 1:ndom]
  for iel in 1:nel                                            # the traversal of :430-497
    idom = epart[iel]
    for (i, inode) in enumerate(cells[:, iel]), (j, jnode) in enumerate(cells[:, iel])
      io, jo = node_owner[inode], node_owner[jnode]
      (io == 0 || jo == 0) && continue
      c = 9 * (iel - 1) + 3 * (i - 1) + j
      if io == -1 && jo == -1
        push!(lists[idom][3][pos(A_ΓΓdd[idom], ind_Γd_g2l[idom][inode], ind_Γd_g2l[idom][jnode])], c)
      elseif io > 0 && jo > 0
        push!(lists[idom][1][pos(A_IIdd[idom], ind_Id_g2l[idom][inode], ind_Id_g2l[idom][jnode])], c)
      elseif io > 0 && jo == -1
        push!(lists[idom][2][pos(A_IΓdd[idom], ind_Id_g2l[idom][inode], ind_Γd_g2l[idom][jnode])], c)
      end
    end
  end
  for d in 1:ndom, which in 1:3
    l = lists[d][which]
    ptr = Int64[1]; for v in l; push!(ptr, ptr[end] + length(v)); end
    src = isempty(l) ? Int64[] : reduce(vcat, l)
    chk(ccall((:kb200_upload_assembly_map, LIB), Cint, (Ptr{Cvoid}, Int64, Cint, Int64, Ptr{Int64}, Ptr{Int64}),
              S.h, d - 1, which - 1, length(l), ptr, src))
  end
  return S
end
"""
    enable_device_rhs!(S, cells, points, epart, node_owner, ind_Id_g2l, ind_Γ_g2l, b_Id, b_Γ, f, uexact)

Right-hand sides of prepare_local_schurs (:499-536) on the GPU as well: `b_Id`, `b_Γ` as returned for ANY coefficient give
the load vector; the entries of nodes that share an element with a Dirichlet node are re-assembled per realization from
their ordered terms (`b -= ΔKij * uexact(...)` at :502/:506, `b += Δb` at :529/:533).  Afterwards
`get_schur_rhs(S)` and `get_subdomain_solutions(S, u_Γ)` need no right-hand-side arguments.
"""
function enable_device_rhs!(S::B200SchurOperator, cells::Array{Int,2}, points::Array{Float64,2}, epart::Array{Int,1},
                            node_owner::Array{Int,1}, ind_Id_g2l::Vector{Dict{Int,Int}}, ind_Γ_g2l::Dict{Int,Int},
                            b_Id::Vector{Vector{Float64}}, b_Γ::Vector{Float64}, f::Function, uexact::Function)
  nel = size(cells, 2)
  off = cumsum(vcat(0, S.n_I)); nItot = off[end]
  index(idom, node) = node_owner[node] == -1 ? nItot + ind_Γ_g2l[node] : off[idom] + ind_Id_g2l[idom][node]
  terms = Dict{Int,Vector{Tuple{Int64,Float64}}}()              # target index => ordered (src, val)
  touched = falses(length(node_owner))                          # nodes with a lift term
  for iel in 1:nel, inode in cells[:, iel]
    node_owner[inode] == 0 && (touched[cells[:, iel]] .= true)
  end
  for iel in 1:nel                                              # the traversal of :430-536
    nodes = cells[:, iel]
    any(touched[nodes]) || continue
    idom = epart[iel]
    x = points[1, nodes]; y = points[2, nodes]
    for (i, inode) in enumerate(nodes), (j, jnode) in enumerate(nodes)
      (node_owner[inode] == 0 && node_owner[jnode] != 0 && touched[jnode]) || continue
      push!(get!(terms, index(idom, jnode), Tuple{Int64,Float64}[]), (9 * (iel - 1) + 3 * (i - 1) + j, uexact(x[i], y[i])))
    end
    Δx = (x[3] - x[2], x[1] - x[3], x[2] - x[1]); Δy = (y[2] - y[3], y[3] - y[1], y[1] - y[2])
    Area = (Δx[3] * Δy[2] - Δx[2] * Δy[3]) / 2.
    for (i, inode) in enumerate(nodes)
      (node_owner[inode] != 0 && touched[inode]) || continue
      j = i % 3 + 1; k = (i + 1) % 3 + 1
      Δb = (2 * f(x[i], y[i]) + f(x[j], y[j]) + f(x[k], y[k])) * Area / 12
      push!(get!(terms, index(idom, inode), Tuple{Int64,Float64}[]), (0, Δb))
    end
  end
  tgt = sort(collect(keys(terms)))
  ptr = Int64[1]; src = Int64[]; val = Float64[]
  for t in tgt
    for (s, v) in terms[t]; push!(src, s); push!(val, v); end
    push!(ptr, length(src) + 1)
  end
  chk(ccall((:kb200_upload_rhs_map, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
            S.h, vcat(b_Id...), b_Γ, length(tgt), Vector{Int64}(tgt), ptr, src, val))
  return S
end
# get_schur_rhs / get_subdomain_solutions on the right-hand sides the device assembled
function get_schur_rhs(S::B200SchurOperator)
  out = Vector{Float64}(undef, S.N)
  chk(ccall((:kb200_schur_rhs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), S.h, C_NULL, C_NULL, out))
  return out
end
function get_subdomain_solutions(S::B200SchurOperator, u_Γ::Vector{Float64})
  u = Vector{Float64}(undef, sum(S.n_I))
  chk(ccall((:kb200_back_substitute, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), S.h, u_Γ, C_NULL, u))
  off = cumsum(vcat(0, S.n_I))
  return [u[off[d]+1:off[d+1]] for d in 1:length(S.n_I)]
end
function assemble_realization!(S::B200SchurOperator, coeff::Vector{Float64})
  chk(ccall((:kb200_assemble_realization, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), S.h, coeff))
  S.assembled = false
  return S
end

# ---------------------------------------------------------------------------------------------------
# Post-hoc recycling variants (rrdefpcg.jl:48-151, :204-275; hrdefpcg.jl:48-164, :217-295): ONE ccall each
# ---------------------------------------------------------------------------------------------------
function posthoc(kind::Int, S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64}, W::Union{Matrix{Float64},Nothing},
                 M::Union{B200NeumannNeumann,Nothing}, nvec::Int, spdim::Int, maxit::Int)
  cap = (maxit == 0 ? S.N : maxit) + 1
  res = Vector{Float64}(undef, cap); it = Ref{Int64}(0); tm = Timings()
  W2 = Matrix{Float64}(undef, S.N, nvec)
  chk(ccall((:kb200_posthoc_recycle, LIB), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Float64, Cint, Cint, Ref{Int64},
             Ptr{Float64}, Ptr{Float64}, Ref{Timings}),
            S.h, kind, b, x, W === nothing ? C_NULL : W, nvec, spdim, maxit, eps_rtol, S.assembled ? 1 : 0, M === nothing ? 0 : 1,
            it, res, W2, tm))
  return x, Int(it[]), res[1:it[]], W2
end
rrpcg(S::B200SchurOperator, b, x, M, nvec::Int, spdim::Int; maxit::Int=0) = posthoc(0, S, b, x, nothing, M, nvec, spdim, maxit)
rrdefpcg(S::B200SchurOperator, b, x, W::Matrix{Float64}, M, spdim::Int; maxit::Int=0) = posthoc(0, S, b, x, W, M, size(W, 2), spdim, maxit)
hrpcg(S::B200SchurOperator, b, x, M, nvec::Int, spdim::Int; maxit::Int=0) = posthoc(1, S, b, x, nothing, M, nvec, spdim, maxit)
hrdefpcg(S::B200SchurOperator, b, x, W::Matrix{Float64}, M, spdim::Int; maxit::Int=0) = posthoc(1, S, b, x, W, M, size(W, 2), spdim, maxit)
# thick-restart siblings (trrrdefpcg.jl:48-199, :234-355; trhrdefpcg.jl:48-208, :247-389): kinds 2 / 3 of the same entry point
trrrpcg(S::B200SchurOperator, b, x, M, nvec::Int, spdim::Int; maxit::Int=0) = posthoc(2, S, b, x, nothing, M, nvec, spdim, maxit)
trrrdefpcg(S::B200SchurOperator, b, x, W::Matrix{Float64}, M, spdim::Int; maxit::Int=0) = posthoc(2, S, b, x, W, M, size(W, 2), spdim, maxit)
trhrpcg(S::B200SchurOperator, b, x, M, nvec::Int, spdim::Int; maxit::Int=0) = posthoc(3, S, b, x, nothing, M, nvec, spdim, maxit)
trhrdefpcg(S::B200SchurOperator, b, x, W::Matrix{Float64}, M, spdim::Int; maxit::Int=0) = posthoc(3, S, b, x, W, M, size(W, 2), spdim, maxit)
# locally-optimal thick-restart siblings (lotrrrdefpcg.jl:48-214, :252-389; lotrhrdefpcg.jl:48-229, :267-428): kinds 4 / 5
lotrrrpcg(S::B200SchurOperator, b, x, M, nvec::Int, spdim::Int; maxit::Int=0) = posthoc(4, S, b, x, nothing, M, nvec, spdim, maxit)
lotrrrdefpcg(S::B200SchurOperator, b, x, W::Matrix{Float64}, M, spdim::Int; maxit::Int=0) = posthoc(4, S, b, x, W, M, size(W, 2), spdim, maxit)
lotrhrpcg(S::B200SchurOperator, b, x, M, nvec::Int, spdim::Int; maxit::Int=0) = posthoc(5, S, b, x, nothing, M, nvec, spdim, maxit)
lotrhrdefpcg(S::B200SchurOperator, b, x, W::Matrix{Float64}, M, spdim::Int; maxit::Int=0) = posthoc(5, S, b, x, W, M, size(W, 2), spdim, maxit)

# ---------------------------------------------------------------------------------------------------
# LORASC preconditioner of the full system (EllipticPdeDomainDecomposition.jl:1406-1428, apply :1908-1976)
# ---------------------------------------------------------------------------------------------------
struct B200Lorasc
  S::B200SchurOperator
  idx_I::Vector{Int}      # not_dirichlet index of every interior node, subdomain after subdomain (x_Id[idom], :1921-1926)
  idx_Γ::Vector{Int}      # not_dirichlet index of every interface node (:1928-1933)
end
"""
    B200Lorasc(S, Πlorasc)        # Πlorasc::LorascPreconditioner from prepare_lorasc_precond (:1502-1676), host side

Uploads `chol_A_ΓΓ` (`F.p`, `sparse(F.L)`) and the correction vectors `E`; the interior factors and `A_IΓ,d` are the
ones `S` already holds.  `M \ x` / `ldiv!(z, M, x)` then run apply_lorasc on the GPU(s) of `S`.
"""
function B200Lorasc(S::B200SchurOperator, Π)
  F = Π.chol_A_ΓΓ
  L = sparse(F.L); p = Vector{Int64}(F.p)
  nev = length(Π.E)
  E = nev == 0 ? zeros(Float64, S.N, 0) : hcat(Π.E...)
  chk(ccall((:kb200_upload_lorasc, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Int64),
            S.h, p, L.colptr, L.rowval, L.nzval, E, nev))
  ndom = length(Π.ind_Id_g2l)
  idx_I = Int[]
  for idom in 1:ndom
    loc = Vector{Int}(undef, Π.ind_Id_g2l[idom].count)
    for (node, node_in_I) in Π.ind_Id_g2l[idom]
      loc[node_in_I] = Π.not_dirichlet_inds_g2l[node]
    end
    append!(idx_I, loc)
  end
  idx_Γ = Vector{Int}(undef, Π.ind_Γ_g2l.count)
  for (node, node_in_Γ) in Π.ind_Γ_g2l
    idx_Γ[node_in_Γ] = Π.not_dirichlet_inds_g2l[node]
  end
  return B200Lorasc(S, idx_I, idx_Γ)
end
function ldiv!(z::Vector{Float64}, M::B200Lorasc, x::Vector{Float64})
  xI = x[M.idx_I]; xΓ = x[M.idx_Γ]
  uI = similar(xI); uΓ = similar(xΓ)
  chk(ccall((:kb200_apply_lorasc, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            M.S.h, xI, xΓ, uI, uΓ))
  z[M.idx_I] = uI; z[M.idx_Γ] = uΓ
  return z
end
\(M::B200Lorasc, x::Vector{Float64}) = ldiv!(similar(x), M, x)
ldiv!(M::B200Lorasc, x::Vector{Float64}) = (x .= M \ copy(x))

export defpcg, eigpcg, eigdefpcg, update_realization!, enable_device_refactorization!, refactor_realization!, enable_device_assembly!, enable_device_rhs!, assemble_realization!, rrpcg, rrdefpcg, hrpcg, hrdefpcg, trrrpcg, trrrdefpcg, trhrpcg, trhrdefpcg, lotrrrpcg, lotrrrdefpcg, lotrhrpcg, lotrhrdefpcg, B200Lorasc

end # module

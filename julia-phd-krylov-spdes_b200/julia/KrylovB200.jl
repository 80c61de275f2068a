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
"""
function B200SchurOperator(A_IΓdd::Vector{SparseMatrixCSC{Float64,Int}},
                           A_ΓΓdd::Vector{SparseMatrixCSC{Float64,Int}},
                           F::Vector, ind_Γd_Γ2l::Vector{Dict{Int,Int}}, node_Γ_cnt::Vector{Int};
                           device::Int=0, csize::Int=4)
  ndom = length(A_IΓdd)
  hr = Ref{Ptr{Cvoid}}(C_NULL)
  chk(ccall((:kb200_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, hr))
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
  f = S.assembled ? :kb200_apply_schur_assembled : :kb200_apply_schur
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
# Deflated / recycling solvers (defcg.jl:242-308, eigcg.jl:143-267, defcg.jl:337-473).
# The reference's loops are kept; every n_Γ-sized array becomes a column of a device block
# (kb200_mat_alloc) and the BLAS-1/2 calls become kb200_op_* calls; eigpcg / eigdefpcg run the
# iterations between two restarts as device bursts (kb200_pcg_burst) and keep the reference's own LAPACK
# lines for the Rayleigh-Ritz restarts (eigvecs / rank / svd / eigen, eigcg.jl:220-228) on the small host
# matrix VtAV.  `defpcg` here is the host-driven pattern; kb200_defpcg (one ccall) runs it on the device.
# ---------------------------------------------------------------------------------------------------
module DeviceOps
using ..KrylovB200: LIB, chk, B200SchurOperator
mat_alloc(S, ncols) = (id = Ref{Int64}(0); chk(ccall((:kb200_mat_alloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Int64}), S.h, ncols, id)); id[])
mat_free(S, id) = chk(ccall((:kb200_mat_free, LIB), Cint, (Ptr{Cvoid}, Int64), S.h, id))
mat_set(S, id, col0, A::VecOrMat{Float64}) = chk(ccall((:kb200_mat_set, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}), S.h, id, col0, size(A, 2), A))
mat_get!(out::VecOrMat{Float64}, S, id, col0) = (chk(ccall((:kb200_mat_get, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}), S.h, id, col0, size(out, 2), out)); out)
apply(S, which, src, scol, dst, dcol, ncols=1) = chk(ccall((:kb200_op_apply, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int64, Int64), S.h, which, src, scol, dst, dcol, ncols))
axpby(S, a, x, xc, b, y, yc) = chk(ccall((:kb200_op_axpby, LIB), Cint, (Ptr{Cvoid}, Float64, Int64, Int64, Float64, Int64, Int64), S.h, a, x, xc, b, y, yc))
dot(S, x, xc, y, yc) = (o = Ref{Float64}(0); chk(ccall((:kb200_op_dot, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ref{Float64}), S.h, x, xc, y, yc, o)); o[])
gemv_t(S, A, col0, k, x, xc) = (y = Vector{Float64}(undef, k); chk(ccall((:kb200_op_gemv_t, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Ptr{Float64}), S.h, A, col0, k, x, xc, y)); y)
gemv_n(S, a, A, col0, k, mu, b, y, yc) = chk(ccall((:kb200_op_gemv_n, LIB), Cint, (Ptr{Cvoid}, Float64, Int64, Int64, Int64, Ptr{Float64}, Float64, Int64, Int64), S.h, a, A, col0, k, mu, b, y, yc))
gemm_tn(S, A, ca, ka, B, cb, kb) = (C = Matrix{Float64}(undef, ka, kb); chk(ccall((:kb200_op_gemm_tn, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64, Ptr{Float64}), S.h, A, ca, ka, B, cb, kb, C)); C)
gemm_nn(S, A, ca, ka, Q::Matrix{Float64}, C, cc) = chk(ccall((:kb200_op_gemm_nn, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Int64), S.h, A, ca, ka, Q, size(Q, 2), C, cc))
# Up to m iterations of the (deflated) recurrence on the device without host synchronisation: the body
# of the loops eigcg.jl:185-246 / defcg.jl:400-448 between two restarts (see ../solvers.py: eigpcg,
# eigdefpcg).  Returns (n_done, alpha, beta, res, rTz); V[:, vcol0 + j] receives z / sqrt(rTz).
function pcg_burst(S, precond::Bool, wk, m, rTz, tol, it0, maxit; V=-1, vcol0=0, W=-1, AW=-1, nvec=0,
                   WtAW::Union{Matrix{Float64},Nothing}=nothing, WtW::Union{Matrix{Float64},Nothing}=nothing)
  al = Vector{Float64}(undef, m); be = similar(al); rs = similar(al)
  nd = Ref{Int64}(0); rz = Ref{Float64}(0)
  ptr(A) = A === nothing ? Ptr{Float64}(C_NULL) : pointer(A)
  A1i = WtAW === nothing ? nothing : inv(WtAW); A2i = WtW === nothing ? nothing : inv(WtW)
  GC.@preserve WtAW WtW A1i A2i chk(ccall((:kb200_pcg_burst, LIB), Cint,
      (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64,
       Int64, Int64, Float64, Float64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Float64}),
      S.h, S.assembled ? 1 : 0, precond ? 1 : 0, wk, W, AW, nvec, ptr(WtAW), ptr(A1i), ptr(WtW), ptr(A2i), m,
      V, vcol0, rTz, tol, it0, maxit, al, be, rs, nd, rz))
  k = nd[]
  return k, al[1:k], be[1:k], rs[1:k], rz[]
end
end

"""
    defpcg(S, b, x, W, M; maxit=0) -> (x, it, res_norm)          # defcg.jl:242-308
"""
function defpcg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64}, W::Matrix{Float64},
                M::Union{B200NeumannNeumann,Nothing}; maxit::Int=0)
  D = DeviceOps
  n, nvec = size(W)
  op = S.assembled ? 1 : 0
  X, R, Z, P, AP, B = 0, 1, 2, 3, 4, 5                       # columns of the work block (0-based)
  wk = D.mat_alloc(S, 7); Wd = D.mat_alloc(S, nvec); AWd = D.mat_alloc(S, nvec)
  try
    D.mat_set(S, Wd, 0, W); D.mat_set(S, wk, B, b); D.mat_set(S, wk, X, x)
    D.apply(S, op, Wd, 0, AWd, 0, nvec)                       # WtA[ivec, :] .= A * W[:, ivec]   (:261-264)
    WtAW = D.gemm_tn(S, AWd, 0, nvec, Wd, 0, nvec)            # :268
    resid!() = (D.apply(S, op, wk, X, wk, AP); D.axpby(S, 1.0, wk, B, 0.0, wk, R); D.axpby(S, -1.0, wk, AP, 1.0, wk, R))
    precond!() = M === nothing ? D.axpby(S, 1.0, wk, R, 0.0, wk, Z) : D.apply(S, 2, wk, R, wk, Z)
    resid!()
    mu = WtAW \ D.gemv_t(S, Wd, 0, nvec, wk, R)               # :272-274
    D.gemv_n(S, 1.0, Wd, 0, nvec, mu, 1.0, wk, X)             # x .+= W * mu
    maxit == 0 ? maxit = n : nothing
    res_norm = Vector{Float64}(undef, maxit + 1)
    it = 1
    resid!()
    rTr = D.dot(S, wk, R, wk, R); res_norm[it] = sqrt(rTr)
    precond!()
    rTz = D.dot(S, wk, R, wk, Z)
    mu = WtAW \ D.gemv_t(S, AWd, 0, nvec, wk, Z)              # :284
    D.axpby(S, 1.0, wk, Z, 0.0, wk, P); D.gemv_n(S, -1.0, Wd, 0, nvec, mu, 1.0, wk, P)
    tol = eps_rtol * norm(b)
    while (it < maxit) && (res_norm[it] > tol)                # :290
      D.apply(S, op, wk, P, wk, AP)
      d = D.dot(S, wk, P, wk, AP)
      alpha = rTz / d
      beta = 1. / rTz
      D.axpby(S, alpha, wk, P, 1.0, wk, X)
      D.axpby(S, -alpha, wk, AP, 1.0, wk, R)
      rTr = D.dot(S, wk, R, wk, R)
      precond!()
      rTz = D.dot(S, wk, R, wk, Z)
      beta *= rTz
      mu = WtAW \ D.gemv_t(S, AWd, 0, nvec, wk, Z)            # :301
      D.axpby(S, 1.0, wk, Z, beta, wk, P); D.gemv_n(S, -1.0, Wd, 0, nvec, mu, 1.0, wk, P)   # :302
      it += 1
      res_norm[it] = sqrt(rTr)
    end
    D.mat_get!(x, S, wk, X)
    return x, it, res_norm[1:it]
  finally
    D.mat_free(S, wk); D.mat_free(S, Wd); D.mat_free(S, AWd)
  end
end
export defpcg

# Restart algebra of eigcg.jl:220-228 / defcg.jl:425-432, the reference's own LAPACK lines on the small host
# matrix: returns (Q*Z, vals, nev).
function rayleigh_ritz(Tm::Symmetric{Float64,Matrix{Float64}}, nvec::Int)
  m = size(Tm, 1)
  Y = zeros(m, 2 * nvec)
  Y[:, 1:nvec] = eigvecs(Tm)[:, 1:nvec]
  Y[1:m-1, nvec+1:end] = eigvecs(Symmetric(Tm[1:m-1, 1:m-1]))[:, 1:nvec]
  nev = rank(Y)
  Q = svd(Y).U[:, 1:nev]
  H = Q' * (Tm * Q)
  vals, Z = eigen(Symmetric(H))
  return Q * Z, vals, nev
end

"""
    eigpcg(S, b, x, M, nvec, spdim; maxit=0) -> (x, it, res_norm, V[:, 1:nvec])      # eigcg.jl:143-267

The reference's loop with every n_Γ-sized array on the device.  The iterations between two restarts are
plain PCG steps that append a column to V and two entries to V'AV: they run as ONE device burst
(kb200_pcg_burst, no host synchronisation inside); the restart iteration and the one after it, which
touch `tvec` and the Rayleigh-Ritz algebra, are driven from here.  Mirror of ../solvers.py::eigpcg
(tested against the oracle on the GPU and, through a numpy twin of the block API, on the CPU).
"""
function eigpcg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64},
                M::Union{B200NeumannNeumann,Nothing}, nvec::Int, spdim::Int; maxit::Int=0)
  D = DeviceOps
  n = length(x)
  op = S.assembled ? 1 : 0
  X, R, Z, P, AP, B, T = 0, 1, 2, 3, 4, 5, 6                 # columns of the work block (0-based)
  wk = D.mat_alloc(S, 7); Vd = D.mat_alloc(S, spdim); AVd = D.mat_alloc(S, spdim)
  try
    VtAV = zeros(spdim, spdim)
    D.mat_set(S, wk, B, b); D.mat_set(S, wk, X, x)
    maxit == 0 ? maxit = n : nothing
    res_norm = Vector{Float64}(undef, maxit + 1)
    precond!() = M === nothing ? D.axpby(S, 1.0, wk, R, 0.0, wk, Z) : D.apply(S, 2, wk, R, wk, Z)
    it = 1
    D.apply(S, op, wk, X, wk, AP); D.axpby(S, 1.0, wk, B, 0.0, wk, R); D.axpby(S, -1.0, wk, AP, 1.0, wk, R)
    rTr = D.dot(S, wk, R, wk, R)
    precond!()
    rTz = D.dot(S, wk, R, wk, Z)
    D.axpby(S, 1.0, wk, Z, 0.0, wk, P)
    res_norm[it] = sqrt(rTr)
    ivec = 1                                                  # 1-based column of V, as in the reference
    D.axpby(S, 1.0 / sqrt(rTz), wk, Z, 0.0, Vd, ivec - 1)
    just_restarted = false
    hlpr = 0.0
    tol = eps_rtol * norm(b)
    while (it < maxit) && (res_norm[it] > tol)                # eigcg.jl:185
      if !just_restarted && ivec < spdim
        nd, al, be, rs, rTz = D.pcg_burst(S, M !== nothing, wk, spdim - ivec, rTz, tol, it, maxit; V=Vd, vcol0=ivec)
        for j in 1:nd
          it += 1
          res_norm[it] = rs[j]
          VtAV[ivec, ivec] += 1.0 / al[j]                     # :203
          ivec += 1                                           # V[:, ivec] was stored by the device
          VtAV[ivec-1, ivec] = -sqrt(be[j]) / al[j]
          VtAV[ivec, ivec] = be[j] / al[j]
        end
        continue
      end
      D.apply(S, op, wk, P, wk, AP)
      d = D.dot(S, wk, P, wk, AP)
      alpha = rTz / d
      beta = 1.0 / rTz
      D.axpby(S, alpha, wk, P, 1.0, wk, X)
      D.axpby(S, -alpha, wk, AP, 1.0, wk, R)
      rTr = D.dot(S, wk, R, wk, R)
      precond!()
      just_restarted && (hlpr = sqrt(rTz))                    # :194-196
      rTz = D.dot(S, wk, R, wk, Z)
      beta *= rTz
      ivec == spdim && D.axpby(S, -beta, wk, AP, 1.0, wk, T)  # tvec .-= beta * Ap   :199-201
      D.axpby(S, 1.0, wk, Z, beta, wk, P)
      it += 1
      res_norm[it] = sqrt(rTr)
      VtAV[ivec, ivec] += 1.0 / alpha
      if just_restarted
        D.axpby(S, 1.0, wk, AP, 1.0, wk, T)                   # tvec .+= Ap          :205
        nev = ivec - 1
        VtAV[1:nev, ivec] = D.gemv_t(S, Vd, 0, nev, wk, T) ./ hlpr
        just_restarted = false
      end
      if ivec == spdim                                        # :211
        D.apply(S, op, Vd, 0, AVd, 0, spdim)                  # AV = A * V
        VtAV .= D.gemm_tn(S, Vd, 0, spdim, AVd, 0, spdim)     # V'AV
        QZ, vals, nev = rayleigh_ritz(Symmetric(VtAV), nvec)
        D.gemm_nn(S, Vd, 0, spdim, QZ, Vd, 0)                 # V[:, 1:nev] = V * (Q * Z)
        ivec = nev + 1
        D.axpby(S, 1.0 / sqrt(rTz), wk, Z, 0.0, Vd, ivec - 1)
        VtAV .= 0
        for j in 1:nev
          VtAV[j, j] = vals[j]
        end
        VtAV[ivec, ivec] = beta / alpha
        D.axpby(S, -beta, wk, AP, 0.0, wk, T)                 # tvec .= -beta * Ap   :237
        just_restarted = true
      else
        ivec += 1
        D.axpby(S, 1.0 / sqrt(rTz), wk, Z, 0.0, Vd, ivec - 1)
        VtAV[ivec-1, ivec] = -sqrt(beta) / alpha
        VtAV[ivec, ivec] = beta / alpha
      end
    end
    if !just_restarted && ivec > nvec                         # :248-264
      m = ivec - 1
      QZ, vals, nev = rayleigh_ritz(Symmetric(VtAV[1:m, 1:m]), nvec)
      D.gemm_nn(S, Vd, 0, m, QZ, Vd, 0)
    end
    D.mat_get!(x, S, wk, X)
    Vout = Matrix{Float64}(undef, n, nvec)
    D.mat_get!(Vout, S, Vd, 0)
    return x, it, res_norm[1:it], Vout
  finally
    D.mat_free(S, wk); D.mat_free(S, Vd); D.mat_free(S, AVd)
  end
end

"""
    eigdefpcg(S, b, x, M, W, spdim; maxit=0) -> (x, it, res_norm, V[:, 1:nvec])      # defcg.jl:337-473

Deflated with W and recycling; all iterations but the restart ones run as device bursts with the
deflation (p -= W (WtAW \\ (WtA z)), defcg.jl:301-302) and the projection r -= W (WtW \\ (W'r)) (:407) on
the device.  Mirror of ../solvers.py::eigdefpcg.
"""
function eigdefpcg(S::B200SchurOperator, b::Vector{Float64}, x::Vector{Float64},
                   M::Union{B200NeumannNeumann,Nothing}, W::Matrix{Float64}, spdim::Int; maxit::Int=0)
  D = DeviceOps
  n, nvec = size(W)
  op = S.assembled ? 1 : 0
  X, R, Z, P, AP, B = 0, 1, 2, 3, 4, 5
  wk = D.mat_alloc(S, 7); Wd = D.mat_alloc(S, nvec); AWd = D.mat_alloc(S, nvec); Vd = D.mat_alloc(S, spdim)
  try
    D.mat_set(S, Wd, 0, W); D.mat_set(S, wk, B, b); D.mat_set(S, wk, X, x)
    D.apply(S, op, Wd, 0, AWd, 0, nvec)                       # :357-360
    WtAW = D.gemm_tn(S, AWd, 0, nvec, Wd, 0, nvec)
    WtW = D.gemm_tn(S, Wd, 0, nvec, Wd, 0, nvec)
    resid!() = (D.apply(S, op, wk, X, wk, AP); D.axpby(S, 1.0, wk, B, 0.0, wk, R); D.axpby(S, -1.0, wk, AP, 1.0, wk, R))
    precond!() = M === nothing ? D.axpby(S, 1.0, wk, R, 0.0, wk, Z) : D.apply(S, 2, wk, R, wk, Z)
    resid!()
    mu = WtAW \ D.gemv_t(S, Wd, 0, nvec, wk, R)
    D.gemv_n(S, 1.0, Wd, 0, nvec, mu, 1.0, wk, X)
    VtAV = zeros(spdim, spdim)
    first_restart = true
    maxit == 0 ? maxit = n : nothing
    res_norm = Vector{Float64}(undef, maxit + 1)
    it = 1
    resid!()
    rTr = D.dot(S, wk, R, wk, R)
    precond!()
    rTz = D.dot(S, wk, R, wk, Z)
    mu = WtAW \ D.gemv_t(S, AWd, 0, nvec, wk, Z)
    D.axpby(S, 1.0, wk, Z, 0.0, wk, P); D.gemv_n(S, -1.0, Wd, 0, nvec, mu, 1.0, wk, P)
    res_norm[it] = sqrt(rTr)
    VtAV[1:nvec, 1:nvec] = WtAW                               # :390
    D.gemm_nn(S, Wd, 0, nvec, Matrix{Float64}(I, nvec, nvec), Vd, 0)   # V[:, 1:nvec] = W
    just_restarted = false
    ivec = nvec + 1
    D.axpby(S, 1.0 / sqrt(rTz), wk, Z, 0.0, Vd, ivec - 1)
    tol = eps_rtol * norm(b)
    while (it < maxit) && (res_norm[it] > tol)                # :400
      if ivec < spdim
        nd, al, be, rs, rTz = D.pcg_burst(S, M !== nothing, wk, spdim - ivec, rTz, tol, it, maxit; V=Vd, vcol0=ivec,
                                          W=Wd, AW=AWd, nvec=nvec, WtAW=WtAW, WtW=WtW)
        for j in 1:nd
          it += 1
          res_norm[it] = rs[j]
          VtAV[ivec, ivec] += 1.0 / al[j]                     # :417
          ivec += 1
          VtAV[ivec-1, ivec] = -sqrt(be[j]) / al[j]
          VtAV[ivec, ivec] = be[j] / al[j]
        end
        nd > 0 && (just_restarted = false)
        continue
      end
      # restart iteration (ivec == spdim)
      D.apply(S, op, wk, P, wk, AP)
      d = D.dot(S, wk, P, wk, AP)
      alpha = rTz / d
      beta = 1.0 / rTz
      D.axpby(S, alpha, wk, P, 1.0, wk, X)
      D.axpby(S, -alpha, wk, AP, 1.0, wk, R)
      mu2 = WtW \ D.gemv_t(S, Wd, 0, nvec, wk, R)             # r -= W (WtW \ W'r)   :407
      D.gemv_n(S, -1.0, Wd, 0, nvec, mu2, 1.0, wk, R)
      rTr = D.dot(S, wk, R, wk, R)
      precond!()
      rTz = D.dot(S, wk, R, wk, Z)
      beta *= rTz
      mu = WtAW \ D.gemv_t(S, AWd, 0, nvec, wk, Z)
      D.axpby(S, 1.0, wk, Z, beta, wk, P); D.gemv_n(S, -1.0, Wd, 0, nvec, mu, 1.0, wk, P)
      it += 1
      res_norm[it] = sqrt(rTr)
      VtAV[ivec, ivec] += 1.0 / alpha
      if first_restart                                        # :420-423
        VtAV[1:nvec, nvec+1:spdim] = D.gemm_tn(S, AWd, 0, nvec, Vd, nvec, spdim - nvec)
        first_restart = false
      end
      QZ, vals, nev = rayleigh_ritz(Symmetric(VtAV), nvec)
      D.gemm_nn(S, Vd, 0, spdim, QZ, Vd, 0)
      ivec = nev + 1
      D.axpby(S, 1.0 / sqrt(rTz), wk, Z, 0.0, Vd, ivec - 1)
      VtAV .= 0
      for j in 1:nev
        VtAV[j, j] = vals[j]
      end
      VtAV[ivec, ivec] = beta / alpha
      just_restarted = true
    end
    if !just_restarted && ivec > nvec                         # :450-466
      m = ivec - 1
      if first_restart && m > nvec
        VtAV[1:nvec, nvec+1:m] = D.gemm_tn(S, AWd, 0, nvec, Vd, nvec, m - nvec)
      end
      QZ, vals, nev = rayleigh_ritz(Symmetric(VtAV[1:m, 1:m]), nvec)
      D.gemm_nn(S, Vd, 0, m, QZ, Vd, 0)
    end
    D.mat_get!(x, S, wk, X)
    Vout = Matrix{Float64}(undef, n, nvec)
    D.mat_get!(Vout, S, Vd, 0)
    return x, it, res_norm[1:it], Vout
  finally
    D.mat_free(S, wk); D.mat_free(S, Wd); D.mat_free(S, AWd); D.mat_free(S, Vd)
  end
end
export eigpcg, eigdefpcg

end # module

/* libkrylov_b200.so -- C ABI of the B200-native Schur-complement PCG hot path.
 *
 * Drop-in boundary for venkovic/julia-phd-krylov-spdes: the reference's Julia API keeps its
 * call shapes and `ccall`s these entry points for everything executed per PCG iteration
 * (INTEGRATION.md shows the Julia binding).  Each entry point cites the reference code it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every function returns 0 on success and a negative kb200_status otherwise; the message
 *    is available from kb200_last_error() (thread-local); nothing throws or exits across the ABI;
 *  - sparse matrices are Julia-native SparseMatrixCSC{Float64,Int}: Int64 **1-based** colptr /
 *    rowval + Float64 nzval; index maps are Int64 1-based; dense matrices are column-major;
 *  - the library copies every input during the call and never keeps or frees caller memory;
 *  - one handle = one calling thread at a time (Julia's synchronous ccall model); a handle drives one GPU
 *    (kb200_create, kb200_create_dist: one process per GPU) or several (kb200_create_multi: one process);
 *  - there is no CPU fallback: without a usable CUDA device kb200_create fails.
 */
#ifndef KRYLOV_B200_H
#define KRYLOV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb200_ctx* kb200_handle;

typedef enum {
  KB200_OK = 0,
  KB200_ERR_INVALID = -1,   /* bad argument / call order */
  KB200_ERR_CUDA = -2,      /* CUDA runtime error */
  KB200_ERR_NCCL = -3,      /* NCCL error */
  KB200_ERR_INPUT = -4,     /* malformed matrix / factor */
  KB200_ERR_NOT_SPD = -5,   /* p'Ap <= 0 met during a solve (reported, solve stopped) */
  KB200_ERR_INTERNAL = -6
} kb200_status;

/* Per-solve measurements (device times from CUDA events on the library's stream). */
typedef struct {
  double solve_ms;          /* whole solve on the device, host<->device copies excluded */
  double schur_ms;          /* sum over launches of the local-Schur stream kernel */
  double nn_ms;             /* sum over launches of the Neumann-Neumann GEMV kernel */
  int64_t schur_launches;
  int64_t nn_launches;
  int64_t kernel_launches;  /* every kernel of this library launched by the call */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
} kb200_timings;

const char* kb200_last_error(void);

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Single-process handle on device `device_id`. */
int kb200_create(int device_id, kb200_handle* out);
/* One process per GPU: `nccl_id` is the 128-byte ncclUniqueId produced on rank 0 by
 * kb200_nccl_unique_id and distributed by the host (torch.distributed / MPI / files).  Subdomains
 * are sharded over ranks; every apply ends in the sum of the ranks' partial interface vectors
 * (reference sketch: Fem/EllipticPdePllDomainDecomposition.jl:1-18, `@distributed (+)`), carried out
 * inside the assembly kernel over NVLink peer memory (CUDA IPC buffers exchanged through the
 * communicator at the first operation) or, if that mapping is unavailable, by ncclAllReduce.
 * All ranks must issue the same sequence of operator applications / solves. */
int kb200_nccl_unique_id(void* out128);
int kb200_create_dist(int device_id, int rank, int world, const void* nccl_id, kb200_handle* out);
/* ONE process, several GPUs of one box (the reference's examples are a single Julia process; SURVEY 8(b)
 * `kb200_create(ngpus, device_ids)`): the handle owns one member context per GPU (device_ids == NULL: devices
 * 0 .. ngpus-1) and one host thread per member.  Every call of this header made on the handle is carried out by
 * all members at once -- subdomains sharded as for kb200_create_dist, per-subdomain uploads taken by the owning
 * member, interface sums inside the fused assembly + peer-memory exchange kernels (buffers mapped with
 * cudaDeviceEnablePeerAccess; no NCCL, no fallback: creation of the exchange fails without peer access) -- and the
 * caller's output buffers are written once.  Results are bit-identical to the multi-process handle and, for the
 * operators, to a single-GPU handle (owner-ordered sums).  The device-resident block API (kb200_mat_*, kb200_op_*,
 * kb200_pcg_burst) and the development aids are not available on such a handle; the whole-solve entry points are. */
int kb200_create_multi(int ngpus, const int* device_ids, kb200_handle* out);
int kb200_destroy(kb200_handle h);
/* Tunables (0 = automatic where it applies): "csize" (CTAs per subdomain cluster: 1,2,4,8,16),
 * "warps" (consumer warps of the stream kernel: 8, or 16 with one CTA per SM), "stages" (ring depth,
 * 2..12), "lane_entries" / "pad_tol_pct" / "min_block" / "adaptive_lanes" / "warp_private" (host
 * analysis of the stream, see csrc/plan.h), "profile" (0/1: sample the two dominant kernels with
 * CUDA events), "check_every" (iterations between convergence polls, default 32), "fuse_pcg" (1:
 * plain PCG runs assembly + dot + vector update in one kernel), "peer_exchange" (1: interface sums
 * over NVLink peer memory; 0: NCCL all-reduce), "vec_threads" / "gemv_rows" (block shapes of the
 * interface-vector / dense GEMV kernels), "pdl" (programmatic dependent launch, default 0),
 * "shard_world" / "shard_rank" (hold one shard of the subdomains without a communicator). */
int kb200_set_option(kb200_handle h, const char* name, int64_t value);

/* ---- setup (once per mesh / per realization) ------------------------------------------------- */
/* Flattened outputs of set_subdomains (Fem/EllipticPdeDomainDecomposition.jl:86-193):
 * gamma_l2g[d][l] = interface id of local interface node l of subdomain d, i.e. the inverse of the
 * Dict ind_Γd_Γ2l[d] (cf. ind_Γd_l2Γ at :1090-1095); node_gamma_cnt = node_Γ_cnt. */
int kb200_set_topology(kb200_handle h, int64_t ndom, int64_t n_gamma, const int64_t* n_I,
                       const int64_t* n_gd, const int64_t* const* gamma_l2g,
                       const int64_t* node_gamma_cnt);
/* Subdomains [d0, d1) (0-based) live on this handle's GPU (all of them without NCCL). Uploads for
 * other subdomains are accepted and ignored. */
int kb200_local_range(kb200_handle h, int64_t* d0, int64_t* d1);
/* Blocks of prepare_local_schurs (:389-546) for subdomain d (0-based): A_IΓ,d (n_I x n_gd, local
 * interface columns) and A_ΓΓ,d (n_gd x n_gd). */
int kb200_upload_local_blocks(kb200_handle h, int64_t d,
                              const int64_t* AIG_colptr, const int64_t* AIG_rowval, const double* AIG_nzval,
                              const int64_t* AGG_colptr, const int64_t* AGG_rowval, const double* AGG_nzval);
/* Exact interior solve A_II,d^{-1} = P' L^-T L^-1 P from F = cholesky(A_IIdd[d]) (:1536;
 * MyPreconditioners/CholPreconditioners.jl:11): perm = F.p, L = sparse(F.L) (lower, diagonal
 * stored, sorted rows).  Replaces the inner IterativeSolvers.cg(A_IIdd, ., Pl=...) of :648-650. */
int kb200_upload_cholesky(kb200_handle h, int64_t d, const int64_t* perm,
                          const int64_t* L_colptr, const int64_t* L_rowval, const double* L_nzval);
/* prepare_global_schur variant (:212-369; operator apply_global_schur :596-625, the one Example03:175 checks
 * against the local-column operator): A_IΓ,d with GLOBAL interface columns (n_I x n_gamma; every non-empty
 * column must be an interface node of subdomain d) instead of kb200_upload_local_blocks, and the interface
 * block as ONE assembled A_ΓΓ (n_gamma x n_gamma).  A handle fed this way applies
 * Sx = A_ΓΓ x - Σ_d A_IΓd' A_IId^{-1} A_IΓd x in every operator application and solver. */
int kb200_upload_global_blocks(kb200_handle h, int64_t d, const int64_t* AIGg_colptr, const int64_t* AIGg_rowval,
                               const double* AIGg_nzval);
int kb200_upload_global_interface_block(kb200_handle h, const int64_t* AGG_colptr, const int64_t* AGG_rowval,
                                        const double* AGG_nzval);
/* New coefficient realization on the same mesh / partition / ordering (Example07:160-204 loop): same
 * sparsity patterns, new values.  On a finalized handle the values go straight into the device-resident value
 * pool; before the next operation three kernels re-invert the supernode blocks and rewrite the slots of the
 * resident stream (no host analysis, no re-upload: ~0.02 s at 4M DOF).  The uploads above may also simply be
 * repeated (that re-runs the host analysis).  A dense S_d of the previous realization is dropped. */
int kb200_update_cholesky_values(kb200_handle h, int64_t d, const double* L_nzval);
int kb200_update_local_block_values(kb200_handle h, int64_t d, const double* AIG_nzval, const double* AGG_nzval);
/* ---- numeric re-factorisation on the device (next row f1) ----------------------------------------
 * Example07_PcgSchurStochasticEllipticPde.jl:160-204 refills A_II,d and calls cholesky() per realization
 * (EllipticPdeDomainDecomposition.jl:389-546, :1536).  With these two calls the factorisation itself moves to the
 * GPU: the permutation and the pattern of the factor uploaded once by kb200_upload_cholesky ARE the symbolic
 * factorisation; kb200_upload_interior_pattern gives the pattern of A_II,d (SparseMatrixCSC colptr / rowval, Int64
 * 1-based, both triangles or one, rows of a column sorted), kb200_refactor_interior its new values (nzval, same
 * pattern).  Before the next operation a level-scheduled numeric Cholesky (dot-product form, one launch per
 * dependency level, all local subdomains at once, deterministic sums) recomputes the values of L_d in the resident
 * value pool, after which the value refresh of kb200_update_cholesky_values re-inverts the supernode blocks and
 * rewrites the stream.  All local subdomains must have been given values at least once before the first operation
 * that follows; a non-positive pivot is reported as KB200_ERR_INPUT.  A_IΓ,d / A_ΓΓ,d still arrive through
 * kb200_update_local_block_values. */
int kb200_upload_interior_pattern(kb200_handle h, int64_t d, const int64_t* colptr, const int64_t* rowval);
int kb200_refactor_interior(kb200_handle h, int64_t d, const double* A_II_nzval);
/* ---- re-assembly of the local blocks on the device (next row f1) ---------------------------------
 * The element loop of prepare_local_schurs (EllipticPdeDomainDecomposition.jl:430-497) for a new coefficient, on the
 * GPU: a realization then costs the host ONE vector (`coeff`, one double per mesh node) instead of the three nzval
 * arrays of every subdomain.  The loop nest of the reference visits "contributions" in a fixed order -- element iel
 * (1-based), vertex i (outer), vertex j (inner) -- and pushes ΔK_ij = Δa (Δy_i Δy_j + Δx_i Δx_j) / 4 / Area (:470)
 * onto the triplet list of one of the three blocks; sparse() (:540-543) adds the duplicates of an entry in the order
 * they were pushed.  The caller describes exactly that:
 *   kb200_upload_assembly_mesh: `cells` (3 x nel, Int64 1-based, column-major as the reference's array) and `points`
 *     (2 x nnode) of Mesh.jl:49-91;
 *   kb200_upload_assembly_map(h, d, which, nnz, ptr, src): for block `which` of subdomain d (0: A_II,d  1: A_IΓ,d
 *     2: A_ΓΓ,d) and every stored entry p of its nzval array (CSC order, the pattern uploaded before) the
 *     contributions src[ptr[p] .. ptr[p+1]-1] (ptr 1-based like colptr, nnz + 1 entries), each the 1-based position
 *     9 (iel - 1) + 3 (i - 1) + j of a contribution in the loop nest, in ascending order;
 *   kb200_assemble_realization(h, coeff): coeff[nnode] -> Δa per element as :437-446 (sum of the three vertex values in
 *     the order of `cells`, divided by 3), ΔK_ij with the reference's association order and no fused multiply-adds,
 *     duplicates added in ascending contribution order: the values have the BITS of prepare_local_schurs.  A_IΓ,d /
 *     A_ΓΓ,d go straight to the value pool, A_II,d to the staging area of kb200_refactor_interior (which must have
 *     been prepared with kb200_upload_interior_pattern); the next operation runs the numeric Cholesky and the
 *     stream refresh on the device.  The right-hand sides b_I,d / b_Γ (:499-536) stay with the caller.
 *   kb200_get_local_block_values reads the current values back (any pointer may be null; A_II only once staged).
 *   kb200_upload_rhs_map (optional) puts the right-hand sides b_I,d / b_Γ of the same loop (:499-536) on the device too.
 *     Only the Dirichlet-lift terms (:499-508) depend on the coefficient, and only in the entries of nodes that share
 *     an element with a Dirichlet node; the caller passes the load vector once (bI0_concat: sum_d n_I,d entries in
 *     subdomain order, bG0: n_gamma entries; what prepare_local_schurs returns for ANY coefficient is fine, the nt
 *     entries below are overwritten) and, for each of those nt entries, tgt[t] = its 1-based index in [b_I ; b_Γ]
 *     and its terms src/val[ptr[t] .. ptr[t+1]-1] (ptr 1-based) in the order the loop nest applies them:
 *     src > 0: `b -= ΔK_src * val` (val = uexact at the Dirichlet vertex, src the contribution index as above),
 *     src = 0: `b += val` (val = Δb of :524-526).  After that every kb200_assemble_realization also assembles the
 *     right-hand sides (bit for bit the reference's), kb200_schur_rhs(h, NULL, NULL, b_schur) and
 *     kb200_back_substitute(h, u_gamma, NULL, u_I) use them, kb200_get_rhs reads them back. */
int kb200_upload_assembly_mesh(kb200_handle h, int64_t nnode, int64_t nel, const int64_t* cells, const double* points);
int kb200_upload_assembly_map(kb200_handle h, int64_t d, int which, int64_t nnz, const int64_t* ptr, const int64_t* src);
int kb200_assemble_realization(kb200_handle h, const double* coeff);
int kb200_get_local_block_values(kb200_handle h, int64_t d, double* AII_nzval, double* AIG_nzval, double* AGG_nzval);
int kb200_upload_rhs_map(kb200_handle h, const double* bI0_concat, const double* bG0, int64_t nt, const int64_t* tgt,
                         const int64_t* ptr, const int64_t* src, const double* val);
int kb200_get_rhs(kb200_handle h, double* b_I_concat, double* b_gamma);
/* ΠSd[d] of NeumannNeumannSchurPreconditioner (:1125), n_gd x n_gd column-major. */
int kb200_upload_nn_pinv(kb200_handle h, int64_t d, const double* PiS);
/* Dense local Schur complement S_d (assemble_local_schurs, :667-695), n_gd x n_gd column-major. */
int kb200_upload_dense_local_schur(kb200_handle h, int64_t d, const double* Sd);
/* Host analysis (level schedules, cluster mapping) + upload to the GPU.  Called implicitly by the
 * first operation that needs it. */
int kb200_finalize(kb200_handle h);

/* ---- operators (host vectors in, host vectors out) -------------------------------------------- */
/* y = S x, matrix-free: apply_local_schurs (:711-747).  x, y: n_gamma x nrhs column-major. */
int kb200_apply_schur(kb200_handle h, const double* x, double* y, int64_t nrhs);
/* y = S x, global-column operator: apply_global_schur (:596-625); needs the two kb200_upload_global_* calls. */
int kb200_apply_global_schur(kb200_handle h, const double* x, double* y, int64_t nrhs);
/* y = S x with the dense S_d: apply_local_schurs assembled variant (:761-785). */
int kb200_apply_schur_assembled(kb200_handle h, const double* x, double* y, int64_t nrhs);
/* z = M \ r: apply_neumann_neumann_schur / `\` / ldiv! (:1361-1403). */
int kb200_apply_nn(kb200_handle h, const double* r, double* z, int64_t nrhs);
/* b_schur = b_Γ - Σ_d R_d' A_ΓI,d A_II,d^{-1} b_I,d : get_schur_rhs (:835-864).
 * b_I_concat = [b_Id[1]; b_Id[2]; ...] (all subdomains, every rank).  b_I_concat = b_gamma = NULL (here), b_I_concat =
 * NULL (kb200_back_substitute): the right-hand sides the device assembled (kb200_upload_rhs_map). */
int kb200_schur_rhs(kb200_handle h, const double* b_I_concat, const double* b_gamma, double* b_schur);
/* u_I,d = A_II,d^{-1} (b_I,d - A_IΓ,d u_Γ): get_subdomain_solutions (:1014-1025). */
int kb200_back_substitute(kb200_handle h, const double* u_gamma, const double* b_I_concat,
                          double* u_I_concat);
/* Dense S_d of every local subdomain computed on the device column by column and kept for
 * kb200_apply_schur_assembled: assemble_local_schurs (:667-695; Symmetric() of the upper triangle
 * as at :692).  kb200_get_dense_local_schur copies one block out (column-major). */
int kb200_assemble_local_schurs(kb200_handle h);
int kb200_get_dense_local_schur(kb200_handle h, int64_t d, double* Sd_out);

/* ---- solvers (whole solve on the device) ------------------------------------------------------- */
/* pcg(A, b, x, M; maxit) of RecyclingKrylovSolvers/cg.jl:67-109 with A = the Schur operator
 * (assembled != 0: the dense-S_d variant) and M = the Neumann-Neumann preconditioner
 * (precond == 0: identity, i.e. cg.jl:14-50).  x is updated in place; *it and res_norm[0..*it)
 * follow the reference (it = iterations + 1); rtol is RecyclingKrylovSolvers.jl:21 (1e-7);
 * maxit == 0 means n_gamma.  res_norm must hold maxit (or n_gamma) + 1 entries. */
int kb200_pcg(kb200_handle h, const double* b, double* x, int64_t maxit, double rtol, int assembled,
              int precond, int64_t* it, double* res_norm, kb200_timings* timings);

/* pcg for nrhs right-hand sides (B, X: n_gamma x nrhs column-major): independent recurrences per RHS,
 * i.e. the reference's sequential loop (Example11_EigInitPcgEllipticPdeMultipleRhs_Functions.jl:140-179);
 * its[k] and res_norm[k*res_stride ...] per RHS.  Round 1 solves one RHS after the other. */
int kb200_pcg_multi(kb200_handle h, const double* B, double* X, int64_t nrhs, int64_t maxit, double rtol,
                    int assembled, int precond, int64_t* its, double* res_norm, int64_t res_stride);
/* defpcg(A, b, x, W, M; maxit) of RecyclingKrylovSolvers/defcg.jl:242-308, whole solve on the device:
 * W (n_gamma x nvec, column-major) is uploaded, WtA' = S W and WtAW are formed on the device, the
 * small system WtAW \ . is applied per iteration as explicit inverse + one refinement step (no host
 * round trip), p = beta p + z - W mu by a dense GEMV kernel. */
int kb200_defpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t maxit,
                 double rtol, int assembled, int precond, int64_t* it, double* res_norm, kb200_timings* timings);

/* eigpcg(A, b, x, M, nvec, spdim; maxit) of RecyclingKrylovSolvers/eigcg.jl:143-267 and eigdefpcg(A, b, x, M, W, spdim; maxit)
 * of defcg.jl:337-473, whole solve inside the library: the iterations between two restarts run as one device burst, the
 * Rayleigh-Ritz restart applies the operator to the spdim basis vectors with the multi-vector stream kernel (eigcg.jl:212-218),
 * forms V'AV and V (Q Z) with FP64 DMMA kernels (:219, :228) and does the spdim-sized eigen / SVD / rank algebra (:222-227, LAPACK in
 * the reference) with cuSOLVER on the device.  Return values as the reference: x in place, it, res_norm[0..it), and
 * V_out = V[:, 1:nvec] (n_gamma x nvec column-major), the approximate least dominant eigenvectors to recycle as W. */
int kb200_eigpcg(kb200_handle h, const double* b, double* x, int64_t nvec, int64_t spdim, int64_t maxit, double rtol,
                 int assembled, int precond, int64_t* it, double* res_norm, double* V_out, kb200_timings* timings);
int kb200_eigdefpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim, int64_t maxit,
                    double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* V_out, kb200_timings* timings);

/* ---- device-resident vectors and bases (deflated / recycling solvers) ----------------------------- */
/* defpcg / eigpcg / eigdefpcg (RecyclingKrylovSolvers/defcg.jl:242-308, eigcg.jl:143-267,
 * defcg.jl:337-473) keep their control flow -- including the Rayleigh-Ritz restarts on small dense
 * matrices (LAPACK eigen / svd / rank, eigcg.jl:220-228) -- in the host language; every n_gamma-sized
 * object (x, r, z, p, Ap, W, WtA', V, AV, tvec) lives on the GPU as a column of a block allocated
 * here, and only scalars and nvec / spdim-sized dense data cross the boundary.
 * A block is n_gamma x ncols, column-major, zero-initialised; (id, col) addresses one column. */
int kb200_mat_alloc(kb200_handle h, int64_t ncols, int64_t* id);
int kb200_mat_free(kb200_handle h, int64_t id);
int kb200_mat_set(kb200_handle h, int64_t id, int64_t col0, int64_t ncols, const double* host);
int kb200_mat_get(kb200_handle h, int64_t id, int64_t col0, int64_t ncols, double* host);
/* dst[:, dcol+k] = Op * src[:, scol+k], k < ncols; which = 0 S matrix-free (`A * W[:, ivec]`,
 * defcg.jl:263), 1 S assembled, 2 M \ . (Neumann-Neumann). */
int kb200_op_apply(kb200_handle h, int which, int64_t src, int64_t scol, int64_t dst, int64_t dcol, int64_t ncols);
/* y = a x + b y  (axpy! / axpby!, cg.jl:97-103) */
int kb200_op_axpby(kb200_handle h, double a, int64_t x, int64_t xcol, double b, int64_t y, int64_t ycol);
/* *out = x . y  (dot, cg.jl:94) */
int kb200_op_dot(kb200_handle h, int64_t x, int64_t xcol, int64_t y, int64_t ycol, double* out);
/* y_host[0:k] = A[:, col0:col0+k]' x        (WtA * z, W' * r: defcg.jl:272, 301) */
int kb200_op_gemv_t(kb200_handle h, int64_t A, int64_t col0, int64_t k, int64_t x, int64_t xcol, double* y_host);
/* y = b y + a A[:, col0:col0+k] mu_host     (W * mu: defcg.jl:275, 302) */
int kb200_op_gemv_n(kb200_handle h, double a, int64_t A, int64_t col0, int64_t k, const double* mu_host, double b,
                    int64_t y, int64_t ycol);
/* C_host (ka x kb, column-major) = A[:, ca:ca+ka]' B[:, cb:cb+kb]   (WtA * W, V'AV: defcg.jl:268, eigcg.jl:219) */
int kb200_op_gemm_tn(kb200_handle h, int64_t A, int64_t ca, int64_t ka, int64_t B, int64_t cb, int64_t kb,
                     double* C_host);
/* C[:, cc:cc+kq] = A[:, ca:ca+ka] Q_host (ka x kq, column-major); C may alias A (V[:, 1:nev] = V*(Q*Z), eigcg.jl:228) */
int kb200_op_gemm_nn(kb200_handle h, int64_t A, int64_t ca, int64_t ka, const double* Q_host, int64_t kq, int64_t C,
                     int64_t cc);

/* Continue a (deflated) PCG recurrence for up to m iterations WITHOUT host synchronisation inside: the
 * body of the loops eigcg.jl:185-246 / defcg.jl:400-448 between two restarts of the recycling solvers
 * (cg.jl:93-105 per iteration, plus p -= W (WtAW \ (WtA z)) defcg.jl:301-302 when nvec > 0, plus
 * r -= W (WtW \ (W'r)) defcg.jl:407 when WtW is given).  Vectors are columns 0..4 = x, r, z, p, Ap of block
 * `wk`; W, AW are blocks of nvec columns; WtAW / WtW and their inverses are nvec x nvec column-major host
 * arrays (the device applies the inverse with one step of iterative refinement).  After iteration j
 * (0-based) the column V[:, vcol0 + j] receives z / sqrt(rTz) (eigcg.jl:242; V < 0: not stored) and
 * alpha_out[j], beta_out[j], res_out[j] = alpha, beta, ||r||_2 of that iteration.  The recurrence starts from
 * the host-side scalars rTz, `it0` (the reference's `it`) and stops like the reference's loop
 * (it < maxit && ||r|| > tol).  n_done iterations were run; rTz_out is the final r'z. */
int kb200_pcg_burst(kb200_handle h, int assembled, int precond, int64_t wk, int64_t W, int64_t AW, int64_t nvec,
                    const double* WtAW, const double* WtAW_inv, const double* WtW, const double* WtW_inv, int64_t m,
                    int64_t V, int64_t vcol0, double rTz, double tol, int64_t it0, int64_t maxit, double* alpha_out,
                    double* beta_out, double* res_out, int64_t* n_done, double* rTz_out);

/* ---- introspection (tests / benchmarks) -------------------------------------------------------- */
/* stats[0..n): 0 Σ nnz(L) strictly lower, 1 Σ stream entries, 2 Σ rows, 3 Σ slices,
 * 4 stream bytes read per Schur apply, 5 max levels, 6 max top size, 7 shared memory per CTA,
 * 8 CTAs per launch, 9 cluster size, 10 dense bytes read per NN apply. */
int kb200_get_stats(kb200_handle h, double* stats, int64_t n);
/* Gather / scatter maps the device uses, for bit-for-bit comparison with the reference's
 * (0-based): adj_ptr (n_gamma+1), adj_dom / adj_loc (Σ n_gd): contributions to each interface
 * node in ascending subdomain order. */
int kb200_get_interface_maps(kb200_handle h, int64_t* adj_ptr, int64_t* adj_dom, int64_t* adj_loc);
/* Average device time (ms) of `reps` back-to-back launches of one kernel on resident data:
 * which = 0 Schur stream kernel, 1 NN GEMV kernel, 2 assembled-Schur GEMV kernel, 3 the fused assembly
 * (+ peer exchange) / vector-update kernel (development probe: it leaves garbage in the work vectors; on
 * several ranks every rank must make the call), 4 / 5 the multi-vector stream kernel with 2 / 4 right-hand sides per
 * pass over the factor, 6..9 the batched DMMA Neumann-Neumann product with 8 / 16 / 32 / 64 right-hand sides. */
int kb200_time_kernel(kb200_handle h, int which, int reps, double* avg_ms);
/* Test aid: the Rayleigh-Ritz restart algebra of eigpcg / eigdefpcg (eigcg.jl:220-228) alone, on a symmetric m x m matrix
 * Tm (column-major): QZ_out (m x 2 nvec), vals_out (2 nvec), *nev_out = rank found. */
int kb200_debug_rayleigh_ritz(kb200_handle h, const double* Tm, int64_t m, int64_t nvec, double* QZ_out, double* vals_out,
                              int64_t* nev_out);
/* ---- post-hoc recycling variants (next row f3) --------------------------------------------------
 * rrpcg / rrdefpcg (RecyclingKrylovSolvers/rrdefpcg.jl:204-275, :48-151: Rayleigh-Ritz projection of [W Z], Z = the
 * first spdim (- nvec) preconditioned residuals z / sqrt(r'z)) and hrpcg / hrdefpcg (hrdefpcg.jl:217-295, :48-164:
 * harmonic Ritz projection of [W P], P = the first search directions): the (deflated) solve plus ONE projection at the
 * end, whole on the device; W2_out (n_Γ x nvec) = the new deflation basis.  Arguments as kb200_eigpcg / kb200_eigdefpcg.
 * When the solve stops before the search space is full the projection uses the columns that exist (the reference
 * would read uninitialised memory there).
 * trrrpcg / trrrdefpcg (trrrdefpcg.jl:234-355, :48-199) and trhrpcg / trhrdefpcg (trhrdefpcg.jl:247-389, :48-208): the
 * thick-restart siblings -- the window of captured columns is cyclic and every time it is full W2 becomes the nvec
 * lowest (harmonic) Ritz vectors of [W2 | window]; a solve that stops before its first window is full returns W2 = W
 * (deflated) or zeros.
 * lotrrrpcg / lotrrrdefpcg (lotrrrdefpcg.jl:252-389, :48-214) and lotrhrpcg / lotrhrdefpcg (lotrhrdefpcg.jl:267-428,
 * :48-229): the locally-optimal thick-restart siblings -- the projection is the restart of eigCG (eigcg.jl:220-228)
 * applied to the pencil of the search space [W2 | window], W2 carries nvec..2 nvec columns inside the solve and its
 * first nvec columns are returned; spdim >= 2 nvec + 2. */
int kb200_rrpcg(kb200_handle h, const double* b, double* x, int64_t nvec, int64_t spdim, int64_t maxit, double rtol, int assembled,
                int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_rrdefpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim, int64_t maxit,
                   double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_hrpcg(kb200_handle h, const double* b, double* x, int64_t nvec, int64_t spdim, int64_t maxit, double rtol, int assembled,
                int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_hrdefpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim, int64_t maxit,
                   double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_trrrpcg(kb200_handle h, const double* b, double* x, int64_t nvec, int64_t spdim, int64_t maxit, double rtol, int assembled,
                  int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_trrrdefpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim, int64_t maxit,
                     double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_trhrpcg(kb200_handle h, const double* b, double* x, int64_t nvec, int64_t spdim, int64_t maxit, double rtol, int assembled,
                  int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_trhrdefpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim, int64_t maxit,
                     double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_lotrrrpcg(kb200_handle h, const double* b, double* x, int64_t nvec, int64_t spdim, int64_t maxit, double rtol, int assembled,
                    int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_lotrrrdefpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim, int64_t maxit,
                       double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_lotrhrpcg(kb200_handle h, const double* b, double* x, int64_t nvec, int64_t spdim, int64_t maxit, double rtol, int assembled,
                    int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
int kb200_lotrhrdefpcg(kb200_handle h, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim, int64_t maxit,
                       double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* W2_out, kb200_timings* timings);
/* all of them behind one entry: kind 0 = Rayleigh-Ritz, 1 = harmonic Ritz, 2 / 3 = their thick-restart siblings, 4 / 5 = the
 * locally-optimal thick-restart ones; W == NULL: no deflation basis */
int kb200_posthoc_recycle(kb200_handle h, int kind, const double* b, double* x, const double* W, int64_t nvec, int64_t spdim,
                          int64_t maxit, double rtol, int assembled, int precond, int64_t* it, double* res_norm, double* W2_out,
                          kb200_timings* timings);
/* ---- LORASC preconditioner of the full system (next row f2) ------------------------------------
 * LorascPreconditioner / apply_lorasc (EllipticPdeDomainDecomposition.jl:1406-1428, :1908-1976; `\` :1980-1982, ldiv!
 * :1984-1995): u = apply_lorasc(x) on the (interior..., interface) splitting of the full system.
 * kb200_upload_lorasc takes chol_A_ΓΓ as Julia extracts it (F.p, sparse(F.L) of cholesky(A_ΓΓ), :1525) and the nev
 * correction vectors E (n_Γ x nev, column-major, :1418, as prepare_lorasc_precond :1502-1676 leaves them -- that set-up
 * stays on the host); the interior factors and A_IΓ,d are the ones already on the handle.  kb200_apply_lorasc:
 * x_I_concat / u_I_concat are the interior parts concatenated in subdomain order (x_Id[idom], :1921-1926), x_gamma /
 * u_gamma the interface parts; the gather / scatter through not_dirichlet_inds_g2l (:1921-1933, :1963-1973) is the
 * caller's (host-language) side.  Works on single-GPU and kb200_create_multi handles. */
int kb200_upload_lorasc(kb200_handle h, const int64_t* perm, const int64_t* L_colptr, const int64_t* L_rowval,
                        const double* L_nzval, const double* E, int64_t nev);
int kb200_apply_lorasc(kb200_handle h, const double* x_I_concat, const double* x_gamma, double* u_I_concat, double* u_gamma);
/* Development aid: with option "debug_timing" = 1 the stream kernel records 16 clock64 stamps per
 * CTA (phase boundaries); this copies them out. */
int kb200_debug_stamps(kb200_handle h, long long* out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* KRYLOV_B200_H */

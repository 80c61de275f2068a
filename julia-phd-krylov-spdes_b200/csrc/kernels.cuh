// Device kernels of libkrylov_b200 (sm_100a).  FP64 throughout.
//
//   schur_stream_kernel   batched local Schur apply / interior solve: one thread-block cluster per
//                         subdomain, interior vectors resident in (distributed) shared memory,
//                         level-scheduled sparse triangular solves + SpMVs streamed from HBM
//                         (reference: apply_local_schur, EllipticPdeDomainDecomposition.jl:639-654)
//   dense_gemv_kernel     batched dense local products  Pi S_d (D r)_d  (Neumann-Neumann apply,
//                         :1361-1386) and  S_d x_d  (assembled Schur apply, :761-785)
//   assemble_*            deterministic gather-form interface assembly  y[g] = sum_d R_d^T (.)
//   pcg_*                 fused PCG vector updates with two-stage deterministic reductions
//                         (RecyclingKrylovSolvers/cg.jl:92-106)
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "plan.h"

namespace kb {
namespace cg = cooperative_groups;

struct StreamParams {
  const CtaDesc* ctas;
  const SliceHdr* slices;
  const RowRec* rows;
  const double* vals;
  const uint16_t* cols;
  const int32_t* l2g;
  const int32_t* imap;
  const double* x_gamma;  // interface vector (MODE_APPLY / MODE_BACKSUB)
  const double* bI;       // concatenated interior vector in (MODE_RHS / MODE_BACKSUB)
  double* uI;             // concatenated interior vector out (MODE_BACKSUB)
  double* out_loc;        // concatenated local-interface vector out (MODE_APPLY / MODE_RHS)
  double* sd_pool;        // dense S_d blocks, column-major (MODE_UNIT)
  const int64_t* sd_off;  // per local subdomain offset into sd_pool
  const int* done_at;     // optional: skip when *done_at < iter (solver already converged)
  int iter;
  uint64_t n_vals, n_rows, n_slices;  // pool sizes (prefetch clamps)
  int unit_col0;          // MODE_UNIT: column = unit_col0 + blockIdx.y
};

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
constexpr int kPrefetchAhead = 1536;  // entries (12 KB of values) per warp stream
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// One section of a CTA's program: every warp walks its slice range; rows of a level are independent.
__device__ __forceinline__ void run_section(const CtaDesc& D, const StreamParams& P, double* S, int sec,
                                            double sign, int warp, int lane) {
  const SliceHdr* slices = P.slices + D.slice_base[warp];
  const RowRec* rows = P.rows + D.row_base[warp];
  const double* vals = P.vals + D.ent_base[warp];
  const uint16_t* cols = P.cols + D.ent_base[warp];
  uint32_t s = D.sec_begin[sec][warp];
  const uint32_t e = D.sec_end[sec][warp];
  const uint32_t nlev = D.sec_nlevels[sec];
  uint32_t lvl = 0;
  for (; s < e; ++s) {
    const SliceHdr h = slices[s];
    while (lvl < h.level) {
      cta_barrier();
      ++lvl;
    }
    const int tl = h.tlog2 & 7;
    const bool fresh = (h.tlog2 & 0x80) != 0;
    const int T = 1 << tl;
    const int g = lane >> tl;
    const int t = lane & (T - 1);
    const bool act = g < (int)h.nrows;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    uint32_t dst = 0;
    double dg = 1.0;
    // keep the L2 ahead of this warp's strictly sequential streams
    if (lane == 1) prefetch_l2(slices + s + 48);
    if (lane == 2) prefetch_l2(rows + h.row0 + 96);
    if (act) {
      const RowRec rr = rows[h.row0 + g];
      dst = rr.dst;
      dg = rr.diag;
      const double* v = vals + rr.ebeg;
      const uint16_t* c = cols + rr.ebeg;
      const int len = rr.len;
      int k = t;
      for (; k + 3 * T < len; k += 4 * T) {
        const double v0 = __ldg(v + k), v1 = __ldg(v + k + T), v2 = __ldg(v + k + 2 * T),
                     v3 = __ldg(v + k + 3 * T);
        const uint32_t c0 = __ldg(c + k), c1 = __ldg(c + k + T), c2 = __ldg(c + k + 2 * T),
                       c3 = __ldg(c + k + 3 * T);
        acc0 = fma(v0, S[c0], acc0);
        acc1 = fma(v1, S[c1], acc1);
        acc2 = fma(v2, S[c2], acc2);
        acc3 = fma(v3, S[c3], acc3);
      }
      for (; k < len; k += T) acc0 = fma(__ldg(v + k), S[__ldg(c + k)], acc0);
      if (t == 0) {  // one lane per row: the lines this row will be followed by, kPrefetchAhead later
        const double* pv = v + kPrefetchAhead;
        const uint16_t* pc = c + kPrefetchAhead;
        for (int q = 0; q < len; q += 16) prefetch_l2(pv + q);
        for (int q = 0; q < len; q += 64) prefetch_l2(pc + q);
      }
    }
    double acc = (acc0 + acc1) + (acc2 + acc3);
    for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (act && t == 0) S[dst] = ((fresh ? 0.0 : S[dst]) - sign * acc) / dg;
  }
  while (lvl < nlev) {
    cta_barrier();
    ++lvl;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32) schur_stream_kernel(const StreamParams P) {
  extern __shared__ double S[];
  __shared__ CtaDesc D;
  if (P.done_at != nullptr && *P.done_at < P.iter) return;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.ctas + blockIdx.x);
    uint32_t* dstw = reinterpret_cast<uint32_t*>(&D);
    for (int i = tid; i < (int)(sizeof(CtaDesc) / 4); i += blockDim.x) dstw[i] = src[i];
  }
  __syncthreads();
  for (int i = tid; i < D.smem_doubles; i += blockDim.x) S[i] = 0.0;
  __syncthreads();
  const int C = D.csize, rank = D.rank;
  // ---- inputs ---------------------------------------------------------------------------------
  if (MODE == MODE_APPLY || MODE == MODE_BACKSUB) {
    const int32_t* l2g = P.l2g + D.l2g_off;
    for (int l = tid; l < D.n_gd; l += blockDim.x) S[D.xd_off + l] = P.x_gamma[l2g[l]];
  }
  if (MODE == MODE_UNIT) {
    const int col = P.unit_col0 + (int)blockIdx.y;
    if (tid == 0 && col < D.n_gd) S[D.xd_off + col] = 1.0;
  }
  if (MODE == MODE_RHS || MODE == MODE_BACKSUB) {
    const double* b = P.bI + D.I_base;
    const int32_t* om = P.imap + D.ownmap_off;
    for (int k = tid; k < D.n_own; k += blockDim.x) S[D.own_off + k] = b[om[k]];
    if (rank == 0) {
      const int32_t* tm = P.imap + D.topmap_off;
      for (int k = tid; k < D.n_top; k += blockDim.x) S[D.top_off + k] = b[tm[k]];
    }
  }
  __syncthreads();
  // ---- forward ---------------------------------------------------------------------------------
  if (MODE == MODE_APPLY || MODE == MODE_UNIT) run_section(D, P, S, SEC_INIT, -1.0, warp, lane);
  if (MODE == MODE_BACKSUB) run_section(D, P, S, SEC_INIT, 1.0, warp, lane);
  run_section(D, P, S, SEC_FWD, 1.0, warp, lane);
  if (C > 1) {
    run_section(D, P, S, SEC_PART, 1.0, warp, lane);
    cluster.sync();
    if (rank == 0) {
      for (int i = tid; i < D.n_top; i += blockDim.x) {
        double a = S[D.top_off + i];
        for (int c = 0; c < C; ++c) {
          const double* rs = cluster.map_shared_rank(S, c);
          a += rs[D.part_off + i];
        }
        S[D.top_off + i] = a;
      }
      __syncthreads();
      run_section(D, P, S, SEC_TOPF, 1.0, warp, lane);
      run_section(D, P, S, SEC_TOPB, 1.0, warp, lane);
    }
    cluster.sync();
    if (rank != 0) {
      const double* r0 = cluster.map_shared_rank(S, 0);
      for (int i = tid; i < D.n_top; i += blockDim.x) S[D.top_off + i] = r0[D.top_off + i];
      __syncthreads();
    }
  }
  // ---- backward --------------------------------------------------------------------------------
  run_section(D, P, S, SEC_BWD, 1.0, warp, lane);
  if (MODE == MODE_BACKSUB) {
    double* u = P.uI + D.I_base;
    const int32_t* om = P.imap + D.ownmap_off;
    for (int k = tid; k < D.n_own; k += blockDim.x) u[om[k]] = S[D.own_off + k];
    if (rank == 0) {
      const int32_t* tm = P.imap + D.topmap_off;
      for (int k = tid; k < D.n_top; k += blockDim.x) u[tm[k]] = S[D.top_off + k];
    }
    if (C > 1) cluster.sync();  // nobody leaves while rank 0's top values may still be read
    return;
  }
  // ---- interface products ----------------------------------------------------------------------
  if ((MODE == MODE_APPLY || MODE == MODE_UNIT) && rank == 0)
    run_section(D, P, S, SEC_GG, -1.0, warp, lane);
  run_section(D, P, S, SEC_OUT, 1.0, warp, lane);
  if (C > 1) cluster.sync();
  if (rank == 0) {
    double* o;
    if (MODE == MODE_UNIT) {
      const int col = P.unit_col0 + (int)blockIdx.y;
      o = (col < D.n_gd) ? P.sd_pool + P.sd_off[D.dom] + (int64_t)col * D.n_gd : nullptr;
    } else {
      o = P.out_loc + D.G_base;
    }
    if (o != nullptr)
      for (int l = tid; l < D.n_gd; l += blockDim.x) {
        double a = S[D.out_off + l];
        for (int c = 1; c < C; ++c) {
          const double* rs = cluster.map_shared_rank(S, c);
          a += rs[D.out_off + l];
        }
        o[l] = a;
      }
  }
  if (C > 1) cluster.sync();
}

// ---------------------------------------------------------------------------------------------
// Batched dense GEMV over subdomains:  out_loc[G_base+i] = w_i * sum_j M_d[i,j] * (w_j * x[l2g[j]])
// with w = 1/cnt (Neumann-Neumann) or 1 (assembled local Schur).  M_d row-major.
// grid = (row tiles, subdomains)
// ---------------------------------------------------------------------------------------------
struct GemvParams {
  const double* mats;
  const int64_t* mat_off;   // per subdomain
  const int32_t* n_gd;      // per subdomain
  const int64_t* G_base;    // per subdomain
  const int32_t* l2g;       // pool (concatenated, indexed by G_base)
  const double* x;          // interface vector
  const double* inv_cnt;    // 1/node_gamma_cnt or nullptr
  double* out_loc;
  const int* done_at;
  int iter;
};
constexpr int kGemvRowsPerCta = 32;

__global__ void __launch_bounds__(256) dense_gemv_kernel(const GemvParams P) {
  extern __shared__ double xs[];
  if (P.done_at != nullptr && *P.done_at < P.iter) return;
  const int d = blockIdx.y;
  const int n = P.n_gd[d];
  const int row0 = blockIdx.x * kGemvRowsPerCta;
  if (row0 >= n) return;
  const int64_t gb = P.G_base[d];
  const int32_t* l2g = P.l2g + gb;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int g = l2g[j];
    double v = P.x[g];
    if (P.inv_cnt != nullptr) v *= P.inv_cnt[g];
    xs[j] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* M = P.mats + P.mat_off[d];
  for (int r = warp; r < kGemvRowsPerCta; r += 8) {
    const int i = row0 + r;
    if (i >= n) break;
    const double* mrow = M + (int64_t)i * n;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = lane;
    for (; j + 96 < n; j += 128) {
      const double m0 = __ldg(mrow + j), m1 = __ldg(mrow + j + 32), m2 = __ldg(mrow + j + 64),
                   m3 = __ldg(mrow + j + 96);
      a0 = fma(m0, xs[j], a0);
      a1 = fma(m1, xs[j + 32], a1);
      a2 = fma(m2, xs[j + 64], a2);
      a3 = fma(m3, xs[j + 96], a3);
    }
    for (; j < n; j += 32) a0 = fma(__ldg(mrow + j), xs[j], a0);
    double a = (a0 + a1) + (a2 + a3);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      if (P.inv_cnt != nullptr) a *= P.inv_cnt[l2g[i]];
      P.out_loc[gb + i] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Block-level deterministic reductions
// ---------------------------------------------------------------------------------------------
constexpr int kVecThreads = 256;

__device__ __forceinline__ double block_reduce_sum(double v, double* red /*>=8*/) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

// Sum of `nb` per-block partials in a fixed order; every block obtains the identical value.
__device__ __forceinline__ double sum_partials(const double* part, int nb, double* red) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) v += part[i];
  return block_reduce_sum(v, red);
}

// Device-resident scalar state of one PCG recurrence (cg.jl:76-90).
// `done_at` is the index of the iteration whose end satisfied the stopping test (INT_MAX while
// running); kernels of iteration k are skipped when done_at < k, so blocks of the kernel that sets
// it still complete their own work.
struct PcgState {
  double rTz, rTz_old, d, rTr, alpha, beta, tol, bnorm;
  int it, done_at, maxit, fail;
};

// y[g] = base[g] (optional) + sum over the subdomains touching g, ascending subdomain order
// (the order of the reference's sequential loop, EllipticPdeDomainDecomposition.jl:740-742).
// Optionally accumulates the block partial of dot(y, q).
__global__ void __launch_bounds__(kVecThreads)
assemble_kernel(int n, const int32_t* adj_ptr, const int32_t* adj_src, const double* loc, const double* base,
                double* y, const double* q, double* part, const int* done_at, int iter) {
  __shared__ double red[8];
  if (done_at != nullptr && *done_at < iter) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double prod = 0.0;
  if (g < n) {
    double a = (base != nullptr) ? base[g] : 0.0;
    for (int k = adj_ptr[g]; k < adj_ptr[g + 1]; ++k) a += loc[adj_src[k]];
    y[g] = a;
    if (q != nullptr) prod = a * q[g];
  }
  if (part != nullptr) {
    const double s = block_reduce_sum(prod, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
}

// part[b] = block partial of dot(a, b)
__global__ void __launch_bounds__(kVecThreads)
dot_kernel(int n, const double* a, const double* b, double* part, const int* done_at, int iter) {
  __shared__ double red[8];
  if (done_at != nullptr && *done_at < iter) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const double prod = (g < n) ? a[g] * b[g] : 0.0;
  const double s = block_reduce_sum(prod, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// r = b - Ax ; partials of r.r (partA) and b.b (partB)                            (cg.jl:83-84, 89)
__global__ void __launch_bounds__(kVecThreads)
pcg_init_residual_kernel(int n, const double* b, const double* Ax, double* r, double* partA, double* partB) {
  __shared__ double red[8];
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double rr = 0.0, bb = 0.0;
  if (g < n) {
    const double rv = b[g] - Ax[g];
    r[g] = rv;
    rr = rv * rv;
    bb = b[g] * b[g];
  }
  const double s1 = block_reduce_sum(rr, red);
  const double s2 = block_reduce_sum(bb, red);
  if (threadIdx.x == 0) {
    partA[blockIdx.x] = s1;
    partB[blockIdx.x] = s2;
  }
}

// p = z ; state <- (rTr, rTz, tol, it = 1, res_norm[0]) ; partC holds partials of r.z  (cg.jl:82-90)
__global__ void __launch_bounds__(kVecThreads)
pcg_init_direction_kernel(int n, int nb, const double* z, double* p, const double* partA, const double* partB,
                          const double* partC, PcgState* st, double* res_norm, double rtol, int maxit) {
  __shared__ double red[8];
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[g] = z[g];
  if (blockIdx.x == 0) {
    const double rTr = sum_partials(partA, nb, red);
    const double bb = sum_partials(partB, nb, red);
    const double rTz = sum_partials(partC, nb, red);
    if (threadIdx.x == 0) {
      st->rTr = rTr;
      st->rTz = rTz;
      st->bnorm = sqrt(bb);
      st->tol = rtol * sqrt(bb);
      st->it = 1;
      st->maxit = maxit;
      st->fail = 0;
      res_norm[0] = sqrt(rTr);
      st->done_at = ((1 < maxit) && (sqrt(rTr) > st->tol)) ? 0x7fffffff : 0;
    }
  }
}

// d = p.Ap ; alpha = rTz / d ; x += alpha p ; r -= alpha Ap ; partB <- partials of r.r   (cg.jl:94-99)
__global__ void __launch_bounds__(kVecThreads)
pcg_update_xr_kernel(int n, int nb, const double* p, const double* Ap, double* x, double* r, const double* partA,
                     double* partB, PcgState* st, int iter) {
  __shared__ double red[8];
  if (st->done_at < iter) return;
  const double d = sum_partials(partA, nb, red);
  const double rTz = st->rTz;
  const double alpha = rTz / d;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double rr = 0.0;
  if (g < n) {
    x[g] = fma(alpha, p[g], x[g]);
    const double rv = fma(-alpha, Ap[g], r[g]);
    r[g] = rv;
    rr = rv * rv;
  }
  const double s = block_reduce_sum(rr, red);
  if (threadIdx.x == 0) {
    partB[blockIdx.x] = s;
    if (blockIdx.x == 0) {
      st->d = d;
      st->alpha = alpha;
      st->rTz_old = rTz;
      if (!(d > 0.0)) st->fail = 1;  // operator not SPD along p: reported, not asserted
    }
  }
}

// rTz' = r.z ; beta = (1/rTz) rTz' ; p = beta p + z ; it += 1 ; res_norm[it] = sqrt(r.r)  (cg.jl:101-105)
// Optionally stores vcol = z / sqrt(rTz') (eigcg.jl:242).
__global__ void __launch_bounds__(kVecThreads)
pcg_update_p_kernel(int n, int nb, const double* z, double* p, const double* partB, const double* partC,
                    PcgState* st, double* res_norm, double* vcol, int iter) {
  __shared__ double red[8];
  if (st->done_at < iter) return;
  const double rTr = sum_partials(partB, nb, red);
  const double rTz_new = sum_partials(partC, nb, red);
  double beta = 1.0 / st->rTz_old;
  beta *= rTz_new;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) {
    p[g] = fma(beta, p[g], z[g]);
    if (vcol != nullptr) vcol[g] = z[g] / sqrt(rTz_new);
  }
  // fields read above by every block (rTz_old, done_at < iter) are not invalidated by the
  // publication below: rTz_old is never written here and done_at is set to iter, not below it.
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int it = st->it + 1;
    st->it = it;
    st->rTr = rTr;
    st->rTz = rTz_new;
    st->beta = beta;
    res_norm[it - 1] = sqrt(rTr);
    if (!((it < st->maxit) && (sqrt(rTr) > st->tol))) st->done_at = iter;
  }
}

}  // namespace kb

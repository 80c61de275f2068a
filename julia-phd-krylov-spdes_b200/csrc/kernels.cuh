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
  const uint8_t* stream;  // packet streams of all CTAs (chunk aligned)
  const uint16_t* chunk_units;  // used 16-byte units of every chunk of the pool (plan.h): only those are copied
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
  // R right-hand sides per launch (template parameter of the kernel): column j of the interface vector at
  // x_gamma + j*ldx, of the interior vectors at bI / uI + j*ldi, of the local-interface output at out_loc + j*ldo
  int64_t ldx, ldi, ldo;
  int unit_col0;          // MODE_UNIT: columns unit_col0 + R*blockIdx.y + j
  int nstages;            // ring depth (<= kMaxStages)
  int prefetch_chunks;    // L2 prefetch distance of the producer warp in chunks (0: off, <= 32)
  int prefetch_mode;      // 0: cp.async.bulk.prefetch.L2 per chunk, 1: prefetch.global.L2 per 128-byte line
  long long* dbg;         // optional: 16 clock64 stamps per CTA (development aid)
  int dbg_flags;          // development aid: 1 = skip packet arithmetic, 2 = skip level barriers
};

constexpr int kMaxStages = 12;
#ifdef KB200_SLICE_PROF
// instrumented build: 16 phase stamps + 4 counters per consumer warp (slots 16..47) + from slot 64 on the clock of
// consumer warp 0 at every section start (stored negated) and after every level barrier
constexpr int kDbgSlots = 320;
#define KB_PROF(x) x
#else
constexpr int kDbgSlots = 64;  // clock stamps per CTA: 16 phase stamps
#define KB_PROF(x)
#endif
// consumer warps per CTA: kWarps (8; two CTAs per SM) or 16 (one CTA per SM, for grids that leave SMs
// with a single CTA anyway); one more warp feeds the ring
constexpr int stream_threads(int nw) { return (nw + 1) * 32; }

// ---- mbarrier / TMA bulk-copy primitives (PTX) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy (TMA, 1-D), completion signalled on `bar` as transaction bytes
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <int NW>
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); }

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; `pdl_trigger` lets OUR
// successor be scheduled early, `pdl_wait` blocks until the predecessor has completed and flushed its
// memory.  Both are no-ops for ordinary launches.  Every block triggers at its very start, so a
// successor only begins to launch once all blocks of this grid are resident (the grid-wide and
// cross-GPU spin barriers of the vector kernels rely on that).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct StreamCtx {
  const CtaDesc* D;
  const uint8_t* gstream;  // this CTA's stream in global memory
  uint8_t* ring;
  double* S;
  double* scratch;
  uint64_t* full_bar;
  uint64_t* empty_bar;
  uint64_t* lvl_bar;
  uint32_t lvl_phase;
  cg::cluster_group* cluster;
  uint32_t stage, phase;   // ring position (same sequence on both sides); phase = parity of the stage's use count
  uint32_t nst;
  // producer warp: sliding window over the chunk-size table (lane l holds the entry of chunk ubase + l in
  // ucur and of chunk ubase + 32 + l in unxt, fetched one window ahead of its use)
  const uint16_t* units;
  uint32_t nchunks, ubase, ucur, unxt;
  int pf_mode;             // 0: bulk prefetches, 1: line prefetches
  uint32_t pf, pf_done;    // L2 prefetch distance in chunks (0: off) / first chunk not yet asked for
  int dbg_flags;
  int warp, lane;
  KB_PROF(long long t_slice; long long n_slice; long long t_bar; long long t_wait; long long* lvlp; int nlvl;)
};

// Producer warp: stream the chunks of one section into the ring.
__device__ __forceinline__ uint32_t window_load(const StreamCtx& cx, uint32_t first) {
  const uint32_t i = first + (uint32_t)cx.lane;
  return i < cx.nchunks ? (uint32_t)__ldg(cx.units + i) : 0u;
}
// L2 prefetch of a byte range (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void produce_section(StreamCtx& cx, int sec) {
  const uint32_t c0 = cx.D->sec_chunk0[sec], c1 = cx.D->sec_chunk1[sec];
  for (uint32_t ch = c0; ch < c1; ++ch) {
    const uint32_t stage = cx.stage;
    // bytes of the chunk that carry data (the padding behind them stays in HBM)
    while (ch >= cx.ubase + 32u) {  // (twice only when a whole window was skipped with its section)
      cx.ubase += 32u;
      cx.ucur = cx.unxt;
      cx.unxt = window_load(cx, cx.ubase + 32u);
    }
    const uint32_t bytes = __shfl_sync(0xffffffffu, cx.ucur, (int)(ch - cx.ubase)) * 16u;
    // L2 PREFETCH.  The CTAs of a launch walk the same program on similar subdomains: they sit in the latency-bound
    // small levels together and ask for the data-heavy levels together, so HBM idles and saturates in turns while
    // the ring (a few chunks) cannot look further ahead.  The producer therefore asks the L2 for the chunks up to
    // `pf` ahead of the one it copies: HBM streams at its own pace into the 126 MB L2, the bulk copies of the
    // data-heavy levels then run at L2 speed.  (The window holds the sizes of chunks ubase .. ubase + 63.)
    if (cx.pf > 0u) {
      if (cx.pf_done < ch + 1u) cx.pf_done = ch + 1u;
      const uint32_t target = min(min(ch + 1u + cx.pf, cx.nchunks), cx.ubase + 64u);
      while (cx.pf_done < target) {
        if (cx.pf_mode == 0) {   // one bulk prefetch per chunk (they share the SM's bulk-copy engine with the real copies)
          const uint32_t mine = cx.pf_done + (uint32_t)cx.lane;
          const uint32_t rel = mine - cx.ubase;     // < 64 + 32
          const uint32_t lo = __shfl_sync(0xffffffffu, cx.ucur, (int)(rel & 31u));
          const uint32_t hi = __shfl_sync(0xffffffffu, cx.unxt, (int)(rel & 31u));
          const uint32_t u = rel < 32u ? lo : hi;
          if (mine < target && u != 0u) prefetch_l2(cx.gstream + (size_t)mine * kChunkBytes, u * 16u);
          cx.pf_done = min(cx.pf_done + 32u, target);
        } else {                 // line prefetches from the producer warp's own load pipe (128 B per lane and instruction)
          const uint32_t rel = cx.pf_done - cx.ubase;   // < 64
          const uint32_t nb = (rel < 32u ? __shfl_sync(0xffffffffu, cx.ucur, (int)rel)
                                         : __shfl_sync(0xffffffffu, cx.unxt, (int)(rel - 32u))) * 16u;
          const uint8_t* src = cx.gstream + (size_t)cx.pf_done * kChunkBytes;
          for (uint32_t o = (uint32_t)cx.lane * 128u; o < nb; o += 32u * 128u)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
          cx.pf_done += 1u;
        }
      }
    }
    // the producer waits for the previous use of the stage to be released: parity phase^1; the very
    // first pass over the ring finds the barriers in their initial phase, which try_wait(1) accepts
    mbar_wait(&cx.empty_bar[stage], cx.phase ^ 1u);
    if (cx.lane == 0) {
      mbar_expect_tx(&cx.full_bar[stage], bytes);
      tma_load_1d(cx.ring + stage * kChunkBytes, cx.gstream + (size_t)ch * kChunkBytes, bytes, &cx.full_bar[stage]);
    }
    __syncwarp();
    if (++cx.stage == cx.nst) {
      cx.stage = 0;
      cx.phase ^= 1u;
    }
  }
}

// Software barrier over the consumer threads of all CTAs of the cluster (the producer warps keep
// streaming): CTA-local barrier, then one thread releases at cluster scope (ONE fence: the membar
// sequence it costs is paid once, not once per peer) and arrives (relaxed) on the level mbarrier of
// every CTA of the cluster; everybody waits on its own CTA's barrier and acquires with the fence
// restricted to shared::cluster memory -- the only data exchanged between levels lives in
// (distributed) shared memory, so no L1 invalidation (CCTL.IVALL) is needed on the polling path.
template <int NW>
__device__ __forceinline__ void cluster_level_sync(StreamCtx& cx) {
  // the remote (DSMEM) stores of all consumer threads happen before thread 0's release fence through
  // the CTA barrier (the release pattern is cumulative)
  consumer_barrier<NW>();
  const uint32_t local = smem_u32(cx.lvl_bar);
  if (cx.warp == 0 && cx.lane == 0) {
    asm volatile("fence.release.cluster;" ::: "memory");
    for (int c = 0; c < cx.D->csize; ++c) {
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(c));
      asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    }
  }
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(local), "r"(cx.lvl_phase)
        : "memory");
  } while (!ok);
  asm volatile("fence.acquire.sync_restrict::shared::cluster.cluster;" ::: "memory");
  cx.lvl_phase ^= 1u;
}

// One slice (see plan.h): T lanes per row, entries compacted per step.  Written for a short
// instruction stream -- at ~270 entries per slice the fixed part (header decode, tail, butterfly,
// row update) weighs as much as the entry loop: everything the header can precompute is precomputed
// on the host, the row record and the old destination value are fetched before the loop, the loop is
// kept at 4 steps per trip (no wide unrolling with its set-up code) and the tail is 2 + 1 steps.
// R right-hand sides at once (MULTI-VECTOR kernel, the batched SpMM apply of
// Example11_EigInitPcgEllipticPdeMultipleRhs_Functions.jl:140-179 / eigcg.jl:212-218): the shared vector holds
// R interleaved values per slot, S[slot * R + j]; a stream entry (value, 16-bit slot) is read ONCE and applied
// to the R columns, every column with exactly the association order of the single-vector kernel (so column j
// of a batched apply has the bits of a single apply of that column).
// the R values of slot c (16-byte vector loads: S is 16-byte aligned and R is 1, 2 or 4)
template <int R>
__device__ __forceinline__ void load_slot(const double* S, uint32_t c, double (&v)[R]) {
  if (R == 1) {
    v[0] = S[c];
  } else {
#pragma unroll
    for (int h = 0; h < R / 2; ++h) {
      const double2 t = reinterpret_cast<const double2*>(S)[c * (R / 2) + h];
      v[2 * h] = t.x;
      v[2 * h + 1] = t.y;
    }
  }
}

template <bool kTop, int R>
__device__ __forceinline__ void process_slice(StreamCtx& cx, const uint8_t* sl, double sign) {
  double* S = cx.S;
  const int lane = cx.lane;
  const uint4 hq = *reinterpret_cast<const uint4*>(sl);
  const int kmax = (int)(hq.x >> 16);
  const int tl = (int)(hq.y & 0xffu), nrows = (int)((hq.y >> 8) & 0xffu);
  const int W = (int)hq.z;                       // nrows << tl active lanes
  const int g = min(lane >> tl, nrows - 1);      // row of this lane (clamped: idle lanes read a valid record)
  const bool writer = lane < W && (lane & ((1 << tl) - 1)) == 0;
  const uint4 q = *reinterpret_cast<const uint4*>(sl + 16 + 16 * g);
  const uint32_t dst = q.x >> 16, flags = q.y & 0xffffu;
  double old[R];
  load_slot<R>(S, dst, old);
  if (flags & kRowFresh) {
#pragma unroll
    for (int j = 0; j < R; ++j) old[j] = 0.0;
  }
  // entry (k, lane) at k * W + lane; padding entries are (0.0, col 0)
  const double* vp = reinterpret_cast<const double*>(sl + 16 + 16 * nrows) + lane;
  // columns: one 32-bit word per lane and PAIR of steps (low half = even step): half the shared-memory
  // wavefronts and load instructions of a 16-bit load per step
  const uint32_t* cp = reinterpret_cast<const uint32_t*>(sl + hq.w) + lane;
  double a0[R], a1[R], a2[R], a3[R];
#pragma unroll
  for (int j = 0; j < R; ++j) a0[j] = a1[j] = a2[j] = a3[j] = 0.0;
  if (lane < W) {
    int k = kmax;
#pragma unroll 1
    for (; k >= 4; k -= 4) {
      const double v0 = vp[0], v1 = vp[W], v2 = vp[2 * W], v3 = vp[3 * W];
      const uint32_t cc0 = cp[0], cc1 = cp[W];
      const uint32_t c0 = cc0 & 0xffffu, c1 = cc0 >> 16, c2 = cc1 & 0xffffu, c3 = cc1 >> 16;
      double s0[R], s1[R], s2[R], s3[R];
      load_slot<R>(S, c0, s0);
      load_slot<R>(S, c1, s1);
      load_slot<R>(S, c2, s2);
      load_slot<R>(S, c3, s3);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        a0[j] = fma(v0, s0[j], a0[j]);
        a1[j] = fma(v1, s1[j], a1[j]);
        a2[j] = fma(v2, s2[j], a2[j]);
        a3[j] = fma(v3, s3[j], a3[j]);
      }
      vp += 4 * W;
      cp += 2 * W;
    }
    // association of the sums: step j feeds accumulator j & 3 (the emulator follows the same rule)
    if (k & 2) {
      const double v0 = vp[0], v1 = vp[W];
      const uint32_t cc0 = cp[0];
      const uint32_t c0 = cc0 & 0xffffu, c1 = cc0 >> 16;
      double s0[R], s1[R];
      load_slot<R>(S, c0, s0);
      load_slot<R>(S, c1, s1);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        a0[j] = fma(v0, s0[j], a0[j]);
        a1[j] = fma(v1, s1[j], a1[j]);
      }
      vp += 2 * W;
      cp += W;
    }
    if (k & 1) {
      const double v0 = vp[0];
      const uint32_t c0 = cp[0] & 0xffffu;
      double s0[R];
      load_slot<R>(S, c0, s0);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (k & 2) a2[j] = fma(v0, s0[j], a2[j]);
        else a0[j] = fma(v0, s0[j], a0[j]);
      }
    }
  }
  const double dg = __hiloint2double((int)q.w, (int)q.z);  // 1 / diag
#pragma unroll
  for (int j = 0; j < R; ++j) {
    double acc = (a0[j] + a1[j]) + (a2[j] + a3[j]);
    if (tl >= 1) acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (tl >= 2) acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (tl >= 3) acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (tl >= 4) acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    if (tl >= 5) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (writer) {
      const double r = (old[j] - sign * acc) * dg;
      S[dst * R + j] = r;
      if (kTop && (flags & kRowBcast)) {
        for (int cc = 0; cc < cx.D->csize; ++cc)
          if (cc != cx.D->rank) cx.cluster->map_shared_rank(S, cc)[dst * R + j] = r;
      }
    }
  }
}

// Consumer warps: run the warp programs of one section out of the ring (see plan.h).  A warp reads
// only its own token list of every chunk: a token is a slice to process or the end of a level
// (CTA barrier; TOP sections (kTop): barrier over the whole cluster, the rows of a level having been
// shared out over the CTAs and their results stored into every CTA's replica of the top vector).
template <bool kTop, int NW, int R>
__device__ __forceinline__ void consume_section(StreamCtx& cx, int sec, double sign) {
  const uint32_t c0 = cx.D->sec_chunk0[sec], c1 = cx.D->sec_chunk1[sec];
  const int lane = cx.lane, warp = cx.warp;
  KB_PROF(if (cx.lvlp != nullptr && warp == 0 && lane == 0 && cx.nlvl < kDbgSlots - 64) cx.lvlp[cx.nlvl++] = -clock64();)
  for (uint32_t ch = c0; ch < c1; ++ch) {
    const uint32_t stage = cx.stage;
    KB_PROF(const long long tw0 = clock64();)
    mbar_wait(&cx.full_bar[stage], cx.phase);
    KB_PROF(cx.t_wait += clock64() - tw0;)
    const uint8_t* chunk = cx.ring + stage * kChunkBytes;
    const uint32_t t0 = reinterpret_cast<const uint16_t*>(chunk)[warp];
    const uint32_t dnext = reinterpret_cast<const uint16_t*>(chunk)[warp + 1];
    const uint16_t* tok = reinterpret_cast<const uint16_t*>(chunk + kChunkDirBytes);
    for (uint32_t i = t0; i < dnext; ++i) {
      const uint32_t t = tok[i];
      if (t == kTokBarrier) {
#ifdef KB200_SLICE_PROF
        if (kTop) cluster_level_sync<NW>(cx);
        else if (!(cx.dbg_flags & 2)) consumer_barrier<NW>();
        if (cx.lvlp != nullptr && warp == 0 && lane == 0 && cx.nlvl < kDbgSlots - 64) cx.lvlp[cx.nlvl++] = clock64();
#else
        if (kTop) cluster_level_sync<NW>(cx);
        else consumer_barrier<NW>();
#endif
      } else {
#ifdef KB200_SLICE_PROF
        if (cx.dbg_flags & 1) continue;
#endif
        KB_PROF(const long long ts0 = clock64();)
        process_slice<kTop, R>(cx, chunk + t * 16u, sign);
        __syncwarp();  // the warp's next slice may read these rows (warp-private subtrees)
        KB_PROF(cx.t_slice += clock64() - ts0; cx.n_slice++;)
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&cx.empty_bar[stage]);
    if (++cx.stage == cx.nst) {
      cx.stage = 0;
      cx.phase ^= 1u;
    }
  }
}

template <int MODE, int NW, int R>
__global__ void __launch_bounds__(stream_threads(NW), (NW == kWarps && R == 1) ? 2 : 1) schur_stream_kernel(const StreamParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ CtaDesc D;
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t lvl_bar;
  pdl_trigger();
  cg::cluster_group cluster = cg::this_cluster();
  double* scratch = reinterpret_cast<double*>(smem_raw + (size_t)P.nstages * kChunkBytes);
  double* S = scratch + kMaxSlots * R;   // slot i of column j at S[i * R + j]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.ctas + blockIdx.x);
    uint32_t* dstw = reinterpret_cast<uint32_t*>(&D);
    for (int i = tid; i < (int)(sizeof(CtaDesc) / 4); i += blockDim.x) dstw[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < P.nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NW);
    }
    mbar_init(&lvl_bar, (uint32_t)P.ctas[blockIdx.x].csize);
    mbar_fence_init();
  }
  __syncthreads();
  for (int i = tid; i < D.smem_doubles * R; i += blockDim.x) S[i] = 0.0;
  // everything above touches only static data and this CTA's shared memory: with PDL it overlaps the
  // tail of the previous kernel; from here on the kernel reads what its predecessors wrote
  pdl_wait();
  if (P.done_at != nullptr && *P.done_at < P.iter) return;  // uniform over the grid (and the cluster)
  __syncthreads();
  const int C = D.csize, rank = D.rank;
  // ---- inputs ---------------------------------------------------------------------------------
  if (MODE == MODE_APPLY || MODE == MODE_BACKSUB) {
    const int32_t* l2g = P.l2g + D.l2g_off;
    for (int l = tid; l < D.n_gd; l += blockDim.x) {
      const int32_t g = l2g[l];
#pragma unroll
      for (int j = 0; j < R; ++j) S[(D.xd_off + l) * R + j] = P.x_gamma[g + j * P.ldx];
    }
  }
  if (MODE == MODE_UNIT) {
    const int col0 = P.unit_col0 + (int)blockIdx.y * R;
    if (tid < R && col0 + tid < D.n_gd) S[(D.xd_off + col0 + tid) * R + tid] = 1.0;
  }
  if (MODE == MODE_RHS || MODE == MODE_BACKSUB) {
    const double* b = P.bI + D.I_base;
    const int32_t* om = P.imap + D.ownmap_off;
    for (int k = tid; k < D.n_own; k += blockDim.x) {
      const int32_t o = om[k];
#pragma unroll
      for (int j = 0; j < R; ++j) S[(D.own_off + k) * R + j] = b[o + j * P.ldi];
    }
    if (rank == 0) {
      const int32_t* tm = P.imap + D.topmap_off;
      for (int k = tid; k < D.n_top; k += blockDim.x) {
        const int32_t o = tm[k];
#pragma unroll
        for (int j = 0; j < R; ++j) S[(D.top_off + k) * R + j] = b[o + j * P.ldi];
      }
    }
  }
  __syncthreads();

  StreamCtx cx;
  cx.D = &D;
  cx.gstream = P.stream + D.stream_off;
  cx.ring = smem_raw;
  cx.S = S;
  cx.scratch = scratch;
  cx.full_bar = full_bar;
  cx.empty_bar = empty_bar;
  cx.lvl_bar = &lvl_bar;
  cx.lvl_phase = 0;
  cx.cluster = &cluster;
  cx.stage = 0;
  cx.phase = 0;
  cx.nst = (uint32_t)P.nstages;
  cx.units = P.chunk_units + D.stream_off / kChunkBytes;
  cx.nchunks = D.sec_chunk1[NSEC - 1];
  cx.ubase = 0;
  cx.ucur = cx.unxt = 0;
  cx.pf = (uint32_t)P.prefetch_chunks;
  cx.pf_mode = P.prefetch_mode;
  cx.pf_done = 0;
  cx.dbg_flags = P.dbg_flags;
  cx.warp = warp;
  cx.lane = lane;
  KB_PROF(cx.t_slice = 0; cx.n_slice = 0; cx.t_bar = 0; cx.t_wait = 0; cx.nlvl = 0;
          cx.lvlp = P.dbg != nullptr ? P.dbg + (size_t)blockIdx.x * kDbgSlots + 64 : nullptr;)
  const bool producer = warp == NW;
  if (producer) {
    cx.ucur = window_load(cx, 0u);
    cx.unxt = window_load(cx, 32u);
  }
  auto stamp = [&](int i) { if (P.dbg != nullptr && tid == 0) P.dbg[(size_t)blockIdx.x * kDbgSlots + i] = clock64(); };
  stamp(0);
  constexpr bool kHasInit = MODE == MODE_APPLY || MODE == MODE_UNIT || MODE == MODE_BACKSUB;
  constexpr double kInitSign = (MODE == MODE_BACKSUB) ? 1.0 : -1.0;
  constexpr bool kHasGG = MODE == MODE_APPLY || MODE == MODE_UNIT;
  constexpr bool kHasOut = MODE != MODE_BACKSUB;

  // ---- phase 1: SpMV, forward sweep on the own subtrees, partial products for the top -------------
  if (producer) {
    if (kHasInit) produce_section(cx, SEC_INIT);
    produce_section(cx, SEC_FWD);
    if (C > 1) produce_section(cx, SEC_PART);
  } else {
    if (kHasInit) consume_section<false, NW, R>(cx, SEC_INIT, kInitSign);
    stamp(9);
    consume_section<false, NW, R>(cx, SEC_FWD, 1.0);
    stamp(10);
    if (C > 1) consume_section<false, NW, R>(cx, SEC_PART, 1.0);
  }
  stamp(1);
  // ---- phase 2: top nodes, shared out over the cluster ----------------------------------------------
  if (C > 1) {
    cluster.sync();  // partial products of every CTA are complete (all threads incl. producers)
    stamp(2);
    if (producer) {
      produce_section(cx, SEC_TOPF);
      produce_section(cx, SEC_TOPB);
    } else {
      // every CTA forms the same right-hand side of the top system, b_top (held by rank 0) plus
      // the partial products of all CTAs, first into the scratch, and only overwrites its replica of
      // the top vector once all CTAs have finished reading (two cluster-wide barriers per pass)
      const double* r0 = cluster.map_shared_rank(S, 0);
      const int ntopR = D.n_top * R;                 // (node, column) pairs, interleaved like S
      const int topR = D.top_off * R, partR = D.part_off * R;
      for (int base = 0; base < ntopR; base += kMaxSlots * R) {
        for (int i = base + tid; i < ntopR && i < base + kMaxSlots * R; i += NW * 32) {
          double a = r0[topR + i];
          for (int c = 0; c < C; ++c) {
            const double* rs = cluster.map_shared_rank(S, c);
            a += rs[partR + i];
          }
          scratch[i - base] = a;
        }
        cluster_level_sync<NW>(cx);
        for (int i = base + tid; i < ntopR && i < base + kMaxSlots * R; i += NW * 32) S[topR + i] = scratch[i - base];
        cluster_level_sync<NW>(cx);
      }
      consume_section<true, NW, R>(cx, SEC_TOPF, 1.0);
      stamp(3);
      consume_section<true, NW, R>(cx, SEC_TOPB, 1.0);
    }
    stamp(4);
    stamp(5);
  }
  // ---- phase 3: backward sweep, interface products ---------------------------------------------------
  if (producer) {
    produce_section(cx, SEC_BWD);
    if (kHasGG && rank == 0) produce_section(cx, SEC_GG);
    if (kHasOut) produce_section(cx, SEC_OUT);
  } else {
    consume_section<false, NW, R>(cx, SEC_BWD, 1.0);
    stamp(6);
    if (kHasGG && rank == 0) consume_section<false, NW, R>(cx, SEC_GG, -1.0);
    if (kHasOut) consume_section<false, NW, R>(cx, SEC_OUT, 1.0);
  }
  __syncthreads();
  stamp(7);
  KB_PROF(if (P.dbg != nullptr && lane == 0 && warp < 8) {
    long long* o = P.dbg + (size_t)blockIdx.x * kDbgSlots + 16 + warp * 4;
    o[0] = cx.t_slice; o[1] = cx.n_slice; o[2] = cx.t_bar; o[3] = cx.t_wait;
  })
  if (MODE == MODE_BACKSUB) {
    double* u = P.uI + D.I_base;
    const int32_t* om = P.imap + D.ownmap_off;
    for (int k = tid; k < D.n_own; k += blockDim.x) {
      const int32_t o = om[k];
#pragma unroll
      for (int j = 0; j < R; ++j) u[o + j * P.ldi] = S[(D.own_off + k) * R + j];
    }
    if (rank == 0) {
      const int32_t* tm = P.imap + D.topmap_off;
      for (int k = tid; k < D.n_top; k += blockDim.x) {
        const int32_t o = tm[k];
#pragma unroll
        for (int j = 0; j < R; ++j) u[o + j * P.ldi] = S[(D.top_off + k) * R + j];
      }
    }
    if (C > 1) cluster.sync();  // nobody leaves while rank 0's top values may still be read
    return;
  }
  if (C > 1) cluster.sync();
  if (rank == 0) {
    const int col0 = P.unit_col0 + (int)blockIdx.y * R;
    for (int l = tid; l < D.n_gd; l += blockDim.x) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        double a = S[(D.out_off + l) * R + j];
        for (int c = 1; c < C; ++c) {
          const double* rs = cluster.map_shared_rank(S, c);
          a += rs[(D.out_off + l) * R + j];
        }
        if (MODE == MODE_UNIT) {
          if (col0 + j < D.n_gd) P.sd_pool[P.sd_off[D.dom] + (int64_t)(col0 + j) * D.n_gd + l] = a;
        } else {
          P.out_loc[D.G_base + l + j * P.ldo] = a;
        }
      }
    }
  }
  if (C > 1) cluster.sync();
  stamp(8);
}

// ---------------------------------------------------------------------------------------------
// Numeric Cholesky with a fixed symbolic pattern, batched over the subdomains (next row f1: the re-factorisation of
// A_II,d for a new coefficient realization, Example07_PcgSchurStochasticEllipticPde.jl:160-204 calls cholesky() per
// realization; EllipticPdeDomainDecomposition.jl:389-546 refills A_II,d).  Dot-product (up-looking) form on the pattern
// of the uploaded factor: L[i,j] = (A[i,j] - sum_{k<j} L[i,k] L[j,k]) / L[j,j], L[j,j] = sqrt(A[j,j] - sum_k L[j,k]^2).
// Columns are grouped into dependency levels (level(j) = 1 + max level of the columns row j touches); ONE launch per
// level handles the entries of that level's columns of all subdomains, one thread per entry: it walks row j of the
// factor (CSR view of the same values) and finds the matching entries of row i with a monotone scan (the pattern of
// row i contains that of row j below j -- fill).  Every sum runs in ascending column order: deterministic.
// All index arrays are global over the local subdomains (rows of all subdomains numbered consecutively).
// ---------------------------------------------------------------------------------------------
struct CholParams {
  const int32_t* order;    // entries of the factor sorted by the level of their column
  const int32_t* Lrow;     // per entry: (global) row i
  const int32_t* Lcol;     // per entry: (global) column j
  const uint32_t* epos;    // per entry: index of its value in the pool
  const int32_t* asrc;     // per entry: index of A[i,j] in the staged values, -1 where A has no entry
  const int32_t* rptr;     // CSR of the factor pattern: row pointers (global rows), columns ascending, diagonal last
  const int32_t* rcol;     // (global) column of every CSR entry
  const uint32_t* rpos;    // pool index of every CSR entry
  const int32_t* diag_e;   // per (global) row j: entry id of (j, j)
  const double* A;         // staged values of A_II of all local subdomains
  double* pool;
  int* fail;               // set to 1 when a pivot is not positive
};
__global__ void __launch_bounds__(256) chol_level_kernel(const CholParams P, int64_t e0, int64_t e1) {
  const int64_t idx = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= e1) return;
  const int32_t E = P.order[idx];
  const int32_t i = P.Lrow[E], j = P.Lcol[E];
  const int32_t qj0 = P.rptr[j], qj1 = P.rptr[j + 1] - 1;   // row j below the diagonal: [qj0, qj1)
  double sjj = P.A[P.asrc[P.diag_e[j]]];
  for (int32_t q = qj0; q < qj1; ++q) {
    const double v = P.pool[P.rpos[q]];
    sjj = fma(-v, v, sjj);
  }
  if (!(sjj > 0.0)) {
    *P.fail = 1;
    sjj = 1.0;
  }
  const double ljj = sqrt(sjj);
  if (i == j) {
    P.pool[P.epos[E]] = ljj;
    return;
  }
  const int32_t a = P.asrc[E];
  double sij = a >= 0 ? P.A[a] : 0.0;
  int32_t p = P.rptr[i];
  const int32_t pend = P.rptr[i + 1];
  for (int32_t q = qj0; q < qj1; ++q) {
    const int32_t k = P.rcol[q];
    while (p < pend && P.rcol[p] < k) ++p;
    if (p < pend && P.rcol[p] == k) sij = fma(-P.pool[P.rpos[p]], P.pool[P.rpos[q]], sij);
  }
  P.pool[P.epos[E]] = sij / ljj;
}
// The same for the entries of columns whose row is long (the separator rows: hundreds of entries): one WARP per entry,
// the lanes stride over row j and find their partner in row i by binary search; lane sums in ascending order, then a
// butterfly -- deterministic.  (One thread per entry walks such rows one dependent load at a time: the ~130 levels of
// the top separators took 0.3 s of a 4M-DOF re-factorisation.)
__global__ void __launch_bounds__(256) chol_level_warp_kernel(const CholParams P, int64_t e0, int64_t e1) {
  const int64_t idx = e0 + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (idx >= e1) return;   // (whole warps leave together)
  const int32_t E = P.order[idx];
  const int32_t i = P.Lrow[E], j = P.Lcol[E];
  const int32_t qj0 = P.rptr[j], qj1 = P.rptr[j + 1] - 1;
  const int32_t pi0 = P.rptr[i], pend = P.rptr[i + 1];
  double pjj = 0.0, pij = 0.0;
  for (int32_t q = qj0 + lane; q < qj1; q += 32) {
    const double v = P.pool[P.rpos[q]];
    pjj = fma(v, v, pjj);
    if (i != j) {
      const int32_t k = P.rcol[q];
      int32_t lo = pi0, hi = pend;   // first position of row i with column >= k
      while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (P.rcol[mid] < k) lo = mid + 1;
        else hi = mid;
      }
      if (lo < pend && P.rcol[lo] == k) pij = fma(P.pool[P.rpos[lo]], v, pij);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    pjj += __shfl_xor_sync(0xffffffffu, pjj, o);
    pij += __shfl_xor_sync(0xffffffffu, pij, o);
  }
  if (lane != 0) return;
  double sjj = P.A[P.asrc[P.diag_e[j]]] - pjj;
  if (!(sjj > 0.0)) {
    *P.fail = 1;
    sjj = 1.0;
  }
  const double ljj = sqrt(sjj);
  if (i == j) {
    P.pool[P.epos[E]] = ljj;
    return;
  }
  const int32_t a = P.asrc[E];
  P.pool[P.epos[E]] = ((a >= 0 ? P.A[a] : 0.0) - pij) / ljj;
}

// ---------------------------------------------------------------------------------------------
// Re-assembly of the local blocks for a new coefficient (kb200_assemble_realization): the element loop of
// prepare_local_schurs, EllipticPdeDomainDecomposition.jl:430-497, in gather form.  One thread per stored entry of a
// block; its contributions are positions 9 e + 3 i + j of the reference's loop nest (element e, vertices i, j) in
// ascending order = the order in which sparse() (:540-543) adds the duplicates.  Every operation is the reference's,
// in its association order, with explicitly rounded multiplies / adds (no contraction into fused multiply-adds, which
// Julia does not apply either): the values have the bits of the host assembly.
// ---------------------------------------------------------------------------------------------
struct AsmParams {
  const int32_t* cells;   // 3 x nel, 0-based node ids, column-major like the reference's `cells`
  const double* pts;      // 2 x nnode
  const double* coeff;    // nnode
  const int64_t* ptr;     // per entry: its contributions [ptr[s], ptr[s + 1])
  const int32_t* src;     // 9 e + 3 i + j, 0-based
  const uint32_t* dst;    // per entry: index in `out` (nullptr: s itself)
  double* out;
  int64_t n;
};
__device__ __forceinline__ double pick3(int k, double a0, double a1, double a2) { return k == 0 ? a0 : (k == 1 ? a1 : a2); }
// ΔK_ij of contribution c = 9 e + 3 i + j (:437-470), every operation rounded on its own
__device__ __forceinline__ double element_stiffness(const int32_t* __restrict__ cells, const double* __restrict__ pts,
                                                    const double* __restrict__ coeff, int32_t c) {
  const int32_t e = c / 9, ij = c - 9 * e, i = ij / 3, j = ij - 3 * i;
  const int32_t n0 = cells[3 * (int64_t)e], n1 = cells[3 * (int64_t)e + 1], n2 = cells[3 * (int64_t)e + 2];
  const double x0 = pts[2 * (int64_t)n0], y0 = pts[2 * (int64_t)n0 + 1];
  const double x1 = pts[2 * (int64_t)n1], y1 = pts[2 * (int64_t)n1 + 1];
  const double x2 = pts[2 * (int64_t)n2], y2 = pts[2 * (int64_t)n2 + 1];
  // :437-446   Δa = 0.; Δa += coeff[jnode] (three times); Δa /= 3.
  double da = __dadd_rn(__dadd_rn(__dadd_rn(0.0, coeff[n0]), coeff[n1]), coeff[n2]);
  da = __ddiv_rn(da, 3.0);
  // :449-454   (1-based Δx[1..3], Δy[1..3] of the reference = dx0..2, dy0..2)
  const double dx0 = __dsub_rn(x2, x1), dx1 = __dsub_rn(x0, x2), dx2 = __dsub_rn(x1, x0);
  const double dy0 = __dsub_rn(y1, y2), dy1 = __dsub_rn(y2, y0), dy2 = __dsub_rn(y0, y1);
  // :457       Area = (Δx[3] * Δy[2] - Δx[2] * Δy[3]) / 2.
  const double area = __ddiv_rn(__dsub_rn(__dmul_rn(dx2, dy1), __dmul_rn(dx1, dy2)), 2.0);
  // :470       ΔKij = Δa * (Δy[i] * Δy[j] + Δx[i] * Δx[j]) / 4 / Area
  const double dyi = pick3(i, dy0, dy1, dy2), dyj = pick3(j, dy0, dy1, dy2);
  const double dxi = pick3(i, dx0, dx1, dx2), dxj = pick3(j, dx0, dx1, dx2);
  const double g = __dadd_rn(__dmul_rn(dyi, dyj), __dmul_rn(dxi, dxj));
  return __ddiv_rn(__ddiv_rn(__dmul_rn(da, g), 4.0), area);
}
__global__ void __launch_bounds__(256) assemble_values_kernel(const AsmParams P) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.n) return;
  double acc = 0.0;
  const int64_t q0 = P.ptr[s], q1 = P.ptr[s + 1];
  for (int64_t q = q0; q < q1; ++q) {
    const double k = element_stiffness(P.cells, P.pts, P.coeff, P.src[q]);
    acc = q == q0 ? k : __dadd_rn(acc, k);
  }
  P.out[P.dst ? P.dst[s] : s] = acc;
}
// Right-hand sides b_I,d / b_Γ of the same loop (:499-536) for the entries that carry Dirichlet-lift terms (the nodes next
// to the boundary; every other entry is the coefficient-independent load vector, copied from a static image first).
// One thread per entry; its terms in the reference's order: src >= 0: `b -= ΔK_c * uexact` (:502 / :506, val = uexact at
// the Dirichlet vertex), src < 0: `b += Δb` (:529 / :533, val = Δb of :524-526).
struct RhsParams {
  const int32_t* cells;
  const double* pts;
  const double* coeff;
  const int64_t* ptr;     // per entry: its terms [ptr[t], ptr[t + 1])
  const int32_t* src;
  const double* val;
  const uint32_t* dst;    // per entry: index in `out`
  double* out;
  int64_t n;
};
__global__ void __launch_bounds__(128) assemble_rhs_kernel(const RhsParams P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n) return;
  double acc = 0.0;
  for (int64_t q = P.ptr[t]; q < P.ptr[t + 1]; ++q) {
    const int32_t c = P.src[q];
    if (c < 0) acc = __dadd_rn(acc, P.val[q]);
    else acc = __dsub_rn(acc, __dmul_rn(element_stiffness(P.cells, P.pts, P.coeff, c), P.val[q]));
  }
  P.out[P.dst[t]] = acc;
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
  int rows_per_cta;         // 32 (kGemvRowsPerCta) unless overridden by the gemv_rows option
};
constexpr int kGemvRowsPerCta = 32;

// the 16 loads of one 512-column stretch of a row, issued back to back (clamped addresses: a load beyond the row
// end is multiplied by zero later)
__device__ __forceinline__ void gemv_load16(const double* mrow, int j0, int lane, int n, double (&m)[16]) {
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    // unconditional volatile load: left to the compiler, each load is sunk next to its FMA and the row is read one
    // latency at a time
    const int j = min(j0 + lane + 32 * u, n - 1);
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(m[u]) : "l"(mrow + j));
  }
}

__global__ void __launch_bounds__(256, 4) dense_gemv_kernel(const GemvParams P) {
  extern __shared__ double xs[];
  pdl_trigger();
  pdl_wait();
  if (P.done_at != nullptr && *P.done_at < P.iter) return;
  const int d = blockIdx.y;
  const int n = P.n_gd[d];
  const int rpc = P.rows_per_cta;
  const int row0 = blockIdx.x * rpc;
  if (row0 >= n) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gb = P.G_base[d];
  const int32_t* l2g = P.l2g + gb;
  const double* M = P.mats + P.mat_off[d];
  // The matrix does not depend on x: the first row of every warp is requested BEFORE the gather of x_d (two
  // dependent loads and a barrier), so the DRAM latency of the row and the staging of the vector overlap -- on the
  // per-GPU shards of a multi-GPU run (a few rows per warp) that is most of the kernel.
  double m[16];
  const bool have0 = row0 + warp < n && warp < rpc;
  if (have0) gemv_load16(M + (int64_t)(row0 + warp) * n, 0, lane, n, m);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int g = l2g[j];
    double v = P.x[g];
    if (P.inv_cnt != nullptr) v *= P.inv_cnt[g];
    xs[j] = v;
  }
  __syncthreads();
  for (int r = warp; r < rpc; r += 8) {
    const int i = row0 + r;
    if (i >= n) break;
    const double* mrow = M + (int64_t)i * n;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    // 16 loads per lane in flight (a whole row of up to 512 entries at once): the kernel is bound by memory
    // latency, not by arithmetic; element j = j0 + lane + 32 u feeds accumulator u & 3 (fixed order)
    for (int j0 = 0; j0 < n; j0 += 512) {
      if (r != warp || j0 != 0) gemv_load16(mrow, j0, lane, n, m);   // (the first stretch of the first row is already here)
      __syncwarp();  // keeps the arithmetic below the block of loads
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        const int j = j0 + lane + 32 * u;
        // (a clamped load beyond the row end multiplies a zero)
        a0 = fma(m[u], (j < n) ? xs[j] : 0.0, a0);
        a1 = fma(m[u + 1], (j + 32 < n) ? xs[j + 32] : 0.0, a1);
        a2 = fma(m[u + 2], (j + 64 < n) ? xs[j + 64] : 0.0, a2);
        a3 = fma(m[u + 3], (j + 96 < n) ? xs[j + 96] : 0.0, a3);
      }
    }
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
constexpr int kVecThreads = 1024;  // largest block size of the interface-vector kernels (the host picks 256 or 1024)

__device__ __forceinline__ double block_reduce_sum(double v, double* red /*>= warps per block*/) {
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
  __shared__ double red[kVecThreads / 32];
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

// ---------------------------------------------------------------------------------------------
// Interface assembly FUSED with the exchange over the GPUs of the node (subdomains are sharded over
// the ranks, EllipticPdePllDomainDecomposition.jl:11 `@distributed (+)`): every rank owns a receive
// buffer that its peers map through CUDA IPC (NVLink / NVSwitch peer memory).
//
// OWNER-ORDERED COMPACT EXCHANGE.  The reference sums the contributions to an interface node in
// ascending subdomain order (EllipticPdeDomainDecomposition.jl:740-742); with contiguous shards that is
// rank-major order.  For every interface node the LOWEST rank touching it sends one entry, the sum of
// its own contributions in ascending order (a prefix of the reference's sum; the final value when no
// other rank touches the node -- the vast majority); every HIGHER rank touching the node sends its
// contributions one by one.  All ranks receive all entries (an all-gather of ~n_gamma doubles in total,
// not `world` full-length partial vectors) and finish every node's sum in the reference's order: the
// assembled vector has the SAME BITS on every rank and for every number of GPUs, single GPU included,
// so the PCG scalars and iteration counts do not depend on the GPU count.
// The first ceil(m_r / blockDim) blocks of rank r compute and push its m_r entries to every peer; every thread
// then sums the entries of its node as they arrive (self-validating entries, see exchange_gather).  No separate
// collective launch, no flags.  Two data slots alternate between consecutive REAL
// exchanges (a rank can run at most one exchange ahead of a peer: it cannot finish exchange k+1 before
// that peer has pushed k+1, i.e. has left the kernel of exchange k).  Launches that the device skips after
// convergence must not advance the slot: the number of real exchanges is carried ON THE DEVICE in a
// two-word ping-pong counter -- launch L (host-side launch index, skipped ones included) reads
// cnt[L & 1] and writes cnt[(L + 1) & 1] (+1 if it really exchanged), so no block of a launch ever reads
// a word that the same launch writes.  Identical launch sequences on all ranks => identical counters.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct XchParams {
  double* data[kMaxPeers];     // gathered entries at every rank: [2 slots][m_pad] doubles
  uint32_t* flags[kMaxPeers];  // flags at every rank: one per send block of every rank [nflags]
  const int32_t* snd_ptr;      // this rank's send entries: CSR over indices into `loc` (snd_n + 1)
  const int32_t* snd_src;
  const int32_t* snd_g;        // interface node whose `base` value the entry starts from, or -1
  const int32_t* fin_ptr;      // per interface node: entries of the gathered array to sum, in order
  const int32_t* fin_src;
  int world, rank;
  uint32_t* cnt;               // ping-pong counter of REAL exchanges (see above); seq = cnt[par] + 1, slot = seq & 1
  int par;                     // launch index & 1
  int m_pad;                   // entries per slot
  int snd_off, snd_n;          // position / number of this rank's entries in the gathered array
  int flag_off, nflags;        // this rank's first flag / flags of all ranks
  int* fail;                   // set to 2 when a peer does not show up (time-out instead of a hang)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Exchange step of one block (all threads must call): returns the assembled value of interface node g
// (contributions of ALL subdomains, reference order); see above.
// a launch that does not exchange (skipped after convergence) hands the counter on unchanged
__device__ __forceinline__ void exchange_skip(const XchParams& X) {
  if (blockIdx.x == 0 && threadIdx.x == 0) X.cnt[X.par ^ 1] = __ldcg(X.cnt + X.par);
}

// SELF-VALIDATING ENTRIES.  Every entry of the gathered array is one naturally aligned 64-bit word: a slot that has not
// been written for the current exchange holds kXchEmpty (a NaN bit pattern no arithmetic produces), the sender
// overwrites it with ONE relaxed system-scope store (single-copy atomic), the one thread of the receiving rank that
// sums the entry polls the word until it is not empty and puts the empty pattern back.  No flags, no release / acquire
// fences, no barrier between pushing and gathering: an exchange costs one NVLink store latency (round 2 measured
// 16.4 us per fused kernel at 8 GPUs with flag-per-block publication -- a system-scope fence after the pushes, 265
// acquire polls per block -- against 6.5 us for the same kernel without an exchange).  The two data slots still
// alternate between consecutive real exchanges: a peer may push exchange k+1 while this rank still reads k, and
// cannot push k+2 (same slot as k) before this rank has pushed k+1, i.e. has left the kernel that read -- and emptied
// -- slot k.  Every rank sums every entry exactly once (the assembled vector is replicated), so every slot is emptied.
constexpr unsigned long long kXchEmpty = 0xffffffffffffffffull;
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ double exchange_gather(int g, int n, const double* loc, const double* base, const XchParams& X) {
  const int T = blockDim.x;
  const uint32_t seq = __ldcg(X.cnt + X.par) + 1u;
  const int slot = (int)(seq & 1u);
  if (blockIdx.x == 0 && threadIdx.x == 0) X.cnt[X.par ^ 1] = seq;
  if ((int)blockIdx.x * T < X.snd_n) {  // block-uniform: this block pushes entries
    const int e = blockIdx.x * T + threadIdx.x;
    if (e < X.snd_n) {
      const int gb = X.snd_g[e];
      double a = (base != nullptr && gb >= 0) ? base[gb] : 0.0;
      for (int k = X.snd_ptr[e]; k < X.snd_ptr[e + 1]; ++k) a += loc[X.snd_src[k]];
      unsigned long long bits = (unsigned long long)__double_as_longlong(a);
      if (bits == kXchEmpty) bits = 0x7ff8000000000000ull;   // (a NaN all the same: never mistaken for an empty slot)
      const size_t at = (size_t)slot * (size_t)X.m_pad + (size_t)(X.snd_off + e);
      for (int r = 0; r < X.world; ++r) st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(X.data[r] + at), bits);
    }
  }
  double sum = 0.0;
  if (g < n) {
    unsigned long long* G = reinterpret_cast<unsigned long long*>(X.data[X.rank] + (size_t)slot * (size_t)X.m_pad);
    const long long t0 = clock64();
    for (int k = X.fin_ptr[g]; k < X.fin_ptr[g + 1]; ++k) {
      unsigned long long* p = G + X.fin_src[k];
      unsigned long long v = ld_relaxed_sys_u64(p);
      while (v == kXchEmpty) {
        if (clock64() - t0 > 8000000000LL) {  // ~4 s: a peer never arrived
          if (X.fail != nullptr) *X.fail = 2;
          break;
        }
        v = ld_relaxed_sys_u64(p);
      }
      sum += __longlong_as_double((long long)v);
      st_relaxed_sys_u64(p, kXchEmpty);
    }
  }
  return sum;
}

// local partial of interface entry g: contributions of the local subdomains in ascending order (+ base)
__device__ __forceinline__ double assemble_local(int g, int n, const int32_t* adj_ptr, const int32_t* adj_src,
                                                 const double* loc, const double* base) {
  double a = 0.0;
  if (g < n) {
    a = (base != nullptr) ? base[g] : 0.0;
    for (int k = adj_ptr[g]; k < adj_ptr[g + 1]; ++k) a += loc[adj_src[k]];
  }
  return a;
}

__global__ void __launch_bounds__(kVecThreads)
assemble_exchange_kernel(int n, const int32_t* adj_ptr, const int32_t* adj_src, const double* loc, const double* base,
                         double* y, const double* q, double* part, const int* done_at, int iter, const XchParams X) {
  __shared__ double red[kVecThreads / 32];
  if (done_at != nullptr && *done_at < iter) {
    exchange_skip(X);
    return;
  }
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const double sum = exchange_gather(g, n, loc, base, X);
  double prod = 0.0;
  if (g < n) {
    y[g] = sum;
    if (q != nullptr) prod = sum * q[g];
  }
  if (part != nullptr) {
    const double s = block_reduce_sum(prod, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
}

// y = A x for the assembled interface block A_ΓΓ (CSR, n x n): the base of the interface assembly in the
// global-column operator (apply_global_schur, EllipticPdeDomainDecomposition.jl:604)
__global__ void __launch_bounds__(kVecThreads)
spmv_csr_kernel(int n, const int32_t* ptr, const int32_t* col, const double* val, const double* x, double* y,
                const int* done_at, int iter) {
  if (done_at != nullptr && *done_at < iter) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  double a = 0.0;
  for (int k = ptr[g]; k < ptr[g + 1]; ++k) a = fma(val[k], x[col[k]], a);
  y[g] = a;
}

// part[b] = block partial of dot(a, b)
__global__ void __launch_bounds__(kVecThreads)
dot_kernel(int n, const double* a, const double* b, double* part, const int* done_at, int iter) {
  __shared__ double red[kVecThreads / 32];
  if (done_at != nullptr && *done_at < iter) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const double prod = (g < n) ? a[g] * b[g] : 0.0;
  const double s = block_reduce_sum(prod, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// y = src ; part[b] = block partial of dot(src, q) (optional)   -- tail of the NCCL fallback path
__global__ void __launch_bounds__(kVecThreads)
copy_dot_kernel(int n, const double* src, double* y, const double* q, double* part, const int* done_at, int iter) {
  __shared__ double red[kVecThreads / 32];
  if (done_at != nullptr && *done_at < iter) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double prod = 0.0;
  if (g < n) {
    const double v = src[g];
    y[g] = v;
    if (q != nullptr) prod = v * q[g];
  }
  if (part != nullptr) {
    const double s = block_reduce_sum(prod, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
}

// r = b - Ax ; partials of r.r (partA) and b.b (partB)                            (cg.jl:83-84, 89)
__global__ void __launch_bounds__(kVecThreads)
pcg_init_residual_kernel(int n, const double* b, const double* Ax, double* r, double* partA, double* partB) {
  __shared__ double red[kVecThreads / 32];
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
  __shared__ double red[kVecThreads / 32];
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
  __shared__ double red[kVecThreads / 32];
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
                    PcgState* st, double* res_norm, double* vcol, double* hist, int iter) {
  __shared__ double red[kVecThreads / 32];
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
    if (hist != nullptr) {  // alpha, beta of this iteration (the recycling solvers build V'AV from them)
      hist[2 * it] = st->alpha;
      hist[2 * it + 1] = beta;
    }
    if (!((it < st->maxit) && (sqrt(rTr) > st->tol))) st->done_at = iter;
  }
}


// ---------------------------------------------------------------------------------------------
// Fused PCG steps (plain preconditioned CG, cg.jl:93-105): the interface assembly (with the peer
// exchange when kX), the dot product that needs it, and the vector update that needs the dot run in
// ONE kernel -- a grid-wide arrival counter separates "every block has written its partial" from
// "every block sums the partials" (all blocks of these small grids are co-resident).  Four launches
// per iteration instead of six: stream kernel, fused_xr, dense GEMV, fused_p.
// ---------------------------------------------------------------------------------------------
// Arrive on the grid counter and wait until `target` arrivals have been counted (counter is monotone:
// target = launches so far x blocks).  Kernels skipped after convergence still arrive (grid_skip).
__device__ __forceinline__ void grid_arrive_wait(unsigned int* counter, unsigned int target, int* fail) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      // the host only launches these kernels when the whole grid is co-resident (finalize checks the
      // occupancy); should that ever not hold, this turns a dead-lock into a reported error after ~4 s
      if ((int)(v - target) < 0 && clock64() - t0 > 8000000000LL) {
        if (fail != nullptr) *fail = 3;
        break;
      }
    } while ((int)(v - target) < 0);
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_skip(unsigned int* counter) {
  if (threadIdx.x == 0) atomicAdd(counter, 1u);
}
// fixed-order sum of per-block partials written by other blocks of this launch (read through L2)
__device__ __forceinline__ double sum_partials_cg(const double* part, int nb, double* red) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) v += __ldcg(part + i);
  return block_reduce_sum(v, red);
}

// Ap = sum_d R_d' loc_d ; d = p.Ap ; alpha = rTz / d ; x += alpha p ; r -= alpha Ap ; partB <- r.r
template <bool kX>
__global__ void __launch_bounds__(kVecThreads)
pcg_fused_xr_kernel(int n, const int32_t* adj_ptr, const int32_t* adj_src, const double* loc, double* Ap, const double* p,
                    double* x, double* r, double* partA, double* partB, PcgState* st, int iter, unsigned int* counter,
                    unsigned int target, const XchParams X) {
  __shared__ double red[kVecThreads / 32];
  pdl_trigger();
  pdl_wait();
  if (st->done_at < iter) {
    grid_skip(counter);
    if (kX) exchange_skip(X);
    return;
  }
  const int nb = gridDim.x;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const double a = kX ? exchange_gather(g, n, loc, nullptr, X) : assemble_local(g, n, adj_ptr, adj_src, loc, nullptr);
  double pv = 0.0, xv = 0.0, rold = 0.0;
  if (g < n) {
    Ap[g] = a;
    pv = p[g];
    xv = x[g];     // fetched before the grid-wide wait: their latency hides behind it
    rold = r[g];
  }
  const double rTz = st->rTz;
  const double s = block_reduce_sum(a * pv, red);
  if (threadIdx.x == 0) partA[blockIdx.x] = s;
  grid_arrive_wait(counter, target, &st->fail);
  const double d = sum_partials_cg(partA, nb, red);
  const double alpha = rTz / d;
  double rr = 0.0;
  if (g < n) {
    x[g] = fma(alpha, pv, xv);
    const double rv = fma(-alpha, a, rold);
    r[g] = rv;
    rr = rv * rv;
  }
  const double s2 = block_reduce_sum(rr, red);
  if (threadIdx.x == 0) {
    partB[blockIdx.x] = s2;
    if (blockIdx.x == 0) {
      st->d = d;
      st->alpha = alpha;
      st->rTz_old = rTz;
      if (!(d > 0.0)) st->fail = 1;  // operator not SPD along p: reported, not asserted
    }
  }
}

// z = sum_d R_d' loc_d ; rTz' = r.z ; beta = rTz' / rTz ; p = beta p + z ; it += 1 ; res_norm[it] = sqrt(r.r)
template <bool kX>
__global__ void __launch_bounds__(kVecThreads)
pcg_fused_p_kernel(int n, const int32_t* adj_ptr, const int32_t* adj_src, const double* loc, double* z, const double* r,
                   double* p, const double* partB, double* partC, PcgState* st, double* res_norm, double* vcol, double* hist,
                   int iter, unsigned int* counter, unsigned int target, const XchParams X) {
  __shared__ double red[kVecThreads / 32];
  pdl_trigger();
  pdl_wait();
  if (st->done_at < iter) {
    grid_skip(counter);
    if (kX) exchange_skip(X);
    return;
  }
  const int nb = gridDim.x;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const double a = kX ? exchange_gather(g, n, loc, nullptr, X) : assemble_local(g, n, adj_ptr, adj_src, loc, nullptr);
  double rv = 0.0, pold = 0.0;
  if (g < n) {
    z[g] = a;
    rv = r[g];
    pold = p[g];   // fetched before the grid-wide wait
  }
  const double rTz_old = st->rTz_old;
  const double s = block_reduce_sum(a * rv, red);
  if (threadIdx.x == 0) partC[blockIdx.x] = s;
  // r.r of this iteration: its partials were written by the previous launch, so they can be summed before the wait
  const double rTr = sum_partials_cg(partB, nb, red);
  grid_arrive_wait(counter, target, &st->fail);
  const double rTz_new = sum_partials_cg(partC, nb, red);
  double beta = 1.0 / rTz_old;
  beta *= rTz_new;
  if (g < n) {
    p[g] = fma(beta, pold, a);
    if (vcol != nullptr) vcol[g] = a / sqrt(rTz_new);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int it = st->it + 1;
    st->it = it;
    st->rTr = rTr;
    st->rTz = rTz_new;
    st->beta = beta;
    res_norm[it - 1] = sqrt(rTr);
    if (hist != nullptr) {
      hist[2 * it] = st->alpha;
      hist[2 * it + 1] = beta;
    }
    if (!((it < st->maxit) && (sqrt(rTr) > st->tol))) st->done_at = iter;
  }
}

// State of a recurrence continued from host-side scalars (kb200_pcg_burst)
__global__ void pcg_burst_init_kernel(PcgState* st, double rTz, double tol, int it, int maxit) {
  st->rTz = rTz;
  st->rTz_old = rTz;
  st->tol = tol;
  st->it = it;
  st->maxit = maxit;
  st->fail = 0;
  st->done_at = 0x7fffffff;
}

// ---------------------------------------------------------------------------------------------
// Dense tall-skinny products with the deflation / recycling bases W, V (n_gamma x k, column-major):
// reference defcg.jl:301-302 (WtA*z, W*mu), :407, eigcg.jl:219-228 (V'AV, V*(Q*Z)).  HBM-bound, FP64.
// ---------------------------------------------------------------------------------------------
// y = a*x + b*y
__global__ void __launch_bounds__(kVecThreads) axpby_kernel(int n, double a, const double* x, double b, double* y) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) y[g] = (b == 0.0 ? 0.0 : b * y[g]) + a * x[g];
}
// out[j] = A[:, j] . x   (one block per column, fixed reduction order)
__global__ void __launch_bounds__(kVecThreads) gemv_t_kernel(int n, const double* A, int64_t lda, const double* x, double* out,
                                                           const int* done_at = nullptr, int iter = 0) {
  __shared__ double red[kVecThreads / 32];
  if (done_at != nullptr && *done_at < iter) return;
  const double* col = A + (int64_t)blockIdx.x * lda;
  double s = 0.0;
  for (int g = threadIdx.x; g < n; g += blockDim.x) s = fma(col[g], x[g], s);
  const double t = block_reduce_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = t;
}
// y = b*y + a * A[:, 0:k] * mu
__global__ void __launch_bounds__(kVecThreads) gemv_n_kernel(int n, int k, double a, const double* A, int64_t lda, const double* mu,
                                                           double b, double* y, const int* done_at = nullptr, int iter = 0) {
  extern __shared__ double smu[];
  if (done_at != nullptr && *done_at < iter) return;
  for (int j = threadIdx.x; j < k; j += blockDim.x) smu[j] = mu[j];
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  double s0 = 0.0, s1 = 0.0;
  int j = 0;
  for (; j + 1 < k; j += 2) {
    s0 = fma(A[g + (int64_t)j * lda], smu[j], s0);
    s1 = fma(A[g + (int64_t)(j + 1) * lda], smu[j + 1], s1);
  }
  if (j < k) s0 = fma(A[g + (int64_t)j * lda], smu[j], s0);
  y[g] = (b == 0.0 ? 0.0 : b * y[g]) + a * (s0 + s1);
}
// mu = M^{-1} y for the small deflation system (defcg.jl:301, `WtAW \\ (WtA*z)`): explicit inverse
// applied with one step of iterative refinement, mu = Minv y; mu += Minv (y - M mu), which recovers
// the accuracy of a backward-stable solve without any host round trip.  One CTA; matrices k x k,
// column-major, read coalesced.
__global__ void __launch_bounds__(1024) small_solve_kernel(int k, const double* Minv, const double* M, const double* y, double* mu,
                                                           const int* done_at, int iter) {
  extern __shared__ double sh[];  // y | mu | r
  if (done_at != nullptr && *done_at < iter) return;
  double* sy = sh;
  double* smu = sh + k;
  double* sr = sh + 2 * k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) sy[i] = y[i];
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < k; ++j) s = fma(Minv[i + (int64_t)j * k], sy[j], s);
    smu[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    double s = sy[i];
    for (int j = 0; j < k; ++j) s = fma(-M[i + (int64_t)j * k], smu[j], s);
    sr[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    double s = smu[i];
    for (int j = 0; j < k; ++j) s = fma(Minv[i + (int64_t)j * k], sr[j], s);
    mu[i] = s;
  }
}

// C[i, j] = A[:, i] . B[:, j]  (ka x kb, column-major) -- 16x16 output tile per block, rows streamed in
// 64-row panels through shared memory; fixed summation order.
__global__ void __launch_bounds__(256) gemm_tn_kernel(int n, int ka, int kb, const double* A, int64_t lda, const double* B,
                                                      int64_t ldb, double* Cm) {
  __shared__ double sa[64][17], sb[64][17];
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  const int i0 = blockIdx.x * 16, j0 = blockIdx.y * 16;
  double acc = 0.0;
  for (int r0 = 0; r0 < n; r0 += 64) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e & 63, c = e >> 6;
      const int g = r0 + r;
      sa[r][c] = (g < n && i0 + c < ka) ? A[g + (int64_t)(i0 + c) * lda] : 0.0;
      sb[r][c] = (g < n && j0 + c < kb) ? B[g + (int64_t)(j0 + c) * ldb] : 0.0;
    }
    __syncthreads();
#pragma unroll 16
    for (int r = 0; r < 64; ++r) acc = fma(sa[r][ti], sb[r][tj], acc);
    __syncthreads();
  }
  if (i0 + ti < ka && j0 + tj < kb) Cm[(i0 + ti) + (int64_t)(j0 + tj) * ka] = acc;
}
// C[:, j] = A[:, 0:ka] * Q[:, j]   (n x kq); Q (ka x kq, column-major) in global memory
__global__ void __launch_bounds__(kVecThreads) gemm_nn_kernel(int n, int ka, int kq, const double* A, int64_t lda, const double* Q,
                                                            double* Cm, int64_t ldc) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (g >= n) return;
  const double* q = Q + (int64_t)j * ka;
  double s0 = 0.0, s1 = 0.0;
  int kk = 0;
  for (; kk + 1 < ka; kk += 2) {
    s0 = fma(A[g + (int64_t)kk * lda], __ldg(q + kk), s0);
    s1 = fma(A[g + (int64_t)(kk + 1) * lda], __ldg(q + kk + 1), s1);
  }
  if (kk < ka) s0 = fma(A[g + (int64_t)kk * lda], __ldg(q + kk), s0);
  Cm[g + (int64_t)j * ldc] = s0 + s1;
}

}  // namespace kb

// =============================================================================================
// FP64 tensor-core (DMMA) kernels: the dense contractions of the path -- batched dense local products
// with several right-hand sides (assembled S_d X_d, EllipticPdeDomainDecomposition.jl:761-785, and
// Pi S_d (D R)_d, :1361-1386, for the multiple-right-hand-side solves of
// Example11_EigInitPcgEllipticPdeMultipleRhs_Functions.jl:140-179) and the tall-skinny products with
// the recycling bases (V'AV eigcg.jl:219, V (Q Z) eigcg.jl:228 / defcg.jl:432, W'A W defcg.jl:268).
// mma.sync.m8n8k4.f64: A 8x4 (lane l: row l/4, col l%4), B 4x8 (lane l: row l%4, col l/4),
// C 8x8 (lane l: row l/4, cols 2(l%4), 2(l%4)+1).  Fixed summation order: bit-reproducible.
// =============================================================================================
namespace kb {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- batched dense local products with k right-hand sides -------------------------------------
// out_loc[G_base_d + i + j ldo] = w_i sum_m M_d[i, m] (w_m x[l2g_d[m] + j ldx]),  j < k
// grid = (row tiles of 64, subdomains); 8 warps, warp w owns rows 8w..8w+7 of the tile and all KP columns
struct GemmParams {
  const double* mats;       // row-major n_gd x n_gd blocks (symmetric)
  const int64_t* mat_off;
  const int32_t* n_gd;
  const int64_t* G_base;
  const int32_t* l2g;
  const double* x;          // n_gamma x k, column-major
  int64_t ldx;
  const double* inv_cnt;    // 1/node_gamma_cnt or nullptr
  double* out_loc;          // (sum n_gd) x k, column-major
  int64_t ldo;
  int k;                    // right-hand sides (<= KP)
  const int* done_at;
  int iter;
};
constexpr int kGemmBM = 64, kGemmBK = 32;

template <int KP>
__global__ void __launch_bounds__(256) dense_gemm_kernel(const GemmParams P) {
  __shared__ double As[kGemmBM][kGemmBK + 4];
  __shared__ double Xs[kGemmBK][KP + 4];
  if (P.done_at != nullptr && *P.done_at < P.iter) return;
  const int d = blockIdx.y;
  const int n = P.n_gd[d];
  const int row0 = blockIdx.x * kGemmBM;
  if (row0 >= n) return;
  const int64_t gb = P.G_base[d];
  const int32_t* l2g = P.l2g + gb;
  const double* M = P.mats + P.mat_off[d];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NF = KP / 8;
  double c[NF][2];
#pragma unroll
  for (int f = 0; f < NF; ++f) c[f][0] = c[f][1] = 0.0;
  for (int kc = 0; kc < n; kc += kGemmBK) {
    for (int e = tid; e < kGemmBM * kGemmBK; e += 256) {
      const int r = e / kGemmBK, cc = e % kGemmBK;
      const int i = row0 + r, m = kc + cc;
      As[r][cc] = (i < n && m < n) ? __ldg(M + (int64_t)i * n + m) : 0.0;
    }
    for (int e = tid; e < kGemmBK * KP; e += 256) {
      const int kk = e % kGemmBK, j = e / kGemmBK;
      const int m = kc + kk;
      double v = 0.0;
      if (m < n && j < P.k) {
        const int g = l2g[m];
        v = P.x[g + (int64_t)j * P.ldx];
        if (P.inv_cnt != nullptr) v *= P.inv_cnt[g];
      }
      Xs[kk][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < kGemmBK / 4; ++k4) {
      const double a = As[warp * 8 + (lane >> 2)][k4 * 4 + (lane & 3)];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const double b = Xs[k4 * 4 + (lane & 3)][f * 8 + (lane >> 2)];
        dmma884(c[f][0], c[f][1], a, b);
      }
    }
    __syncthreads();
  }
  const int i = row0 + warp * 8 + (lane >> 2);
  if (i < n) {
    const double w = (P.inv_cnt != nullptr) ? P.inv_cnt[l2g[i]] : 1.0;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int j = f * 8 + 2 * (lane & 3);
      if (j < P.k) P.out_loc[gb + i + (int64_t)j * P.ldo] = c[f][0] * w;
      if (j + 1 < P.k) P.out_loc[gb + i + (int64_t)(j + 1) * P.ldo] = c[f][1] * w;
    }
  }
}

// ---- C = A' B, tall-skinny (n rows, ka x kb result), split over row ranges ------------------------
// CTA tile 64 x 128 of C (8 warps as 2 x 4, 32 x 32 each), rows [r0, r1) of this split; operands straight
// from global memory (a fragment is 4 consecutive rows of 8 columns: full 32-byte sectors).  Partial tiles go
// to `part` [nsplit][ka * kb]; gemm_tn_reduce_kernel sums the splits in order.
constexpr int kTnTM = 64, kTnTN = 128;
__global__ void __launch_bounds__(256) gemm_tn_dmma_kernel(int n, int ka, int kb, const double* A, int64_t lda, const double* B,
                                                           int64_t ldb, double* part, int rows_per_split) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * kTnTM + (warp >> 2) * 32, j0 = blockIdx.y * kTnTN + (warp & 3) * 32;
  const int r0 = blockIdx.z * rows_per_split, r1 = min(n, r0 + rows_per_split);
  double c[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) c[mi][ni][0] = c[mi][ni][1] = 0.0;
  // column pointers of this lane's fragments (columns clamped into range; out-of-range results are dropped)
  const double* ap[4];
  const double* bp[4];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) ap[mi] = A + (int64_t)min(i0 + mi * 8 + (lane >> 2), ka - 1) * lda + (lane & 3);
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) bp[ni] = B + (int64_t)min(j0 + ni * 8 + (lane >> 2), kb - 1) * ldb + (lane & 3);
  if (i0 < ka && j0 < kb) {
    int r = r0;
    for (; r + 4 <= r1; r += 4) {
      double a[4], b[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) a[mi] = __ldg(ap[mi] + r);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = __ldg(bp[ni] + r);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(c[mi][ni][0], c[mi][ni][1], a[mi], b[ni]);
    }
    if (r < r1) {  // tail rows: lanes beyond the range feed zeros
      const bool ok = r + (lane & 3) < r1;
      double a[4], b[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) a[mi] = ok ? __ldg(ap[mi] + r) : 0.0;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = ok ? __ldg(bp[ni] + r) : 0.0;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(c[mi][ni][0], c[mi][ni][1], a[mi], b[ni]);
    }
  }
  double* out = part + (int64_t)blockIdx.z * ka * kb;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int i = i0 + mi * 8 + (lane >> 2), j = j0 + ni * 8 + 2 * (lane & 3);
      if (i < ka && j < kb) out[i + (int64_t)j * ka] = c[mi][ni][0];
      if (i < ka && j + 1 < kb) out[i + (int64_t)(j + 1) * ka] = c[mi][ni][1];
    }
}
__global__ void __launch_bounds__(256) gemm_tn_reduce_kernel(int64_t nel, int nsplit, const double* part, double* C) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nel) return;
  double s = 0.0;
  for (int z = 0; z < nsplit; ++z) s += part[(int64_t)z * nel + e];
  C[e] = s;
}

// ---- C[:, 0:kq] = A[:, 0:ka] Q, C may alias A (V[:, 1:nev] = V (Q Z), eigcg.jl:228) -----------------
// One CTA per panel of 32 rows: the panel of A is staged ONCE in shared memory (A is read exactly once from
// HBM, and may be overwritten by C afterwards), then 64 output columns at a time: warp w owns 8 columns and
// the 4 row fragments.  Q (ka x kq, column-major) comes from global memory (L2-resident).  ka <= kNnMaxK.
constexpr int kNnRows = 32, kNnLd = 36, kNnMaxK = 784;
__global__ void __launch_bounds__(256) gemm_nn_dmma_kernel(int n, int ka, int kq, const double* A, int64_t lda, const double* Q,
                                                           double* Cm, int64_t ldc) {
  extern __shared__ __align__(16) double Ap[];  // [ka][kNnLd]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * kNnRows;
  const int kap = (ka + 3) & ~3;
  for (int e = threadIdx.x; e < kap * kNnRows; e += 256) {
    const int kk = e / kNnRows, i = e % kNnRows;
    Ap[kk * kNnLd + i] = (kk < ka && r0 + i < n) ? A[r0 + i + (int64_t)kk * lda] : 0.0;
  }
  __syncthreads();
  for (int jb = 0; jb < kq; jb += 64) {
    const int j = jb + warp * 8 + (lane >> 2);          // column of this lane's B fragment
    const double* qp = Q + (int64_t)min(j, kq - 1) * ka;
    double c[4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) c[mi][0] = c[mi][1] = 0.0;
    if (jb + warp * 8 < kq) {
      for (int k0 = 0; k0 < kap; k0 += 4) {
        const int kk = k0 + (lane & 3);
        const double b = (kk < ka && j < kq) ? __ldg(qp + kk) : 0.0;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) dmma884(c[mi][0], c[mi][1], Ap[kk * kNnLd + mi * 8 + (lane >> 2)], b);
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int i = r0 + mi * 8 + (lane >> 2), jc = jb + warp * 8 + 2 * (lane & 3);
        if (i < n && jc < kq) Cm[i + (int64_t)jc * ldc] = c[mi][0];
        if (i < n && jc + 1 < kq) Cm[i + (int64_t)(jc + 1) * ldc] = c[mi][1];
      }
    }
  }
}

}  // namespace kb

// =============================================================================================
// Independent PCG recurrences for a block of right-hand sides (kb200_pcg_multi): column j of every vector at
// + j * ld, its scalars in st[j], its block partials at part + j * nb.  grid = (nb, columns).  The operator
// launches are shared by the columns (multi-vector stream kernel / DMMA dense products), the recurrences are
// NOT coupled (no block-CG): every column iterates exactly like the reference's sequential loop over
// right-hand sides (Example11_EigInitPcgEllipticPdeMultipleRhs_Functions.jl:140-179, cg.jl:67-109) and stops
// on its own; a converged column is frozen.  Same per-column arithmetic as the single-vector kernels above.
// =============================================================================================
namespace kb {

__global__ void __launch_bounds__(kVecThreads)
assemble_multi_kernel(int n, const int32_t* adj_ptr, const int32_t* adj_src, const double* loc, int64_t ldl, double* y,
                      const double* q, int64_t ld, double* part, const PcgState* st, int iter) {
  __shared__ double red[kVecThreads / 32];
  const int j = blockIdx.y;
  if (st != nullptr && st[j].done_at < iter) return;
  loc += (int64_t)j * ldl;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double prod = 0.0;
  if (g < n) {
    double a = 0.0;
    for (int k = adj_ptr[g]; k < adj_ptr[g + 1]; ++k) a += loc[adj_src[k]];
    y[g + (int64_t)j * ld] = a;
    if (q != nullptr) prod = a * q[g + (int64_t)j * ld];
  }
  if (part != nullptr) {
    const double s = block_reduce_sum(prod, red);
    if (threadIdx.x == 0) part[(int64_t)j * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kVecThreads)
pcg_init_residual_multi_kernel(int n, const double* b, const double* Ax, double* r, int64_t ld, double* partA, double* partB) {
  __shared__ double red[kVecThreads / 32];
  const int64_t o = (int64_t)blockIdx.y * ld;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double rr = 0.0, bb = 0.0;
  if (g < n) {
    const double rv = b[o + g] - Ax[o + g];
    r[o + g] = rv;
    rr = rv * rv;
    bb = b[o + g] * b[o + g];
  }
  const double s1 = block_reduce_sum(rr, red);
  const double s2 = block_reduce_sum(bb, red);
  if (threadIdx.x == 0) {
    partA[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s1;
    partB[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s2;
  }
}

__global__ void __launch_bounds__(kVecThreads)
pcg_init_direction_multi_kernel(int n, const double* z, double* p, int64_t ld, const double* partA, const double* partB,
                                const double* partC, PcgState* st, double* res_norm, int64_t res_stride, double rtol, int maxit) {
  __shared__ double red[kVecThreads / 32];
  const int j = blockIdx.y, nb = gridDim.x;
  const int64_t o = (int64_t)j * ld;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[o + g] = z[o + g];
  if (blockIdx.x == 0) {
    const double rTr = sum_partials(partA + (int64_t)j * nb, nb, red);
    const double bb = sum_partials(partB + (int64_t)j * nb, nb, red);
    const double rTz = sum_partials(partC + (int64_t)j * nb, nb, red);
    if (threadIdx.x == 0) {
      PcgState& s = st[j];
      s.rTr = rTr;
      s.rTz = rTz;
      s.bnorm = sqrt(bb);
      s.tol = rtol * sqrt(bb);
      s.it = 1;
      s.maxit = maxit;
      s.fail = 0;
      res_norm[(int64_t)j * res_stride] = sqrt(rTr);
      s.done_at = ((1 < maxit) && (sqrt(rTr) > s.tol)) ? 0x7fffffff : 0;
    }
  }
}

__global__ void __launch_bounds__(kVecThreads)
pcg_update_xr_multi_kernel(int n, const double* p, const double* Ap, double* x, double* r, int64_t ld, const double* partA,
                           double* partB, PcgState* st, int iter) {
  __shared__ double red[kVecThreads / 32];
  const int j = blockIdx.y, nb = gridDim.x;
  if (st[j].done_at < iter) return;
  const int64_t o = (int64_t)j * ld;
  const double d = sum_partials(partA + (int64_t)j * nb, nb, red);
  const double rTz = st[j].rTz;
  const double alpha = rTz / d;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double rr = 0.0;
  if (g < n) {
    x[o + g] = fma(alpha, p[o + g], x[o + g]);
    const double rv = fma(-alpha, Ap[o + g], r[o + g]);
    r[o + g] = rv;
    rr = rv * rv;
  }
  const double s = block_reduce_sum(rr, red);
  if (threadIdx.x == 0) {
    partB[(int64_t)j * nb + blockIdx.x] = s;
    if (blockIdx.x == 0) {
      st[j].d = d;
      st[j].alpha = alpha;
      st[j].rTz_old = rTz;
      if (!(d > 0.0)) st[j].fail = 1;
    }
  }
}

// (rTz_old is written by block 0 of the xr kernel of the same iteration, a previous launch)
__global__ void __launch_bounds__(kVecThreads)
pcg_update_p_multi_kernel(int n, const double* z, double* p, int64_t ld, const double* partB, const double* partC, PcgState* st,
                          double* res_norm, int64_t res_stride, int iter) {
  __shared__ double red[kVecThreads / 32];
  const int j = blockIdx.y, nb = gridDim.x;
  if (st[j].done_at < iter) return;
  const int64_t o = (int64_t)j * ld;
  const double rTr = sum_partials(partB + (int64_t)j * nb, nb, red);
  const double rTz_new = sum_partials(partC + (int64_t)j * nb, nb, red);
  double beta = 1.0 / st[j].rTz_old;
  beta *= rTz_new;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[o + g] = fma(beta, p[o + g], z[o + g]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int it = st[j].it + 1;
    st[j].it = it;
    st[j].rTr = rTr;
    st[j].rTz = rTz_new;
    st[j].beta = beta;
    res_norm[(int64_t)j * res_stride + it - 1] = sqrt(rTr);
    if (!((it < st[j].maxit) && (sqrt(rTr) > st[j].tol))) st[j].done_at = iter;
  }
}

// min over the columns' done_at (INT_MAX while any column still runs) -> *all_done_at: lets the shared operator
// launches of later iterations exit early once every column has converged
__global__ void pcg_multi_all_done_kernel(int k, const PcgState* st, int* all_done_at) {
  int v = 0;
  for (int j = 0; j < k; ++j) v = max(v, st[j].done_at);
  *all_done_at = v;
}

// upper triangle -> lower triangle of every dense S_d (Symmetric(), EllipticPdeDomainDecomposition.jl:692);
// grid = (tiles, subdomains)
__global__ void __launch_bounds__(256) symmetrize_kernel(double* sd_pool, const int64_t* mat_off, const int32_t* n_gd) {
  const int d = blockIdx.y;
  const int64_t n = n_gd[d];
  double* M = sd_pool + mat_off[d];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % n, j = e / n;       // column-major (i, j)
    if (i > j) M[i + j * n] = M[j + i * n];
  }
}

}  // namespace kb

// =============================================================================================
// VALUE REFRESH for a new coefficient realization (plan.h, "VALUE REFRESH"): the stream keeps its structure,
// only its values are rewritten, on the device.  (1) the in-block entries of L are scattered into the dense
// diagonal blocks D of the inverted supernode groups, (2) every block is re-inverted with the arithmetic of the
// host analysis (invert_lower_block in plan.cpp: same order of operations, no FMA contraction -- the refreshed
// stream has the bits of a stream analysed afresh from the new values), (3) every value slot of the stream is
// rewritten from the value pool.
// =============================================================================================
namespace kb {

__global__ void __launch_bounds__(256) refresh_dmap_kernel(int64_t nmap, const DmapRef* dmap, const double* pool, double* dwork) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nmap) dwork[dmap[e].dslot] = pool[dmap[e].lidx];
}

struct BlockRef {
  uint32_t dwork_off;  // D of the block in the workspace (row-major sz x sz)
  uint32_t pool_off;   // Dinv of the block in the value pool
  uint32_t sz, pad_;
};

// one CTA per block, one thread per column of the inverse (a column only depends on itself)
__global__ void __launch_bounds__(128) refresh_invert_kernel(const BlockRef* blocks, const double* dwork, double* pool) {
  const BlockRef B = blocks[blockIdx.x];
  const int sz = (int)B.sz;
  const double* D = dwork + B.dwork_off;
  double* Dinv = pool + B.pool_off;
  for (int c0 = threadIdx.x; c0 < sz; c0 += blockDim.x) {
    Dinv[(size_t)c0 * sz + c0] = 1.0 / D[(size_t)c0 * sz + c0];
    for (int i = c0 + 1; i < sz; ++i) {
      double acc = 0.0;
      bool any = false;
      const double* Di = D + (size_t)i * sz;
      for (int k = c0; k < i; ++k) {
        const double d = Di[k], v = Dinv[(size_t)k * sz + c0];
        if (d != 0.0 && v != 0.0) {
          acc = __dadd_rn(acc, __dmul_rn(d, v));
          any = true;
        }
      }
      Dinv[(size_t)i * sz + c0] = any ? -acc / Di[i] : 0.0;
    }
  }
}

__global__ void __launch_bounds__(256) refresh_stream_kernel(int64_t nref, const ValRef* refs, const double* pool, double* stream8) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nref) return;
  const ValRef r = refs[e];
  double v = pool[r.src & ~kRefRecip];
  if (r.src & kRefRecip) v = 1.0 / v;
  stream8[r.off8] = v;
}

}  // namespace kb

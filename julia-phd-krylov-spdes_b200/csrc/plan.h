// Host-side analysis ("plan") for the batched interior-solve / local-Schur kernel.
//
// One subdomain d is handled by a thread-block cluster of `csize` CTAs.  The elimination tree of
// its Cholesky factor L_d (A_II,d[p,p] = L L^T, uploaded from the host's CHOLMOD-style factor,
// reference: Fem/EllipticPdeDomainDecomposition.jl:1536, :648-650) is cut into disjoint subtrees
// that are bin-packed onto the CTAs ("own" nodes, solved entirely out of that CTA's shared
// memory) plus a small "top" set of ancestor nodes solved by CTA rank 0 after the partial
// products of the subtrees have been combined over distributed shared memory.
//
// Every numerical step of S_d x_d = A_GG,d x_d - A_GI,d A_II,d^{-1} A_IG,d x_d (reference
// apply_local_schur, EllipticPdeDomainDecomposition.jl:639-654) is expressed as the same row
// operation on a shared-memory vector S:
//        S[dst] = (S[dst] - sign * sum_k val[k] * S[col[k]]) / diag
// grouped into sections (SpMV, forward levels, partial products, top solve, backward levels,
// output products); rows of one dependency level are independent and a CTA barrier separates
// levels.  Each warp owns one contiguous stream (slice headers, row records, values, uint16
// columns) laid out in processing order, so the device reads it strictly sequentially.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kb {

constexpr int kWarps = 8;          // consumer warps per CTA; one more warp feeds the ring (288 threads)
constexpr int kChunkBytes = 16384; // unit of the TMA bulk copies global -> shared ring

enum Section : int {
  SEC_INIT = 0,  // w   = A_IG x_d                 (rows: interior nodes touching the interface)
  SEC_FWD,       // L y = w      on the CTA's own subtrees, level by level
  SEC_PART,      // -sum_{j in own} L[i,j] y_j     for top rows i (combined by rank 0)
  SEC_TOPF,      // forward solve on the top nodes: rows of a level shared out over the cluster's CTAs,
  SEC_TOPB,      // backward solve likewise; results broadcast over DSMEM, cluster barrier per level
  SEC_BWD,       // L^T v = y    on the CTA's own subtrees (top values replicated)
  SEC_GG,        // out = A_GG x_d                 (rank 0)
  SEC_OUT,       // out -= A_IG^T v  restricted to the CTA's own columns
  NSEC
};

// The program of a CTA is ONE byte stream, cut into chunks of kChunkBytes that a producer warp
// copies into a shared-memory ring with cp.async.bulk (TMA) far ahead of the consumers; every
// section starts on a chunk boundary.
//
// Work is organised in LEVELS of independent rows.  A slice is the unit of work of one consumer warp:
// 32 >> tlog2 rows of similar length, T = 1 << tlog2 lanes per row.  Lane (g, t) accumulates entries
// t, t+T, t+2T, ... of row g, the T partial sums are combined with log2(T) butterfly steps and lane
// t = 0 applies
//   S[dst] = (S[dst] - sign * sum) * (1/diag)      (kRowFresh: S[dst] = -sign * sum * (1/diag)),
// optionally storing the result into every other CTA of the cluster too (kRowBcast, top nodes).
// Entries are stored step-major over the nrows*T active lanes: entry (k, lane) at k*nrows*T + lane,
// zero-filled where a row has run out (rows of a slice have similar lengths; a padding entry has
// val 0 and col 0).  Every sum has a fixed association order (bit-reproducible).
//
// WARP PROGRAMS.  A chunk carries, for each of the NW consumer warps (NW = BuildOptions::warps, 8 or
// 16), a list of 16-bit tokens: the offset of a slice inside the chunk (16-byte units), or kTokBarrier
// = "the level ends here" (a CTA barrier; in the top sections a barrier over the whole cluster).  A
// warp only ever reads its own token list, so the per-level control cost is one token per warp
// instead of a header walk by every warp; the host balances the slices of a level over the warps
// (longest-processing-time first on a latency estimate).  Every warp's program holds the same
// sequence of barrier tokens.
// Chunk:  [kChunkDirBytes: (NW+1) x uint16 first token of every warp][tokens, padded to 16 B][slices]
// Slice:  [SliceHdr 16 B][nrows x RowRec 16 B][nent x double][ceil(kmax / 2) x nrows*T x uint32: the columns of
//          two consecutive steps of a lane, low half = even step]  (+ padding to 16 B)
//
// Optionally (BuildOptions::warp_private) the bottom of the elimination forest is processed
// warp-privately: the CTA's own nodes are cut once more into small subtrees that are bin-packed onto
// the consumer warps; a warp walks the rows of its subtrees level by level with only __syncwarp()
// between its slices (tokens without barriers).  Forward sweep: private part first, one barrier,
// then the levels of the remaining ("upper") own nodes; backward sweep: upper levels first.
constexpr uint16_t kTokBarrier = 0xffffu;
constexpr int kChunkDirBytes = 64;  // room for (24 + 1) first-token indices, 16-byte aligned
constexpr int kMaxConsumerWarps = 24;
constexpr int kMaxSlots = 256;     // scratch doubles used when the top right-hand side is combined (passes of kMaxSlots)
struct alignas(16) SliceHdr {
  uint16_t nent;    // stored entries = kmax * nrows * T
  uint16_t kmax;    // steps = ceil(longest row / T)
  uint8_t tlog2;
  uint8_t nrows;
  uint16_t pad0_;
  uint32_t width;     // active lanes = nrows << tlog2 (precomputed for the device)
  uint32_t cols_off;  // byte offset of the packed column indices from the slice start
};
static_assert(sizeof(SliceHdr) == 16, "SliceHdr must be 16 bytes");

constexpr uint16_t kRowFresh = 1, kRowBcast = 2;
struct alignas(16) RowRec {
  uint16_t len;
  uint16_t dst;     // index into S
  uint16_t flags;
  uint16_t pad_;
  double diag;      // RECIPROCAL of the row's divisor
};
static_assert(sizeof(RowRec) == 16, "RowRec must be 16 bytes");

struct CtaDesc {
  int32_t dom, rank, csize;
  int32_t n_own, n_top, n_gd;
  int32_t part_off, top_off, out_off, xd_off, own_off, smem_doubles, top_sz, pad_;
  int64_t l2g_off;     // l2g pool: interface id of local interface node l (n_gd entries)
  int64_t ownmap_off;  // imap pool: original interior index of own position (n_own entries)
  int64_t topmap_off;  // imap pool: original interior index of top position (n_top entries)
  int64_t I_base;      // offset of the subdomain in concatenated interior vectors
  int64_t G_base;      // offset of the subdomain in concatenated local-interface vectors
  uint64_t stream_off; // byte offset of the CTA's stream in the stream pool (chunk aligned)
  uint32_t sec_chunk0[NSEC], sec_chunk1[NSEC];  // chunk range of every section
  uint32_t sec_nlevels[NSEC];
};

// 0-based CSC as converted from the ABI's Int64 1-based arrays
struct Csc {
  int32_t nrow = 0, ncol = 0;
  std::vector<int32_t> colptr, rowval;
  std::vector<double> nzval;
};

struct DomainIn {
  int32_t n_I = 0, n_gd = 0;
  std::vector<int32_t> perm;  // perm[new] = old
  Csc L;                      // lower triangular incl. diagonal, sorted rows
  Csc AIG;                    // n_I x n_gd (original interior row numbering)
  Csc AGG;                    // n_gd x n_gd
};

// VALUE REFRESH (new coefficient realization on the same mesh / partition / ordering, the loop of
// Example07_PcgSchurStochasticEllipticPde.jl:160-204).  The plan is purely structural: which rows, levels,
// slices and tokens exist depends only on the sparsity patterns.  Every value the stream holds comes from the
// subdomain's VALUE POOL
//     [ L.nzval | A_IG.nzval | A_GG.nzval | Dinv of block 0 (row-major sz x sz) | Dinv of block 1 | ... ]
// either directly or as a reciprocal (the row divisors): ValRef says which 8-byte slot of the stream takes which
// pool entry.  For a new realization the caller's nzval arrays are copied into the pool, the supernode blocks
// are re-inverted (DmapRef: which entry of L sits where in the dense diagonal block D of a block) and the
// stream slots are rewritten by a scatter kernel -- no host analysis, no re-upload of the stream.
constexpr uint32_t kRefRecip = 0x80000000u;  // ValRef::src flag: the slot holds 1 / pool[src]
struct ValRef {
  uint32_t off8;  // 8-byte slot in the stream (plan-local; made global by the caller)
  uint32_t src;   // pool entry (plan-local) | kRefRecip
};
struct DmapRef {
  uint32_t dslot;  // entry of the D workspace (Dinv region layout: block offset + row * sz + col)
  uint32_t lidx;   // entry of L.nzval
};

struct DomainPlan {
  std::vector<CtaDesc> ctas;      // csize descriptors (offsets local to this plan's pools)
  std::vector<uint8_t> stream;    // concatenated CTA streams (each a multiple of kChunkBytes)
  // bytes of every chunk that carry data, in 16-byte units (directory + tokens + slices): the producer warp
  // copies only those -- the padding behind them (a slice never straddles a chunk) is not HBM traffic
  std::vector<uint16_t> chunk_units;
  std::vector<int32_t> imap;      // ownmap / topmap entries
  // value refresh (see above)
  std::vector<ValRef> refs;
  std::vector<DmapRef> dmap;
  std::vector<uint32_t> blk_off, blk_sz;   // per inverted block: offset in the Dinv region, size
  uint32_t pool_L = 0, pool_G = 0, pool_GG = 0, pool_D = 0;   // pool segment lengths
  // statistics
  int64_t nnz_L = 0;              // strictly-lower nnz of L
  int64_t n_entries = 0, n_rows = 0, n_packets = 0;   // n_packets = slices
  int64_t lane_steps = 0;                             // sum over slices of 32 * kmax (lane utilisation = n_entries / lane_steps)
  int32_t n_top = 0, max_levels = 0;
  int64_t n_private_rows = 0, n_private_slices = 0;   // rows / slices handled warp-privately
  int64_t max_cta_entries = 0, top_entries = 0;
};

struct BuildOptions {
  int csize = 1;
  int lane_entries = 16;      // aimed entries per lane of a row
  int adaptive_lanes = 0;     // 1: per level, entries per lane = level entries / (warps * 32) clamped to [lane_min, lane_max]
  int lane_min = 4, lane_max = 32;
  double pad_tol = 0.0;       // a slice only takes rows at least pad_tol x its longest row (0: no limit)
  int min_block = 4;          // invert dense diagonal blocks of supernode chains at least this long (0: off)
  int max_block = 1024;
  int merge_depth = 1;        // supernodes of `merge_depth` consecutive depths of the supernode tree share one inverted block
  int max_smem_doubles = 28000;
  int warps = kWarps;         // consumer warps of the kernel that will run the plan (8, 16 or 24): slot k of a level -> warp k % warps
  int avoid_bank_conflicts = 1;  // 1: the entries of a row are dealt to the steps against bank conflicts of the gathered S[col]
  int fill_chunks = 1;        // 1: the tail of a chunk takes any slice of the level that still fits (0: strict order)
  int warp_private = 0;       // 1: subtrees near the leaves are processed warp-privately (no CTA barriers)
  int private_split = 2;      // aim at about private_split x kWarps private subtrees per CTA
};

// Throws std::runtime_error on malformed input (missing diagonal, not lower triangular, ...).
void build_domain_plan(const DomainIn& in, const BuildOptions& opt, DomainPlan& out);

// Inverse of the dense lower-triangular diagonal block D (row-major sz x sz, zeros outside the factor's
// pattern) -- THE arithmetic of the selective inversion, shared by the host analysis and (same order of
// operations) by the device refresh kernel: column by column, terms with a zero factor skipped.
void invert_lower_block(int sz, const double* D, double* Dinv);

// Section signs per mode (see kernels.cu)
enum Mode : int { MODE_APPLY = 0, MODE_RHS = 1, MODE_BACKSUB = 2, MODE_UNIT = 3 };

}  // namespace kb

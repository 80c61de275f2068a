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

constexpr int kWarps = 8;  // warps per CTA (256 threads)

enum Section : int {
  SEC_INIT = 0,  // w   = A_IG x_d                 (rows: interior nodes touching the interface)
  SEC_FWD,       // L y = w      on the CTA's own subtrees, level by level
  SEC_PART,      // -sum_{j in own} L[i,j] y_j     for top rows i (combined by rank 0)
  SEC_TOPF,      // forward solve on the top nodes (rank 0)
  SEC_TOPB,      // backward solve on the top nodes (rank 0)
  SEC_BWD,       // L^T v = y    on the CTA's own subtrees (top values replicated)
  SEC_GG,        // out = A_GG x_d                 (rank 0)
  SEC_OUT,       // out -= A_IG^T v  restricted to the CTA's own columns
  NSEC
};

// A slice = up to 32 >> tlog2 consecutive row records processed by one warp with 1 << tlog2
// lanes per row.
struct SliceHdr {
  uint32_t row0;   // first row record (relative to the warp's row stream)
  uint16_t level;  // dependency level inside the section
  uint8_t tlog2;   // bits 0-2: log2(lanes per row); bit 7: "fresh" rows, S[dst] = -sign*acc/diag
  uint8_t nrows;
};
static_assert(sizeof(SliceHdr) == 8, "SliceHdr must be 8 bytes");

struct alignas(16) RowRec {
  uint32_t ebeg;  // first entry (relative to the warp's entry stream)
  uint16_t len;
  uint16_t dst;   // index into the CTA's shared vector S
  double diag;
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
  int64_t out_g_off;   // out_part pool: this CTA's n_gd partial outputs
  uint64_t slice_base[kWarps], row_base[kWarps], ent_base[kWarps];
  uint32_t n_ent[kWarps];
  uint32_t sec_begin[NSEC][kWarps], sec_end[NSEC][kWarps];
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

struct DomainPlan {
  std::vector<CtaDesc> ctas;  // csize descriptors (offsets local to this plan's pools)
  std::vector<SliceHdr> slices;
  std::vector<RowRec> rows;
  std::vector<double> vals;
  std::vector<uint16_t> cols;
  std::vector<int32_t> imap;  // ownmap / topmap entries
  // statistics
  int64_t nnz_L = 0;          // strictly-lower nnz of L
  int32_t n_top = 0, max_levels = 0;
  int64_t max_cta_entries = 0, top_entries = 0;
};

struct BuildOptions {
  int csize = 1;
  int target_k = 4;           // aimed entries per lane per row
  int min_block = 4;          // invert dense diagonal blocks of supernode chains at least this long (0: off)
  int max_block = 1024;
  int max_smem_doubles = 28000;
};

// Throws std::runtime_error on malformed input (missing diagonal, not lower triangular, ...).
void build_domain_plan(const DomainIn& in, const BuildOptions& opt, DomainPlan& out);

// Section signs per mode (see kernels.cu)
enum Mode : int { MODE_APPLY = 0, MODE_RHS = 1, MODE_BACKSUB = 2, MODE_UNIT = 3 };

}  // namespace kb

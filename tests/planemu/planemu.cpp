// CPU emulator of the device row-stream program (TEST TOOLING, not part of libkrylov_b200.so).
// Builds the plan with the product's plan.cpp and interprets slices / row records exactly as
// csrc/kernels.cu does (same sections, same level barriers -- emulated with snapshot semantics so a
// row that reads a value written in its own level is caught), so the host analysis can be
// validated without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../../julia-phd-krylov-spdes_b200/csrc/plan.h"

using namespace kb;

static void to_csc(Csc& m, int nrow, int ncol, const int32_t* cp, const int32_t* ri, const double* v) {
  m.nrow = nrow; m.ncol = ncol;
  m.colptr.assign(cp, cp + ncol + 1);
  m.rowval.assign(ri, ri + cp[ncol]);
  m.nzval.assign(v, v + cp[ncol]);
}

struct Emu {
  const DomainPlan& P;
  std::vector<std::vector<double>> S;
  explicit Emu(const DomainPlan& p) : P(p) {
    for (auto& d : P.ctas) S.emplace_back((size_t)d.smem_doubles, 0.0);
  }
  void section(int c, int sec, double sign) {
    const CtaDesc& D = P.ctas[c];
    std::vector<double>& s = S[c];
    for (uint32_t lev = 0; lev < D.sec_nlevels[sec]; ++lev) {
      std::vector<std::pair<int, double>> writes;
      for (int w = 0; w < kWarps; ++w)
        for (uint32_t si = D.sec_begin[sec][w]; si < D.sec_end[sec][w]; ++si) {
          const SliceHdr& h = P.slices[D.slice_base[w] + si];
          if (h.level != lev) continue;
          const bool fresh = (h.tlog2 & 0x80) != 0;
          if (h.nrows > (32 >> (h.tlog2 & 7))) throw std::runtime_error("slice holds too many rows");
          for (int g = 0; g < h.nrows; ++g) {
            const RowRec& r = P.rows[D.row_base[w] + h.row0 + g];
            double acc = 0.0;
            for (uint32_t k = 0; k < r.len; ++k) {
              uint64_t e = D.ent_base[w] + r.ebeg + k;
              if (r.ebeg + k >= D.n_ent[w]) throw std::runtime_error("entry out of range");
              if (P.cols[e] >= (uint32_t)D.smem_doubles) throw std::runtime_error("col out of range");
              acc += P.vals[e] * s[P.cols[e]];
            }
            writes.emplace_back((int)r.dst, ((fresh ? 0.0 : s[r.dst]) - sign * acc) / r.diag);
          }
        }
      for (auto& wv : writes) s[wv.first] = wv.second;
    }
  }
};

extern "C" int planemu_run(int n_I, int n_gd, const int32_t* perm, const int32_t* Lp, const int32_t* Li,
                           const double* Lx, const int32_t* Gp, const int32_t* Gi, const double* Gx,
                           const int32_t* GGp, const int32_t* GGi, const double* GGx, int csize, int mode,
                           const double* xd, const double* bI, double* out, double* uI, double* stats,
                           char* err, int errlen) {
  try {
    DomainIn in;
    in.n_I = n_I; in.n_gd = n_gd;
    in.perm.assign(perm, perm + n_I);
    to_csc(in.L, n_I, n_I, Lp, Li, Lx);
    to_csc(in.AIG, n_I, n_gd, Gp, Gi, Gx);
    to_csc(in.AGG, n_gd, n_gd, GGp, GGi, GGx);
    BuildOptions opt;
    opt.csize = csize;
    if (const char* e = std::getenv("KB200_MIN_BLOCK")) opt.min_block = std::atoi(e);
    DomainPlan P;
    build_domain_plan(in, opt, P);
    Emu E(P);
    const int C = csize;
    const CtaDesc& D0 = P.ctas[0];
    for (int c = 0; c < C; ++c) {
      const CtaDesc& D = P.ctas[c];
      if (mode == MODE_APPLY || mode == MODE_BACKSUB)
        for (int l = 0; l < n_gd; ++l) E.S[c][D.xd_off + l] = xd[l];
      if (mode == MODE_RHS || mode == MODE_BACKSUB) {
        for (int k = 0; k < D.n_own; ++k) E.S[c][D.own_off + k] = bI[P.imap[D.ownmap_off + k]];
        if (c == 0)
          for (int k = 0; k < D.n_top; ++k) E.S[c][D.top_off + k] = bI[P.imap[D.topmap_off + k]];
      }
    }
    for (int c = 0; c < C; ++c) {
      if (mode == MODE_APPLY) E.section(c, SEC_INIT, -1.0);
      if (mode == MODE_BACKSUB) E.section(c, SEC_INIT, +1.0);
      E.section(c, SEC_FWD, 1.0);
      E.section(c, SEC_PART, 1.0);
    }
    // cluster sync; rank 0 combines
    for (int t = 0; t < D0.n_top; ++t) {
      double a = E.S[0][D0.top_off + t];
      for (int c = 0; c < C; ++c) a += E.S[c][P.ctas[c].part_off + t];
      E.S[0][D0.top_off + t] = a;
    }
    E.section(0, SEC_TOPF, 1.0);
    E.section(0, SEC_TOPB, 1.0);
    for (int c = 1; c < C; ++c)
      for (int t = 0; t < D0.n_top; ++t) E.S[c][P.ctas[c].top_off + t] = E.S[0][D0.top_off + t];
    for (int c = 0; c < C; ++c) {
      E.section(c, SEC_BWD, 1.0);
      if (mode == MODE_APPLY) E.section(c, SEC_GG, -1.0);
      if (mode == MODE_APPLY || mode == MODE_RHS) E.section(c, SEC_OUT, 1.0);
    }
    if (out) {
      for (int l = 0; l < n_gd; ++l) {
        double a = 0.0;
        for (int c = 0; c < C; ++c) a += E.S[c][P.ctas[c].out_off + l];
        out[l] = a;
      }
    }
    if (uI) {
      for (int c = 0; c < C; ++c) {
        const CtaDesc& D = P.ctas[c];
        for (int k = 0; k < D.n_own; ++k) uI[P.imap[D.ownmap_off + k]] = E.S[c][D.own_off + k];
        if (c == 0)
          for (int k = 0; k < D.n_top; ++k) uI[P.imap[D.topmap_off + k]] = E.S[c][D.top_off + k];
      }
    }
    if (stats) {
      stats[0] = (double)P.nnz_L;
      stats[1] = (double)P.n_top;
      stats[2] = (double)P.max_levels;
      stats[3] = (double)P.max_cta_entries;
      stats[4] = (double)P.vals.size();
      stats[5] = (double)P.slices.size();
      stats[6] = (double)P.rows.size();
      stats[7] = (double)D0.smem_doubles;
      double topent = 0;
      for (int w = 0; w < kWarps; ++w)
        for (int sec : {SEC_TOPF, SEC_TOPB})
          for (uint32_t si = D0.sec_begin[sec][w]; si < D0.sec_end[sec][w]; ++si) {
            const SliceHdr& h = P.slices[D0.slice_base[w] + si];
            for (int g = 0; g < h.nrows; ++g) topent += P.rows[D0.row_base[w] + h.row0 + g].len;
          }
      stats[8] = topent;
      stats[9] = (double)std::max(D0.sec_nlevels[SEC_TOPF], D0.sec_nlevels[SEC_TOPB]);
      stats[10] = (double)std::max(D0.sec_nlevels[SEC_FWD], D0.sec_nlevels[SEC_BWD]);
    }
    return 0;
  } catch (const std::exception& e) {
    if (err) std::snprintf(err, (size_t)errlen, "%s", e.what());
    return -1;
  }
}

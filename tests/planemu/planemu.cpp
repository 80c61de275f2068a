// CPU emulator of the device row-stream program (TEST TOOLING, not part of libkrylov_b200.so).
// Builds the plan with the product's plan.cpp and interprets slices / row records exactly as
// csrc/kernels.cu does (same sections, same level barriers -- emulated with snapshot semantics so a
// row that reads a value written in its own level is caught), so the host analysis can be
// validated without a GPU.
#include <array>
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
  int NW;
  std::vector<std::vector<double>> S;
  long n_barriers = 0, n_slices = 0;
  Emu(const DomainPlan& p, int nw) : P(p), NW(nw) {
    for (auto& d : P.ctas) S.emplace_back((size_t)d.smem_doubles, 0.0);
  }
  // position of one consumer warp in its program
  struct Cursor {
    uint32_t ch, tok, end;   // chunk, next token, one past the warp's last token in the chunk
    bool done;
  };
  // Runs one section on the CTAs `ctas` in lock-step, exactly like the consumer warps: every warp walks its
  // own token list chunk by chunk; a barrier token ends an EPOCH (CTA barrier, or cluster barrier when the
  // section runs on several CTAs).  Checks: chunks / slices well formed; all warps (and CTAs) see the same
  // number of barriers; inside an epoch no slot is written by two warps, no warp reads a slot another
  // warp writes, and a slice does not read what it writes (same-warp slices are ordered by __syncwarp).
  void section(const std::vector<int>& ctas, int sec, double sign) {
    const size_t nc = ctas.size();
    std::vector<std::vector<Cursor>> cur(nc, std::vector<Cursor>(NW));
    auto chunk_ptr = [&](int c, uint32_t ch) { return P.stream.data() + P.ctas[c].stream_off + (size_t)ch * kChunkBytes; };
    auto enter_chunk = [&](int c, int w, Cursor& k) {
      const CtaDesc& D = P.ctas[c];
      for (; k.ch < D.sec_chunk1[sec]; ++k.ch) {
        const uint16_t* dir = reinterpret_cast<const uint16_t*>(chunk_ptr(c, k.ch));
        if (dir[0] != 0) throw std::runtime_error("chunk directory does not start at 0");
        for (int v = 0; v < NW; ++v)
          if (dir[v] > dir[v + 1]) throw std::runtime_error("chunk directory not monotone");
        if ((size_t)kChunkDirBytes + (size_t)dir[NW] * 2 > (size_t)kChunkBytes) throw std::runtime_error("token table too large");
        {
          const size_t gch = (size_t)(chunk_ptr(c, k.ch) - P.stream.data()) / kChunkBytes;
          if (gch >= P.chunk_units.size() || (size_t)kChunkDirBytes + (size_t)dir[NW] * 2 > (size_t)P.chunk_units[gch] * 16)
            throw std::runtime_error("token table beyond the copied part of its chunk");
        }
        k.tok = dir[w];
        k.end = dir[w + 1];
        if (k.tok < k.end) return;
      }
      k.done = true;
    };
    for (size_t i = 0; i < nc; ++i)
      for (int w = 0; w < NW; ++w) {
        cur[i][w] = Cursor{P.ctas[ctas[i]].sec_chunk0[sec], 0, 0, false};
        enter_chunk(ctas[i], w, cur[i][w]);
      }
    struct Todo { size_t ci; int w; const uint8_t* chunk; const uint8_t* sl; };
    for (;;) {
      // gather the slices of this epoch, warp by warp
      std::vector<Todo> todo;
      int at_barrier = 0, finished = 0;
      for (size_t i = 0; i < nc; ++i)
        for (int w = 0; w < NW; ++w) {
          Cursor& k = cur[i][w];
          bool hit = false;
          while (!k.done && !hit) {
            const uint8_t* chunk = chunk_ptr(ctas[i], k.ch);
            const uint16_t* tok = reinterpret_cast<const uint16_t*>(chunk + kChunkDirBytes);
            const uint16_t t = tok[k.tok++];
            if (t == kTokBarrier) hit = true;
            else {
              if ((size_t)t * 16 < (size_t)kChunkDirBytes || (size_t)t * 16 + 32 > (size_t)kChunkBytes) throw std::runtime_error("slice offset out of range");
              todo.push_back(Todo{i, w, chunk, chunk + (size_t)t * 16});
            }
            if (k.tok >= k.end) {
              ++k.ch;
              enter_chunk(ctas[i], w, k);
            }
          }
          if (hit) ++at_barrier;
          else ++finished;
        }
      if (at_barrier != 0 && finished != 0) throw std::runtime_error("warps disagree on the number of barriers");
      // writers of the epoch
      std::vector<std::vector<int>> writer(P.ctas.size());
      for (size_t c = 0; c < P.ctas.size(); ++c) writer[c].assign(S[c].size(), -1);
      auto wid = [&](const Todo& t) { return (int)(t.ci * NW + t.w); };
      for (const Todo& t : todo) {
        const CtaDesc& D = P.ctas[ctas[t.ci]];
        SliceHdr sh;
        std::memcpy(&sh, t.sl, sizeof(sh));
        const int T = 1 << sh.tlog2;
        if (sh.tlog2 > 5 || sh.nrows * T > 32 || sh.nrows == 0) throw std::runtime_error("bad slice shape");
        const size_t bytes = sizeof(SliceHdr) + (size_t)sh.nrows * sizeof(RowRec) + (size_t)sh.nent * 8 +
                             ((size_t)sh.kmax + 1) / 2 * ((size_t)sh.nrows * T) * 4;
        if ((size_t)(t.sl - t.chunk) + bytes > (size_t)kChunkBytes) throw std::runtime_error("slice exceeds its chunk");
        // the producer copies only the used part of a chunk (DomainPlan::chunk_units)
        const size_t gch = (size_t)(t.chunk - P.stream.data()) / kChunkBytes;
        if (gch >= P.chunk_units.size() || (size_t)(t.sl - t.chunk) + bytes > (size_t)P.chunk_units[gch] * 16)
          throw std::runtime_error("slice beyond the copied part of its chunk");
        const RowRec* rr = reinterpret_cast<const RowRec*>(t.sl + sizeof(SliceHdr));
        for (int g = 0; g < sh.nrows; ++g) {
          if (rr[g].dst >= D.smem_doubles) throw std::runtime_error("dst out of range");
          auto claim = [&](int c) {
            if (writer[c][rr[g].dst] >= 0 && writer[c][rr[g].dst] != wid(t))
              throw std::runtime_error("two warps write the same slot between barriers");
            writer[c][rr[g].dst] = wid(t);
          };
          claim(ctas[t.ci]);
          if (rr[g].flags & kRowBcast) {
            if (nc == 1 && P.ctas.size() > 1) throw std::runtime_error("broadcast row outside a cluster-wide section");
            for (size_t c2 = 0; c2 < P.ctas.size(); ++c2)
              if ((int)c2 != ctas[t.ci]) claim((int)c2);
          }
        }
      }
      // execute
      for (const Todo& t : todo) {
        const int c = ctas[t.ci];
        const CtaDesc& D = P.ctas[c];
        std::vector<double>& s = S[c];
        SliceHdr sh;
        std::memcpy(&sh, t.sl, sizeof(sh));
        const int T = 1 << sh.tlog2;
        const RowRec* rr = reinterpret_cast<const RowRec*>(t.sl + sizeof(SliceHdr));
        const double* vals = reinterpret_cast<const double*>(t.sl + sizeof(SliceHdr) + sh.nrows * sizeof(RowRec));
        const uint32_t* cols2 = reinterpret_cast<const uint32_t*>(vals + sh.nent);   // a lane's columns of two steps per word
        if (reinterpret_cast<const uint8_t*>(cols2) != t.sl + sh.cols_off) throw std::runtime_error("cols_off mismatch");
        std::vector<std::array<double, 4>> part(32, std::array<double, 4>{0, 0, 0, 0});
        const uint32_t Wact = (uint32_t)sh.nrows * T;
        if ((uint32_t)sh.kmax * Wact != sh.nent) throw std::runtime_error("slice entry count mismatch");
        std::vector<char> mine_dst(s.size(), 0);
        for (int g = 0; g < sh.nrows; ++g) {
          if (mine_dst[rr[g].dst]) throw std::runtime_error("destination written twice in a slice");
          mine_dst[rr[g].dst] = 1;
        }
        for (uint32_t k = 0; k < sh.kmax; ++k)
          for (uint32_t lane = 0; lane < Wact; ++lane) {
            const int g = lane / T, tt = lane % T;
            const uint32_t pos = k * Wact + lane;
            const uint32_t word = cols2[(k / 2) * Wact + lane];
            const uint32_t col = (k & 1) ? (word >> 16) : (word & 0xffffu);
            if (col >= (uint32_t)D.smem_doubles) throw std::runtime_error("col out of range");
            if (k * T + tt < rr[g].len) {
              if (mine_dst[col]) throw std::runtime_error("a slice reads a value it produces");
              if (writer[c][col] >= 0 && writer[c][col] != wid(t))
                throw std::runtime_error("reads a slot another warp writes before the next barrier");
            } else if (vals[pos] != 0.0) {
              throw std::runtime_error("non-zero padding entry");
            }
            part[lane][k & 3] += vals[pos] * s[col];
          }
        for (int g = 0; g < sh.nrows; ++g) {
          if ((rr[g].len + T - 1) / T > sh.kmax) throw std::runtime_error("kmax too small");
          double sum = 0.0;
          for (int tt = 0; tt < T; ++tt) sum += (part[g * T + tt][0] + part[g * T + tt][1]) + (part[g * T + tt][2] + part[g * T + tt][3]);
          const int dst = rr[g].dst;
          const bool fresh = (rr[g].flags & kRowFresh) != 0;
          s[dst] = ((fresh ? 0.0 : s[dst]) - sign * sum) * rr[g].diag;  // reciprocal stored
          if (rr[g].flags & kRowBcast)
            for (size_t c2 = 0; c2 < S.size(); ++c2) S[c2][dst] = s[dst];
        }
        ++n_slices;
      }
      if (at_barrier == 0) break;
      ++n_barriers;
    }
  }
  void section(int c, int sec, double sign) { section(std::vector<int>{c}, sec, sign); }
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
    if (const char* e = std::getenv("KB200_WARPS")) opt.warps = std::atoi(e);
    if (const char* e = std::getenv("KB200_MERGE_DEPTH")) opt.merge_depth = std::atoi(e);
    if (const char* e = std::getenv("KB200_PAD_TOL_PCT")) opt.pad_tol = std::atoi(e) / 100.0;
    if (const char* e = std::getenv("KB200_LANE_ENTRIES")) opt.lane_entries = std::atoi(e);
    if (const char* e = std::getenv("KB200_ADAPTIVE_LANES")) opt.adaptive_lanes = std::atoi(e);
    if (const char* e = std::getenv("KB200_WARP_PRIVATE")) opt.warp_private = std::atoi(e);
    if (const char* e = std::getenv("KB200_PRIVATE_SPLIT")) opt.private_split = std::atoi(e);
    DomainPlan P;
    build_domain_plan(in, opt, P);
    const int NW = (opt.warps == 16 || opt.warps == 24) ? opt.warps : kWarps;
    Emu E(P, NW);
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
    // cluster sync; every CTA forms the combined top right-hand side; top levels run over all CTAs
    if (C > 1) {
      std::vector<double> comb(D0.n_top);
      for (int t = 0; t < D0.n_top; ++t) {
        double a = E.S[0][D0.top_off + t];
        for (int c = 0; c < C; ++c) a += E.S[c][P.ctas[c].part_off + t];
        comb[t] = a;
      }
      for (int c = 0; c < C; ++c)
        for (int t = 0; t < D0.n_top; ++t) E.S[c][P.ctas[c].top_off + t] = comb[t];
      std::vector<int> all;
      for (int c = 0; c < C; ++c) all.push_back(c);
      E.section(all, SEC_TOPF, 1.0);
      E.section(all, SEC_TOPB, 1.0);
    }
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
      stats[4] = (double)P.n_entries;
      stats[5] = (double)P.n_packets;
      stats[6] = (double)P.n_rows;
      stats[7] = (double)D0.smem_doubles;
      stats[8] = (double)P.stream.size();
      {
        double rb = 0.0;
        for (uint16_t u : P.chunk_units) rb += 16.0 * u;
        if (std::getenv("KB200_STATS_READ_BYTES")) stats[8] = rb;   // bytes the producer copies
      }
      stats[11] = (double)P.lane_steps;
      stats[12] = (double)P.n_private_rows;
      stats[13] = (double)P.n_private_slices;
      stats[14] = (double)E.n_barriers;
      stats[15] = (double)E.n_slices;
      stats[9] = (double)std::max(D0.sec_nlevels[SEC_TOPF], D0.sec_nlevels[SEC_TOPB]);
      stats[10] = (double)std::max(D0.sec_nlevels[SEC_FWD], D0.sec_nlevels[SEC_BWD]);

    }
    return 0;
  } catch (const std::exception& e) {
    if (err) std::snprintf(err, (size_t)errlen, "%s", e.what());
    return -1;
  }
}

// host-side bookkeeping of the multi-GPU exchange (csrc/xchg_seq.h), exposed for the CPU tests
// CPU emulation of the owner-ordered compact interface exchange (csrc/xchg_tables.h + the arithmetic of
// exchange_gather in csrc/kernels.cuh) for `world` ranks: loc_all = concatenation over ALL subdomains of
// their local interface vectors; y_out[rank * ng + g] = what rank `rank` assembles.  Returns the number
// of gathered entries (all ranks), or -1 on error.
#include "../../julia-phd-krylov-spdes_b200/csrc/xchg_tables.h"
extern "C" int64_t planemu_exchange(int world, int64_t ndom, int64_t ng, const int64_t* n_gd, const int32_t* l2g_cat,
                                    const double* loc_all, const double* base, int threads, double* y_out,
                                    int64_t* sent_per_rank) {
  try {
    std::vector<std::vector<int32_t>> l2g(ndom);
    std::vector<int64_t> dom_off(ndom + 1, 0);
    for (int64_t d = 0; d < ndom; ++d) {
      l2g[d].assign(l2g_cat + dom_off[d], l2g_cat + dom_off[d] + n_gd[d]);
      dom_off[d + 1] = dom_off[d] + n_gd[d];
    }
    std::vector<const std::vector<int32_t>*> lp(ndom);
    for (int64_t d = 0; d < ndom; ++d) lp[d] = &l2g[d];
    std::vector<XchTables> T;
    for (int r = 0; r < world; ++r) T.push_back(build_xchg_tables(world, r, lp, ng, threads));
    std::vector<double> gathered((size_t)T[0].m_total, 0.0);
    for (int r = 0; r < world; ++r) {
      if (T[r].m_total != T[0].m_total || T[r].fin_src != T[0].fin_src || T[r].off != T[0].off) return -1;  // same layout on every rank
      const double* loc = loc_all + dom_off[(ndom * r) / world];
      const int m = T[r].off[r + 1] - T[r].off[r];
      if (sent_per_rank) sent_per_rank[r] = m;
      for (int e = 0; e < m; ++e) {
        double a = (base != nullptr && T[r].snd_g[e] >= 0) ? base[T[r].snd_g[e]] : 0.0;
        for (int k = T[r].snd_ptr[e]; k < T[r].snd_ptr[e + 1]; ++k) a += loc[T[r].snd_src[k]];
        gathered[(size_t)T[r].off[r] + e] = a;
      }
    }
    for (int r = 0; r < world; ++r)
      for (int64_t g = 0; g < ng; ++g) {
        double s = 0.0;
        for (int k = T[r].fin_ptr[g]; k < T[r].fin_ptr[g + 1]; ++k) s += gathered[T[r].fin_src[k]];
        y_out[(size_t)r * ng + g] = s;
      }
    return T[0].m_total;
  } catch (const std::exception&) {
    return -1;
  }
}

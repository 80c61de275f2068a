// Host-side plan builder: see plan.h for the design.
#include "plan.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <queue>
#include <stdexcept>

namespace kb {
namespace {

struct TmpRow {
  uint16_t dst;
  double diag;
  int32_t level;
  int32_t kind;  // 0: S[dst] = (S[dst] - sign*acc)/diag ; 1: S[dst] = -sign*acc/diag
  int32_t bcast; // result is replicated into every CTA of the cluster (top nodes)
  int32_t warp;  // consumer warp that owns the row (private phase), -1: rows of a CTA-wide level
  uint32_t ebeg;  // into the builder's scratch entry arrays
  uint32_t len;
  uint32_t diag_src;  // pool entry the divisor comes from (reciprocal), kNoSrc: a constant
};

constexpr uint32_t kNoSrc = 0xffffffffu;

struct Scratch {
  std::vector<uint16_t> cols;
  std::vector<double> vals;
  std::vector<uint32_t> srcs;  // pool entry of every value (value refresh, plan.h)
};

// a serialised slice + where its values come from (slot offsets relative to the slice start)
struct SliceBytes {
  std::vector<uint8_t> b;
  std::vector<ValRef> refs;
};

inline int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
inline int ilog2(int p) {
  int l = 0;
  while ((1 << l) < p) ++l;
  return l;
}

// Byte stream of one CTA.
struct CtaStream {
  std::vector<uint8_t> bytes;
  std::vector<ValRef> refs;  // slot offsets relative to the CTA stream
  std::vector<uint16_t> chunk_units;  // used 16-byte units of every chunk (see DomainPlan::chunk_units)
  int64_t n_entries = 0, n_rows = 0, n_packets = 0, lane_steps = 0;
  size_t room() const { return kChunkBytes - (bytes.size() % kChunkBytes); }
  void pad_to_chunk() {
    if (bytes.size() % kChunkBytes) bytes.resize(bytes.size() + room(), 0);
  }
  template <class T>
  void put(const T& v) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    bytes.insert(bytes.end(), p, p + sizeof(T));
  }
};

// Serialise one slice (rows ord[p..q) of `rows`, T lanes per row) -- layout in plan.h.
SliceBytes make_slice(const std::vector<TmpRow>& rows, const Scratch& sc, const BuildOptions& opt,
                      const std::vector<uint32_t>& ord, size_t p, size_t q, int T, CtaStream& cs) {
  const int nr = (int)(q - p);
  const uint32_t maxlen = rows[ord[p]].len;
  const uint32_t kmax = (maxlen + T - 1) / T;
  // entries in execution order: step k, the nr*T active lanes ascending; lanes whose row has
  // run out get a zero entry (rows of a slice have similar lengths, so this padding is small)
  std::vector<double> pv;
  std::vector<uint16_t> pc;
  std::vector<uint32_t> ps;
  int64_t real = 0;
  // WHICH entry of its row a lane takes at a step is free (any order of a row's terms is a valid sum; the order
  // chosen here is the association order of the plan, followed by the device and by the emulator alike).  It is
  // chosen against shared-memory bank conflicts of the gathered S[col]: a 64-bit load of a warp is served per
  // half-warp, in as many passes as distinct 8-byte words of one bank pair (slot mod 16) are requested; lanes
  // reading the same slot share a pass.  Greedy: every lane takes the first remaining entry of its row whose
  // bank pair is still free in its half-warp at this step (or whose slot is already being read there).
  // Position (k, t) of a row holds a real entry iff k * T + t < len, as before; padding entries repeat a slot of
  // their half-warp (value 0).
  std::vector<std::vector<uint32_t>> left(nr);   // remaining entries of every row, original order
  for (int g = 0; g < nr; ++g) {
    const TmpRow& tr = rows[ord[p + g]];
    left[g].resize(tr.len);
    for (uint32_t e = 0; e < tr.len; ++e) left[g][e] = tr.ebeg + e;
  }
  for (uint32_t k = 0; k < kmax; ++k) {
    int owner_slot[2][16];
    for (auto& h : owner_slot)
      for (int& v : h) v = -1;
    const size_t step0 = pv.size();
    std::vector<int> pad_lanes;
    for (int lane = 0; lane < nr * T; ++lane) {
      const int g = lane / T, t = lane % T, half = lane / 16;
      const TmpRow& tr = rows[ord[p + g]];
      const uint32_t idx = k * T + t;
      if (idx < tr.len) {
        std::vector<uint32_t>& rem = left[g];
        size_t pick = 0;
        if (opt.avoid_bank_conflicts) {
          for (size_t q = 0; q < rem.size(); ++q) {
            const int col = sc.cols[rem[q]];
            const int o = owner_slot[half][col & 15];
            if (o < 0 || o == col) {
              pick = q;
              break;
            }
          }
        }
        const uint32_t e = rem[pick];
        rem.erase(rem.begin() + (std::ptrdiff_t)pick);
        const int col = sc.cols[e];
        if (owner_slot[half][col & 15] < 0) owner_slot[half][col & 15] = col;
        pv.push_back(sc.vals[e]);
        pc.push_back(sc.cols[e]);
        ps.push_back(sc.srcs[e]);
        ++real;
      } else {
        pv.push_back(0.0);
        pc.push_back(0);
        ps.push_back(kNoSrc);
        pad_lanes.push_back(lane);
      }
    }
    if (opt.avoid_bank_conflicts)
      for (int lane : pad_lanes) {   // a padding lane re-reads a slot its half-warp reads anyway
        const int half = lane / 16;
        for (int l2 = half * 16; l2 < std::min(nr * T, half * 16 + 16); ++l2)
          if (ps[step0 + (size_t)l2] != kNoSrc) {
            pc[step0 + (size_t)lane] = pc[step0 + (size_t)l2];
            break;
          }
      }
  }
  if (pv.size() > 65535u || kmax > 65535u) throw std::runtime_error("internal: slice too large");
  SliceBytes out;
  std::vector<uint8_t>& b = out.b;
  auto put = [&](const void* ptr, size_t nb) {
    const uint8_t* u = static_cast<const uint8_t*>(ptr);
    b.insert(b.end(), u, u + nb);
  };
  SliceHdr sh;
  std::memset(&sh, 0, sizeof(sh));
  sh.nent = (uint16_t)pv.size();
  sh.kmax = (uint16_t)kmax;
  sh.tlog2 = (uint8_t)ilog2(T);
  sh.nrows = (uint8_t)nr;
  sh.width = (uint32_t)(nr * T);
  sh.cols_off = (uint32_t)(sizeof(SliceHdr) + (size_t)nr * sizeof(RowRec) + pv.size() * 8);
  put(&sh, sizeof(sh));
  for (int g = 0; g < nr; ++g) {
    const TmpRow& tr = rows[ord[p + g]];
    RowRec rr;
    std::memset(&rr, 0, sizeof(rr));
    rr.len = (uint16_t)tr.len;
    rr.dst = tr.dst;
    rr.flags = (uint16_t)((tr.kind ? kRowFresh : 0) | (tr.bcast ? kRowBcast : 0));
    rr.diag = 1.0 / tr.diag;  // the device multiplies by the reciprocal
    if (tr.diag_src != kNoSrc) out.refs.push_back(ValRef{(uint32_t)((b.size() + 8) / 8), tr.diag_src | kRefRecip});
    put(&rr, sizeof(rr));
  }
  if (!pv.empty()) {
    const uint32_t v0 = (uint32_t)(b.size() / 8);
    for (size_t k = 0; k < ps.size(); ++k)
      if (ps[k] != kNoSrc) out.refs.push_back(ValRef{v0 + (uint32_t)k, ps[k]});
    put(pv.data(), pv.size() * 8);
    // columns: one 32-bit word per lane and pair of steps (low half = even step; an odd last step is repeated)
    const size_t Wd = (size_t)nr * T;
    std::vector<uint32_t> pc2(((size_t)kmax + 1) / 2 * Wd);
    for (size_t pr = 0; pr < ((size_t)kmax + 1) / 2; ++pr)
      for (size_t lane = 0; lane < Wd; ++lane) {
        const uint32_t lo = pc[(2 * pr) * Wd + lane];
        const uint32_t hi = (2 * pr + 1 < kmax) ? pc[(2 * pr + 1) * Wd + lane] : lo;
        pc2[pr * Wd + lane] = lo | (hi << 16);
      }
    put(pc2.data(), pc2.size() * 4);
  }
  while (b.size() % 16) b.push_back(0);
  if (b.size() + (size_t)kChunkDirBytes + 16 > (size_t)kChunkBytes)
    throw std::runtime_error("internal: slice exceeds a chunk");
  cs.n_entries += real;
  cs.n_rows += nr;
  cs.n_packets += 1;
  cs.lane_steps += 32 * (int64_t)kmax;
  return out;
}

// Chunk under construction: per-warp token lists + slice payloads (see plan.h, "warp programs").
struct ChunkBuilder {
  int NW;
  CtaStream& cs;
  std::vector<std::vector<uint16_t>> tok;          // per warp: slice index (into `slices`) or kTokBarrier
  std::vector<SliceBytes> slices;
  size_t payload = 0, ntok = 0;
  ChunkBuilder(int nw, CtaStream& s) : NW(nw), cs(s), tok(nw) {}
  static size_t table_bytes(size_t ntokens) { return (size_t)kChunkDirBytes + ((ntokens * 2 + 15) / 16) * 16; }
  bool empty() const { return ntok == 0; }
  bool fits(size_t more_tokens, size_t more_payload) const {
    return table_bytes(ntok + more_tokens) + payload + more_payload <= (size_t)kChunkBytes;
  }
  void flush() {
    if (empty()) return;
    if (cs.bytes.size() % kChunkBytes) throw std::runtime_error("internal: chunk builder out of phase");
    const size_t base = cs.bytes.size();
    const size_t tb = table_bytes(ntok);
    std::vector<size_t> off(slices.size());
    size_t o = tb;
    for (size_t k = 0; k < slices.size(); ++k) {
      off[k] = o;
      o += slices[k].b.size();
    }
    if (o > (size_t)kChunkBytes) throw std::runtime_error("internal: chunk overflow");
    std::vector<uint16_t> dir(kChunkDirBytes / 2, 0);
    size_t acc = 0;
    for (int w = 0; w < NW; ++w) {
      dir[w] = (uint16_t)acc;
      acc += tok[w].size();
    }
    dir[NW] = (uint16_t)acc;
    for (uint16_t v : dir) cs.put(v);
    size_t written = 0;
    for (int w = 0; w < NW; ++w)
      for (uint16_t t : tok[w]) {
        cs.put(t == kTokBarrier ? kTokBarrier : (uint16_t)(off[t] / 16));
        ++written;
      }
    while ((written * 2) % 16) {
      cs.put((uint16_t)0);
      ++written;
    }
    for (size_t k = 0; k < slices.size(); ++k) {
      const uint32_t at8 = (uint32_t)((base + off[k]) / 8);   // slices start on 16-byte boundaries
      for (const ValRef& r : slices[k].refs) cs.refs.push_back(ValRef{at8 + r.off8, r.src});
      cs.bytes.insert(cs.bytes.end(), slices[k].b.begin(), slices[k].b.end());
    }
    if (cs.bytes.size() - base != o) throw std::runtime_error("internal: chunk size mismatch");
    if (cs.chunk_units.size() != base / kChunkBytes) throw std::runtime_error("internal: chunk table out of phase");
    cs.chunk_units.push_back((uint16_t)((o + 15) / 16));
    cs.pad_to_chunk();
    for (auto& t : tok) t.clear();
    slices.clear();
    payload = 0;
    ntok = 0;
  }
  void add_slice(int warp, SliceBytes&& sb) {
    if (!fits(1, sb.b.size())) flush();
    if (!fits(1, sb.b.size())) throw std::runtime_error("internal: slice exceeds a chunk");
    if (slices.size() >= 0xfffe) throw std::runtime_error("internal: too many slices in a chunk");
    tok[warp].push_back((uint16_t)slices.size());
    payload += sb.b.size();
    slices.push_back(std::move(sb));
    ++ntok;
  }
  void add_barrier() {  // one barrier token in every warp's program
    if (!fits((size_t)NW, 0)) flush();
    for (int w = 0; w < NW; ++w) tok[w].push_back(kTokBarrier);
    ntok += (size_t)NW;
  }
};

// Serialise the rows of one section as warp programs (see plan.h).  Rows without a warp (`warp` < 0)
// form CTA-wide dependency levels: the slices of a level are balanced over the NW consumer warps and
// every warp's program gets a barrier token after each level (`force_levels`: top sections, where
// every CTA of the cluster must execute the same number of barriers, also for levels in which it has
// no rows).  Warp-private rows (`warp` >= 0) go to their warp in level order without barriers, before
// (private_first, forward sweep) or after (backward sweep) the shared levels, fenced by one barrier.
void emit_section(std::vector<TmpRow>& rows, const Scratch& sc, const BuildOptions& opt, CtaStream& cs,
                  uint32_t& chunk0, uint32_t& chunk1, uint32_t& nlevels, int force_levels = -1,
                  bool private_first = true, int64_t* n_private_rows = nullptr, int64_t* n_private_slices = nullptr) {
  cs.pad_to_chunk();
  chunk0 = (uint32_t)(cs.bytes.size() / kChunkBytes);
  const int NW = (opt.warps == 16 || opt.warps == 24) ? opt.warps : kWarps;
  int maxlev = -1;
  bool any_private = false;
  for (auto& r : rows) {
    if (r.level < 0) throw std::runtime_error("internal: negative level");
    if (r.warp >= NW) throw std::runtime_error("internal: warp out of range");
    if (r.warp >= 0) any_private = true;
    else maxlev = std::max(maxlev, r.level);
  }
  nlevels = (uint32_t)(maxlev + 1);
  if (force_levels >= 0) {
    if (force_levels < maxlev + 1) throw std::runtime_error("internal: forced level count too small");
    if (any_private) throw std::runtime_error("internal: private rows in a section with forced levels");
    nlevels = (uint32_t)force_levels;
  }
  int lane_entries = std::max(1, opt.lane_entries);  // entries a lane of a row aims at (set per level below)
  auto t_of = [&](uint32_t len) {
    const int want = (int)((len + lane_entries - 1) / lane_entries);
    return std::min(32, pow2ceil(std::max(1, want)));
  };
  // slices of the rows ord[i..j) (one level, sorted by decreasing length)
  auto level_slices = [&](const std::vector<uint32_t>& ord, size_t i, size_t j, std::vector<SliceBytes>& slices) {
    size_t p = i;
    while (p < j) {
      const int T = t_of(rows[ord[p]].len);
      const int cap = 32 / T;
      size_t q = p;
      // rows of a slice are padded to its longest row: stop when the next row is too much shorter
      const uint32_t len0 = rows[ord[p]].len;
      const uint32_t min_len = (uint32_t)(opt.pad_tol * (double)len0);
      while (q < j && (int)(q - p) < cap && t_of(rows[ord[q]].len) == T &&
             (rows[ord[q]].len >= min_len || len0 - rows[ord[q]].len <= 1))
        ++q;
      slices.push_back(make_slice(rows, sc, opt, ord, p, q, T, cs));
      p = q;
    }
  };
  auto sorted_order = [&](auto keep) {
    std::vector<uint32_t> ord;
    for (uint32_t k = 0; k < rows.size(); ++k)
      if (keep(rows[k])) ord.push_back(k);
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
      if (rows[a].level != rows[b].level) return rows[a].level < rows[b].level;
      return rows[a].len > rows[b].len;
    });
    return ord;
  };
  ChunkBuilder cb(NW, cs);

  auto emit_shared = [&]() {
    const std::vector<uint32_t> ord = sorted_order([](const TmpRow& r) { return r.warp < 0; });
    std::vector<SliceBytes> slices;  // serialised slices of the current level
    size_t i = 0;
    for (int lev = 0; lev < (int)nlevels; ++lev) {
      size_t j = i;
      while (j < ord.size() && rows[ord[j]].level == lev) ++j;
      slices.clear();
      if (j > i) {
        if (opt.adaptive_lanes) {
          // spread the level's entries over all NW warps (about one slice per warp), between lane_min
          // and lane_max entries per lane
          int64_t ent = 0;
          for (size_t k = i; k < j; ++k) ent += rows[ord[k]].len;
          const int64_t per_lane = (ent + (int64_t)NW * 32 - 1) / ((int64_t)NW * 32);
          lane_entries = (int)std::min<int64_t>(std::max<int64_t>(per_lane, opt.lane_min), opt.lane_max);
        }
        level_slices(ord, i, j, slices);
        lane_entries = std::max(1, opt.lane_entries);
      }
      // balance the warps of the level: longest-processing-time first on an estimate of the slice latency
      auto cost = [](const SliceBytes& sb) {
        SliceHdr sh;
        std::memcpy(&sh, sb.b.data(), sizeof(sh));
        return 150 + 30 * (int)sh.kmax + 40 * (int)sh.tlog2;
      };
      std::vector<size_t> so(slices.size());
      std::iota(so.begin(), so.end(), (size_t)0);
      std::stable_sort(so.begin(), so.end(), [&](size_t a, size_t b) { return cost(slices[a]) > cost(slices[b]); });
      std::vector<int64_t> load(NW, 0);
      std::vector<int> owner(slices.size(), 0);
      // rotate the starting warp with the level so that no warp always gets the heaviest slice
      for (size_t k : so) {
        int best = lev % NW;
        for (int d = 1; d < NW; ++d) {
          const int w = (lev + d) % NW;
          if (load[w] < load[best]) best = w;
        }
        load[best] += cost(slices[k]);
        owner[k] = best;
      }
      // Placement into chunks.  A slice cannot straddle a chunk boundary, and the stream is HBM traffic: what is
      // left of a chunk when the next slice does not fit is read for nothing (12 % of the stream when slices are
      // placed strictly in order).  The order of the slices of a level inside the stream is free (a warp finds
      // its slices through its token list), so the tail of a chunk takes the heaviest remaining slice of the
      // level that still fits (first fit, decreasing cost); a chunk is only closed when none does.
      if (opt.fill_chunks) {
        std::vector<char> placed(slices.size(), 0);
        size_t left = slices.size(), first = 0;
        while (left > 0) {
          bool any = false;
          for (size_t q = first; q < so.size(); ++q) {
            const size_t k = so[q];
            if (placed[k] || !cb.fits(1, slices[k].b.size())) continue;
            cb.add_slice(owner[k], std::move(slices[k]));
            placed[k] = 1;
            --left;
            any = true;
          }
          while (first < so.size() && placed[so[first]]) ++first;
          if (!any) {
            if (cb.empty()) throw std::runtime_error("internal: slice exceeds a chunk");
            cb.flush();
          }
        }
      } else {
        for (size_t k : so) cb.add_slice(owner[k], std::move(slices[k]));
      }
      cb.add_barrier();
      i = j;
    }
    if (i != ord.size()) throw std::runtime_error("internal: rows beyond the last level");
  };

  auto emit_private = [&]() {
    // per warp: the slices of its rows in level order; the chunks take from each warp a share
    // proportional to what it has left, so that the warps advance through the ring at the same pace
    std::vector<std::vector<SliceBytes>> ws(NW);
    std::vector<size_t> tot(NW, 0), nxt(NW, 0), rem(NW, 0);
    for (int w = 0; w < NW; ++w) {
      const std::vector<uint32_t> ord = sorted_order([w](const TmpRow& r) { return r.warp == w; });
      size_t i = 0;
      while (i < ord.size()) {
        size_t j = i;
        while (j < ord.size() && rows[ord[j]].level == rows[ord[i]].level) ++j;
        level_slices(ord, i, j, ws[w]);
        i = j;
      }
      for (auto& b : ws[w]) tot[w] += b.b.size();
      rem[w] = tot[w];
      if (n_private_rows) *n_private_rows += (int64_t)ord.size();
      if (n_private_slices) *n_private_slices += (int64_t)ws[w].size();
    }
    for (;;) {
      int best = -1;
      double bf = -1.0;
      for (int w = 0; w < NW; ++w) {
        if (nxt[w] >= ws[w].size()) continue;
        const double f = (double)rem[w] / (double)tot[w];
        if (f > bf) {
          bf = f;
          best = w;
        }
      }
      if (best < 0) break;
      rem[best] -= ws[best][nxt[best]].b.size();
      cb.add_slice(best, std::move(ws[best][nxt[best]]));
      nxt[best]++;
    }
    cb.add_barrier();
  };

  if (any_private && private_first) emit_private();
  emit_shared();
  if (any_private && !private_first) emit_private();
  cb.flush();
  cs.pad_to_chunk();
  chunk1 = (uint32_t)(cs.bytes.size() / kChunkBytes);
}

}  // namespace

void build_domain_plan(const DomainIn& in, const BuildOptions& opt, DomainPlan& out) {
  const int n = in.n_I, ng = in.n_gd, C = std::max(1, opt.csize);
  const Csc& L = in.L;
  if (L.nrow != n || L.ncol != n) throw std::runtime_error("L has wrong dimensions");
  if ((int)in.perm.size() != n) throw std::runtime_error("perm has wrong length");

  // ---- structure of L: diagonal, strictly-lower CSR, elimination tree -------------------------
  std::vector<double> diag(n);
  std::vector<int32_t> rptr(n + 1, 0);
  for (int j = 0; j < n; ++j) {
    int p0 = L.colptr[j], p1 = L.colptr[j + 1];
    if (p0 >= p1 || L.rowval[p0] != j) throw std::runtime_error("L: missing diagonal entry");
    diag[j] = L.nzval[p0];
    if (!(diag[j] > 0.0)) throw std::runtime_error("L: non-positive diagonal");
    for (int p = p0 + 1; p < p1; ++p) {
      if (L.rowval[p] <= L.rowval[p - 1] || L.rowval[p] >= n)
        throw std::runtime_error("L: column rows must be sorted, unique and below the diagonal");
      rptr[L.rowval[p] + 1]++;
    }
  }
  for (int i = 0; i < n; ++i) rptr[i + 1] += rptr[i];
  const int64_t nnzl = rptr[n];
  std::vector<int32_t> rcol(nnzl), rsrc(nnzl);   // rsrc: position of the entry in L.nzval (value refresh)
  std::vector<double> rval(nnzl);
  {
    std::vector<int32_t> fill(rptr.begin(), rptr.end() - 1);
    for (int j = 0; j < n; ++j)
      for (int p = L.colptr[j] + 1; p < L.colptr[j + 1]; ++p) {
        int i = L.rowval[p];
        rcol[fill[i]] = j;
        rval[fill[i]] = L.nzval[p];
        rsrc[fill[i]] = p;
        fill[i]++;
      }
  }
  std::vector<int32_t> parent(n, -1);
  for (int j = 0; j < n; ++j)
    if (L.colptr[j] + 1 < L.colptr[j + 1]) parent[j] = L.rowval[L.colptr[j] + 1];

  // ---- subtree weights -------------------------------------------------------------------------
  std::vector<int64_t> wt(n), sub(n);
  for (int j = 0; j < n; ++j) {
    wt[j] = (int64_t)(rptr[j + 1] - rptr[j]) + (L.colptr[j + 1] - L.colptr[j] - 1) + 4;
    sub[j] = wt[j];
  }
  for (int j = 0; j < n; ++j)
    if (parent[j] >= 0) sub[parent[j]] += sub[j];
  std::vector<int32_t> cptr(n + 1, 0), child(n > 0 ? n : 1);
  for (int j = 0; j < n; ++j)
    if (parent[j] >= 0) cptr[parent[j] + 1]++;
  for (int j = 0; j < n; ++j) cptr[j + 1] += cptr[j];
  {
    std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1);
    for (int j = 0; j < n; ++j)
      if (parent[j] >= 0) child[fill[parent[j]]++] = j;
  }

  // ---- cut the forest into subtrees (bins) + top -----------------------------------------------
  std::vector<char> is_top(n, 0);
  std::vector<int32_t> bin_of_root(n, -1);
  {
    std::vector<int32_t> Q;
    for (int j = 0; j < n; ++j)
      if (parent[j] < 0) Q.push_back(j);
    auto lpt = [&](const std::vector<int32_t>& q, std::vector<int32_t>* assign) -> int64_t {
      std::vector<int32_t> o(q.size());
      std::iota(o.begin(), o.end(), 0);
      std::sort(o.begin(), o.end(), [&](int a, int b) {
        if (sub[q[a]] != sub[q[b]]) return sub[q[a]] > sub[q[b]];
        return q[a] < q[b];
      });
      std::vector<int64_t> load(C, 0);
      if (assign) assign->assign(q.size(), 0);
      for (int k : o) {
        int best = 0;
        for (int c = 1; c < C; ++c)
          if (load[c] < load[best]) best = c;
        load[best] += sub[q[k]];
        if (assign) (*assign)[k] = best;
      }
      return *std::max_element(load.begin(), load.end());
    };
    if (C > 1 && n > 0) {
      // cost of a cut = heaviest bin + serial top work; explore successive splits of the heaviest
      // subtree and keep the best cut seen
      std::vector<int32_t> order;  // nodes moved to top, in order
      int64_t top_w = 0;
      int64_t best_cost = lpt(Q, nullptr);
      size_t best_k = 0;
      std::vector<int32_t> Qw = Q;
      const size_t max_splits = (size_t)std::min<int64_t>(n / 4, 4096);
      for (size_t step = 0; step < max_splits; ++step) {
        size_t hi = 0;
        for (size_t k = 1; k < Qw.size(); ++k)
          if (sub[Qw[k]] > sub[Qw[hi]]) hi = k;
        int r = Qw[hi];
        if (cptr[r + 1] == cptr[r]) break;  // heaviest subtree is a single node
        Qw[hi] = Qw.back();
        Qw.pop_back();
        for (int p = cptr[r]; p < cptr[r + 1]; ++p) Qw.push_back(child[p]);
        order.push_back(r);
        // a top node costs its row and column once per sweep, processed by rank 0 alone while
        // the other CTAs wait (over-estimate: its entries in subtree columns run in parallel)
        top_w += 2 * wt[r];
        if (top_w / C >= best_cost) break;  // cannot improve any more
        if (cptr[r + 1] - cptr[r] >= 2) {  // the cost only drops at branching nodes
          int64_t cost = lpt(Qw, nullptr) + top_w / C;  // top rows are shared out over the C CTAs
          if (cost < best_cost) {
            best_cost = cost;
            best_k = order.size();
          }
        }
      }
      for (size_t k = 0; k < best_k; ++k) is_top[order[k]] = 1;
      Q.clear();
      for (int j = 0; j < n; ++j)
        if (!is_top[j] && (parent[j] < 0 || is_top[parent[j]])) Q.push_back(j);
    }
    std::vector<int32_t> assign;
    lpt(Q, &assign);
    for (size_t k = 0; k < Q.size(); ++k) bin_of_root[Q[k]] = assign[k];
  }
  std::vector<int32_t> owner(n);  // CTA rank, or -1 for top
  for (int j = n - 1; j >= 0; --j) {
    if (is_top[j]) owner[j] = -1;
    else if (parent[j] < 0 || is_top[parent[j]]) owner[j] = bin_of_root[j];
    else owner[j] = owner[parent[j]];
  }

  // ---- warp-private subtrees: the bottom of every CTA's own forest, cut into subtrees that are
  // bin-packed onto the consumer warps (wown = warp, or -1 for the "upper" own nodes and the top) ----
  std::vector<int32_t> wown(n, -1);
  const int NWp = (opt.warps == 16 || opt.warps == 24) ? opt.warps : kWarps;
  if (opt.warp_private && n > 0) {
    std::vector<char> upper(n, 0);
    std::vector<int32_t> wroot(n, -1);
    for (int c = 0; c < C; ++c) {
      std::vector<int32_t> Qw;
      int64_t total = 0;
      for (int j = 0; j < n; ++j)
        if (owner[j] == c && (parent[j] < 0 || owner[parent[j]] != c)) {
          Qw.push_back(j);
          total += sub[j];
        }
      if (Qw.empty()) continue;
      const int64_t target = std::max<int64_t>(total / ((int64_t)NWp * std::max(1, opt.private_split)), 64);
      for (;;) {
        size_t hi = 0;
        for (size_t k = 1; k < Qw.size(); ++k)
          if (sub[Qw[k]] > sub[Qw[hi]] || (sub[Qw[k]] == sub[Qw[hi]] && Qw[k] < Qw[hi])) hi = k;
        const int r = Qw[hi];
        if (sub[r] <= target || cptr[r + 1] == cptr[r]) break;
        Qw[hi] = Qw.back();
        Qw.pop_back();
        upper[r] = 1;
        for (int p = cptr[r]; p < cptr[r + 1]; ++p) Qw.push_back(child[p]);
        if (Qw.empty()) break;
      }
      // longest-processing-time assignment of the subtrees to the warps
      std::sort(Qw.begin(), Qw.end(), [&](int a, int b) {
        if (sub[a] != sub[b]) return sub[a] > sub[b];
        return a < b;
      });
      std::vector<int64_t> load(NWp, 0);
      for (int r : Qw) {
        if (upper[r]) continue;
        int best = 0;
        for (int w = 1; w < NWp; ++w)
          if (load[w] < load[best]) best = w;
        load[best] += sub[r];
        wroot[r] = best;
      }
    }
    for (int j = n - 1; j >= 0; --j) {
      if (owner[j] < 0 || upper[j]) wown[j] = -1;
      else if (wroot[j] >= 0) wown[j] = wroot[j];
      else wown[j] = wown[parent[j]];
    }
  }

  // ---- supernodes with inverted diagonal blocks ---------------------------------------------------
  // A chain a..b of the elimination tree with parent[j] = j+1 and nested column structures (a supernode)
  // has a dense lower-triangular diagonal block D; with Dinv = D^{-1} precomputed, its b-a+1 dependent
  // rows collapse into two parallel steps (off-block sums, then a triangular product), in both sweeps.
  // The same holds for any set J of nodes that is closed along tree paths: x_J = L_JJ^{-1} (b_J - L_{J,rest}
  // x_rest) only needs values of descendants outside J.  merge_depth > 1 groups every supernode at depth
  // d (in the tree of supernodes of one CTA / warp / the top) with its descendants down to depth d +
  // merge_depth - 1: the dependency levels of that part of the tree drop from 2 per supernode to 2 per
  // group, for an inverse that is only as dense as the factor's own coupling between those supernodes.
  // blk[i] = block index or -1; nodes of a block ascending.
  std::vector<int32_t> blk(n, -1), pos_in_blk(n, -1);
  struct Block { std::vector<int32_t> nodes; std::vector<double> Dinv; std::vector<char> pat; uint32_t off = 0; };  // Dinv row-major s x s (lower)
  std::vector<Block> blocks;
  std::vector<DmapRef> dmap_out;
  if (opt.min_block > 0 && n > 0) {
    auto ccount = [&](int j) { return L.colptr[j + 1] - L.colptr[j]; };
    // maximal chains
    std::vector<int32_t> chain_of(n, -1), chain_a, chain_b;
    for (int a0 = 0; a0 < n;) {
      int b0 = a0;
      while (b0 + 1 < n && parent[b0] == b0 + 1 && ccount(b0) == ccount(b0 + 1) + 1 && owner[b0] == owner[b0 + 1] &&
             wown[b0] == wown[b0 + 1] && (b0 + 1 - a0 + 1) <= opt.max_block)
        ++b0;
      for (int j = a0; j <= b0; ++j) chain_of[j] = (int32_t)chain_a.size();
      chain_a.push_back(a0);
      chain_b.push_back(b0);
      a0 = b0 + 1;
    }
    const int nch = (int)chain_a.size();
    // tree of chains (within one scope: same CTA, same warp or the top), depths, groups
    std::vector<int32_t> cpar(nch, -1), depth(nch, 0), groot(nch, -1);
    for (int c = nch - 1; c >= 0; --c) {
      const int pb = parent[chain_b[c]];
      if (pb >= 0 && owner[pb] == owner[chain_b[c]] && wown[pb] == wown[chain_b[c]]) cpar[c] = chain_of[pb];
      depth[c] = cpar[c] >= 0 ? depth[cpar[c]] + 1 : 0;
      const int md = std::max(1, opt.merge_depth);
      // chains shorter than min_block are leaves of the forest of supernodes in practice: they are not
      // worth a block of their own, but they may ride with a group
      groot[c] = (cpar[c] >= 0 && depth[c] % md != 0) ? groot[cpar[c]] : c;
    }
    // group sizes; oversized groups fall back to one block per chain
    std::vector<int64_t> gsize(nch, 0);
    for (int c = 0; c < nch; ++c) gsize[groot[c]] += chain_b[c] - chain_a[c] + 1;
    for (int c = 0; c < nch; ++c)
      if (gsize[groot[c]] > opt.max_block) groot[c] = -2 - c;  // marked: alone
    for (int c = 0; c < nch; ++c)
      if (groot[c] <= -2) groot[c] = c;
    std::fill(gsize.begin(), gsize.end(), 0);
    for (int c = 0; c < nch; ++c) gsize[groot[c]] += chain_b[c] - chain_a[c] + 1;
    // a group of several chains must be closed along tree paths: true by construction unless an oversized
    // group was dissolved above (members become their own roots, which keeps every remaining group closed)
    std::vector<int32_t> block_of_group(nch, -1);
    for (int c = 0; c < nch; ++c) {
      const int g = groot[c];
      if (gsize[g] < opt.min_block) continue;
      if (block_of_group[g] < 0) {
        block_of_group[g] = (int32_t)blocks.size();
        blocks.emplace_back();
      }
      Block& B = blocks[block_of_group[g]];
      for (int j = chain_a[c]; j <= chain_b[c]; ++j) {
        blk[j] = block_of_group[g];
        B.nodes.push_back(j);  // chains are visited in ascending order: nodes ascending
      }
    }
    uint32_t doff = 0;
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
      Block& B = blocks[bi];
      const int sz = (int)B.nodes.size();
      for (int k = 0; k < sz; ++k) pos_in_blk[B.nodes[k]] = k;
      B.off = doff;
      doff += (uint32_t)sz * (uint32_t)sz;
      std::vector<double> D((size_t)sz * sz, 0.0);
      std::vector<char> Dpat((size_t)sz * sz, 0);
      for (int k = 0; k < sz; ++k) {
        const int j = B.nodes[k];
        for (int p = L.colptr[j]; p < L.colptr[j + 1]; ++p) {
          const int i = L.rowval[p];
          if (blk[i] == blk[j]) {
            D[(size_t)pos_in_blk[i] * sz + k] = L.nzval[p];
            Dpat[(size_t)pos_in_blk[i] * sz + k] = 1;
            dmap_out.push_back(DmapRef{B.off + (uint32_t)(pos_in_blk[i] * sz + k), (uint32_t)p});
          }
        }
      }
      B.Dinv.assign((size_t)sz * sz, 0.0);
      invert_lower_block(sz, D.data(), B.Dinv.data());
      // STRUCTURAL pattern of the inverse (which entries the stream carries must not depend on the values:
      // a new realization rewrites values in place): entry (i, c) is present iff some k in [c, i) has D[i,k]
      // stored and Dinv[k,c] present
      B.pat.assign((size_t)sz * sz, 0);
      for (int c0 = 0; c0 < sz; ++c0) {
        B.pat[(size_t)c0 * sz + c0] = 1;
        for (int i = c0 + 1; i < sz; ++i) {
          char any = 0;
          for (int k = c0; k < i && !any; ++k) any = Dpat[(size_t)i * sz + k] && B.pat[(size_t)k * sz + c0];
          B.pat[(size_t)i * sz + c0] = any;
        }
      }
    }
  }

  // ---- dependency levels -----------------------------------------------------------------------
  // fe[i] / be[i]: level at which the (last) step producing x_i / v_i runs; a value is available
  // to steps of strictly higher levels.  For a block, step A (off-block sums) runs at level e-1 and
  // step B (Dinv product) at level e.  A block is handled when its last node is reached: in the forward
  // sweep all its outside dependencies (descendants) have smaller indices, in the backward sweep all of
  // them (ancestors of the group's root) larger ones.
  std::vector<int32_t> fe(n, 0), be(n, 0);
  auto same_scope = [&](int i, int j) { return (owner[i] >= 0) == (owner[j] >= 0); };
  std::vector<int32_t> single(1);
  for (int i = 0; i < n; ++i) {
    if (blk[i] >= 0 && i != blocks[blk[i]].nodes.back()) continue;  // handled with the block's last node
    single[0] = i;
    const std::vector<int32_t>& rows_of = blk[i] >= 0 ? blocks[blk[i]].nodes : single;
    int l = 0;
    for (int r : rows_of)
      for (int p = rptr[r]; p < rptr[r + 1]; ++p) {
        const int j = rcol[p];
        if (blk[i] >= 0 && blk[j] == blk[i]) continue;  // inside the block: absorbed by Dinv
        if (same_scope(r, j)) {
          if (owner[r] >= 0 && owner[j] != owner[r])
            throw std::runtime_error("internal: row dependency crosses subtrees");
          if (wown[r] >= 0 && wown[j] != wown[r])
            throw std::runtime_error("internal: private row depends on another warp's subtree");
          if (blk[j] >= 0 && blocks[blk[j]].nodes.back() > i)
            throw std::runtime_error("internal: block depends on a block that is not complete");
          // the private phase of the forward sweep is complete before the first upper level
          if (wown[r] < 0 && wown[j] >= 0) l = std::max(l, 0);
          else l = std::max(l, fe[j] + 1);
        } else if (owner[r] >= 0) {
          throw std::runtime_error("internal: subtree row depends on a top node");
        }
      }
    if (blk[i] >= 0) l += 1;  // step A at l, step B at l+1
    for (int r : rows_of) fe[r] = l;
  }
  for (int i = n - 1; i >= 0; --i) {
    if (blk[i] >= 0 && i != blocks[blk[i]].nodes.back()) continue;
    single[0] = i;
    const std::vector<int32_t>& rows_of = blk[i] >= 0 ? blocks[blk[i]].nodes : single;
    int l = 0;
    for (int r : rows_of)
      for (int p = L.colptr[r] + 1; p < L.colptr[r + 1]; ++p) {
        const int k = L.rowval[p];
        if (blk[i] >= 0 && blk[k] == blk[i]) continue;
        if (k < i) throw std::runtime_error("internal: block is not closed along tree paths");
        if (owner[r] < 0) {
          if (owner[k] >= 0) throw std::runtime_error("internal: ancestor of a top node is not top");
          l = std::max(l, be[k] + 1);
        } else if (owner[k] >= 0) {
          if (wown[r] < 0 && wown[k] >= 0) throw std::runtime_error("internal: ancestor of an upper node is private");
          if (wown[r] >= 0 && wown[k] >= 0 && wown[k] != wown[r])
            throw std::runtime_error("internal: private row depends on another warp's subtree");
          // the upper levels of the backward sweep are complete before the private phase
          if (wown[r] >= 0 && wown[k] < 0) l = std::max(l, 0);
          else l = std::max(l, be[k] + 1);
        }
      }
    if (blk[i] >= 0) l += 1;
    for (int r : rows_of) be[r] = l;
  }

  // ---- positions in the shared vectors ---------------------------------------------------------
  // every node has a slot; nodes of inverted blocks have a second (scratch) slot.
  std::vector<int32_t> n_own(C, 0), own_sz(C, 0);
  std::vector<int32_t> pos(n, -1), scr(n, -1);
  int n_top = 0, top_sz = 0;
  {
    std::vector<std::vector<int32_t>> own_nodes(C);
    for (int i = 0; i < n; ++i) {
      if (owner[i] < 0) pos[i] = n_top++;
      else own_nodes[owner[i]].push_back(i);
    }
    top_sz = n_top;
    for (int i = 0; i < n; ++i)
      if (owner[i] < 0 && blk[i] >= 0) scr[i] = top_sz++;
    for (int c = 0; c < C; ++c) {
      auto& v = own_nodes[c];
      std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return fe[a] < fe[b]; });
      for (size_t k = 0; k < v.size(); ++k) pos[v[k]] = (int)k;
      n_own[c] = (int)v.size();
      int q = n_own[c];
      for (int node : v)
        if (blk[node] >= 0) scr[node] = q++;
      own_sz[c] = q;
    }
  }
  const int part_off = 0, top_off = n_top, out_off = n_top + top_sz, xd_off = out_off + ng,
            own_off = xd_off + ng;
  int max_own = *std::max_element(own_sz.begin(), own_sz.end());
  if (own_off + max_own > opt.max_smem_doubles || own_off + max_own > 65535)
    throw std::runtime_error("subdomain does not fit the shared-memory vector: use a larger cluster size");

  // slot of b_i / v_i, scratch slot, and where the forward result x_i lives
  auto sidx = [&](int node) -> uint16_t {
    return (uint16_t)(owner[node] < 0 ? top_off + pos[node] : own_off + pos[node]);
  };
  auto scridx = [&](int node) -> uint16_t {
    return (uint16_t)(owner[node] < 0 ? top_off + scr[node] : own_off + scr[node]);
  };
  auto xidx = [&](int node) -> uint16_t { return blk[node] >= 0 ? scridx(node) : sidx(node); };

  // ---- A_IG by permuted row, A_GG by row -------------------------------------------------------
  std::vector<int32_t> pinv(n);
  for (int k = 0; k < n; ++k) {
    if (in.perm[k] < 0 || in.perm[k] >= n) throw std::runtime_error("perm out of range");
    pinv[in.perm[k]] = k;
  }
  const Csc& G = in.AIG;
  if (G.nrow != n || G.ncol != ng) throw std::runtime_error("A_IG has wrong dimensions");
  std::vector<int32_t> gptr(n + 1, 0);
  for (int32_t v : G.rowval) {
    if (v < 0 || v >= n) throw std::runtime_error("A_IG row index out of range");
    gptr[pinv[v] + 1]++;
  }
  for (int i = 0; i < n; ++i) gptr[i + 1] += gptr[i];
  std::vector<int32_t> gcol(G.rowval.size()), gsrc(G.rowval.size());
  std::vector<double> gval(G.rowval.size());
  {
    std::vector<int32_t> fill(gptr.begin(), gptr.end() - 1);
    for (int l = 0; l < ng; ++l)
      for (int p = G.colptr[l]; p < G.colptr[l + 1]; ++p) {
        int i = pinv[G.rowval[p]];
        gcol[fill[i]] = l;
        gval[fill[i]] = G.nzval[p];
        gsrc[fill[i]] = p;
        fill[i]++;
      }
  }
  const Csc& GG = in.AGG;
  if (GG.nrow != ng || GG.ncol != ng) throw std::runtime_error("A_GG has wrong dimensions");
  std::vector<int32_t> ggptr(ng + 1, 0);
  for (int32_t v : GG.rowval) ggptr[v + 1]++;
  for (int i = 0; i < ng; ++i) ggptr[i + 1] += ggptr[i];
  std::vector<int32_t> ggcol(GG.rowval.size()), ggsrc(GG.rowval.size());
  std::vector<double> ggval(GG.rowval.size());
  {
    std::vector<int32_t> fill(ggptr.begin(), ggptr.end() - 1);
    for (int m = 0; m < ng; ++m)
      for (int p = GG.colptr[m]; p < GG.colptr[m + 1]; ++p) {
        int l = GG.rowval[p];
        ggcol[fill[l]] = m;
        ggval[fill[l]] = GG.nzval[p];
        ggsrc[fill[l]] = p;
        fill[l]++;
      }
  }

  // ---- emit ------------------------------------------------------------------------------------
  out = DomainPlan();
  out.nnz_L = nnzl;
  out.n_top = n_top;
  out.ctas.resize(C);
  Scratch sc;
  std::vector<TmpRow> rows;
  int cur_warp = -1;  // warp of the rows being emitted (-1: CTA-wide levels)
  auto begin_row = [&](uint16_t dst, double dg, int level, int kind, uint32_t dg_src = kNoSrc) {
    TmpRow r;
    r.diag_src = dg_src;
    r.warp = cur_warp;
    r.dst = dst;
    r.diag = dg;
    r.level = level;
    r.kind = kind;
    r.bcast = 0;
    r.ebeg = (uint32_t)sc.vals.size();
    r.len = 0;
    rows.push_back(r);
  };
  // pool layout of the value refresh (plan.h): [L | A_IG | A_GG | Dinv blocks]
  const uint32_t poolL = 0, poolG = (uint32_t)L.nzval.size(), poolGG = poolG + (uint32_t)G.nzval.size(),
                 poolD = poolGG + (uint32_t)GG.nzval.size();
  auto push = [&](uint16_t col, double v, uint32_t src) {
    sc.cols.push_back(col);
    sc.vals.push_back(v);
    sc.srcs.push_back(src);
    rows.back().len++;
  };

  for (int c = 0; c < C; ++c) {
    CtaStream cs;
    CtaDesc& D = out.ctas[c];
    std::memset(&D, 0, sizeof(D));
    D.rank = c;
    D.csize = C;
    D.n_own = n_own[c];
    D.n_top = n_top;
    D.n_gd = ng;
    D.part_off = part_off;
    D.top_off = top_off;
    D.out_off = out_off;
    D.xd_off = xd_off;
    D.own_off = own_off;
    D.smem_doubles = own_off + max_own;
    D.top_sz = top_sz;
    auto section = [&](int s, int force_levels = -1, bool private_first = true) {
      emit_section(rows, sc, opt, cs, D.sec_chunk0[s], D.sec_chunk1[s], D.sec_nlevels[s], force_levels, private_first,
                   &out.n_private_rows, &out.n_private_slices);
      rows.clear();
      sc.cols.clear();
      sc.vals.clear();
      sc.srcs.clear();
    };
    auto mine = [&](int node) { return owner[node] == c || (c == 0 && owner[node] < 0); };

    // SEC_INIT
    for (int i = 0; i < n; ++i)
      if (mine(i) && gptr[i + 1] > gptr[i]) {
        begin_row(sidx(i), 1.0, 0, 0);
        for (int p = gptr[i]; p < gptr[i + 1]; ++p) push((uint16_t)(xd_off + gcol[p]), gval[p], poolG + (uint32_t)gsrc[p]);
      }
    section(SEC_INIT);
    // forward / backward rows of the nodes selected by `sel` (own subtree nodes, or top nodes)
    auto emit_forward = [&](auto sel) {
      for (int i = 0; i < n; ++i) {
        if (!sel(i)) continue;
        cur_warp = wown[i];
        if (blk[i] < 0) {
          begin_row(sidx(i), diag[i], fe[i], 0, poolL + (uint32_t)L.colptr[i]);
          for (int p = rptr[i]; p < rptr[i + 1]; ++p)
            if (same_scope(i, rcol[p])) push(xidx(rcol[p]), rval[p], poolL + (uint32_t)rsrc[p]);
        } else {
          const Block& B = blocks[blk[i]];
          const int sz = (int)B.nodes.size();
          const int pi = pos_in_blk[i];
          // step A: t_i = b_i - sum_{j outside the block} L_ij x_j   (in place in slot i)
          bool any = false;
          for (int p = rptr[i]; p < rptr[i + 1]; ++p) {
            const int j = rcol[p];
            if (blk[j] == blk[i] || !same_scope(i, j)) continue;
            if (!any) { begin_row(sidx(i), 1.0, fe[i] - 1, 0); any = true; }
            push(xidx(j), rval[p], poolL + (uint32_t)rsrc[p]);
          }
          // step B: x_i = sum_{k in [a, i]} Dinv[i,k] t_k   (fresh write to the scratch slot)
          begin_row(scridx(i), -1.0, fe[i], 1);
          for (int pk = 0; pk <= pi; ++pk) {
            const size_t e = (size_t)pi * sz + pk;
            if (B.pat[e]) push(sidx(B.nodes[pk]), B.Dinv[e], poolD + B.off + (uint32_t)e);
          }
        }
      }
    };
    auto emit_backward = [&](auto sel) {
      for (int i = 0; i < n; ++i) {
        if (!sel(i)) continue;
        cur_warp = wown[i];
        if (blk[i] < 0) {
          begin_row(sidx(i), diag[i], be[i], 0, poolL + (uint32_t)L.colptr[i]);
          for (int p = L.colptr[i] + 1; p < L.colptr[i + 1]; ++p) push(sidx(L.rowval[p]), L.nzval[p], poolL + (uint32_t)p);
        } else {
          const Block& B = blocks[blk[i]];
          const int sz = (int)B.nodes.size();
          const int pi = pos_in_blk[i];
          // step A': t_i = y_i - sum_{k outside the block} L_ki v_k   (y_i lives in the scratch slot)
          bool any = false;
          for (int p = L.colptr[i] + 1; p < L.colptr[i + 1]; ++p) {
            const int k = L.rowval[p];
            if (blk[k] == blk[i]) continue;
            if (!any) { begin_row(scridx(i), 1.0, be[i] - 1, 0); any = true; }
            push(sidx(k), L.nzval[p], poolL + (uint32_t)p);
          }
          // step B': v_i = sum_{k in [i, b]} Dinv[k,i] t_k   (fresh write to slot i)
          begin_row(sidx(i), -1.0, be[i], 1);
          for (int pk = pi; pk < sz; ++pk) {
            const size_t e = (size_t)pk * sz + pi;
            if (B.pat[e]) push(scridx(B.nodes[pk]), B.Dinv[e], poolD + B.off + (uint32_t)e);
          }
        }
      }
    };
    // SEC_FWD
    emit_forward([&](int i) { return owner[i] == c; });
    cur_warp = -1;
    section(SEC_FWD, -1, true);
    // SEC_PART
    for (int i = 0; i < n; ++i)
      if (owner[i] < 0) {
        bool any = false;
        for (int p = rptr[i]; p < rptr[i + 1]; ++p)
          if (owner[rcol[p]] == c) {
            if (!any) {
              begin_row((uint16_t)(part_off + pos[i]), 1.0, 0, 0);
              any = true;
            }
            push(xidx(rcol[p]), rval[p], poolL + (uint32_t)rsrc[p]);
          }
      }
    section(SEC_PART);
    // SEC_TOPF / SEC_TOPB: every CTA builds the same row list and keeps its share of each level
    // (greedy balance on entries, identical on all ranks); results are broadcast (bcast rows).
    auto keep_share = [&]() -> int {
      int maxlev = -1;
      for (auto& r : rows) maxlev = std::max(maxlev, r.level);
      std::vector<uint32_t> ord(rows.size());
      std::iota(ord.begin(), ord.end(), 0u);
      std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
        if (rows[a].level != rows[b].level) return rows[a].level < rows[b].level;
        return rows[a].len > rows[b].len;
      });
      std::vector<char> mine_row(rows.size(), 0);
      std::vector<uint64_t> load(C, 0);
      int cur = -1;
      for (uint32_t oi : ord) {
        if (rows[oi].level != cur) {
          cur = rows[oi].level;
          std::fill(load.begin(), load.end(), 0);
        }
        int best = 0;
        for (int k = 1; k < C; ++k)
          if (load[k] < load[best]) best = k;
        load[best] += rows[oi].len + 4;
        mine_row[oi] = best == c;
      }
      std::vector<TmpRow> kept;
      for (size_t k = 0; k < rows.size(); ++k)
        if (mine_row[k]) {
          rows[k].bcast = 1;
          kept.push_back(rows[k]);
        }
      rows.swap(kept);
      return maxlev + 1;
    };
    int top_levels = 0;
    if (C > 1) {
      emit_forward([&](int i) { return owner[i] < 0; });
      top_levels = keep_share();
    }
    section(SEC_TOPF, top_levels);
    if (C > 1) {
      emit_backward([&](int i) { return owner[i] < 0; });
      top_levels = keep_share();
    }
    section(SEC_TOPB, top_levels);
    // SEC_BWD
    emit_backward([&](int i) { return owner[i] == c; });
    cur_warp = -1;
    section(SEC_BWD, -1, false);
    // SEC_GG (rank 0)
    if (c == 0)
      for (int l = 0; l < ng; ++l)
        if (ggptr[l + 1] > ggptr[l]) {
          begin_row((uint16_t)(out_off + l), 1.0, 0, 0);
          for (int p = ggptr[l]; p < ggptr[l + 1]; ++p) push((uint16_t)(xd_off + ggcol[p]), ggval[p], poolGG + (uint32_t)ggsrc[p]);
        }
    section(SEC_GG);
    // SEC_OUT
    for (int l = 0; l < ng; ++l) {
      bool any = false;
      for (int p = G.colptr[l]; p < G.colptr[l + 1]; ++p) {
        int i = pinv[G.rowval[p]];
        if (mine(i)) {
          if (!any) {
            begin_row((uint16_t)(out_off + l), 1.0, 0, 0);
            any = true;
          }
          push(sidx(i), G.nzval[p], poolG + (uint32_t)p);
        }
      }
    }
    section(SEC_OUT);

    // append the CTA stream to the plan's stream pool
    D.stream_off = out.stream.size();
    {
      const uint64_t base8 = (uint64_t)out.stream.size() / 8;
      for (const ValRef& r : cs.refs) {
        if (base8 + r.off8 > 0xffffffffull) throw std::runtime_error("internal: stream too long for 32-bit value slots");
        out.refs.push_back(ValRef{(uint32_t)(base8 + r.off8), r.src});
      }
    }
    out.stream.insert(out.stream.end(), cs.bytes.begin(), cs.bytes.end());
    if (cs.chunk_units.size() * (size_t)kChunkBytes != cs.bytes.size()) throw std::runtime_error("internal: chunk table size");
    out.chunk_units.insert(out.chunk_units.end(), cs.chunk_units.begin(), cs.chunk_units.end());
    out.n_entries += cs.n_entries;
    out.n_rows += cs.n_rows;
    out.n_packets += cs.n_packets;
    out.lane_steps += cs.lane_steps;
    out.max_cta_entries = std::max(out.max_cta_entries, cs.n_entries);
    for (int s = 0; s < NSEC; ++s) out.max_levels = std::max(out.max_levels, (int)D.sec_nlevels[s]);
    // position -> original interior index maps
    D.ownmap_off = (int64_t)out.imap.size();
    {
      std::vector<int32_t> m(n_own[c]);
      for (int i = 0; i < n; ++i)
        if (owner[i] == c) m[pos[i]] = in.perm[i];
      out.imap.insert(out.imap.end(), m.begin(), m.end());
    }
    D.topmap_off = (int64_t)out.imap.size();
    {
      std::vector<int32_t> m(n_top);
      for (int i = 0; i < n; ++i)
        if (owner[i] < 0) m[pos[i]] = in.perm[i];
      out.imap.insert(out.imap.end(), m.begin(), m.end());
    }
  }
  out.dmap = dmap_out;
  out.pool_L = (uint32_t)L.nzval.size();
  out.pool_G = (uint32_t)G.nzval.size();
  out.pool_GG = (uint32_t)GG.nzval.size();
  out.pool_D = 0;
  for (auto& B : blocks) {
    out.blk_off.push_back(B.off);
    out.blk_sz.push_back((uint32_t)B.nodes.size());
    out.pool_D += (uint32_t)(B.nodes.size() * B.nodes.size());
  }
}

void invert_lower_block(int sz, const double* D, double* Dinv) {
  for (int c0 = 0; c0 < sz; ++c0) {
    Dinv[(size_t)c0 * sz + c0] = 1.0 / D[(size_t)c0 * sz + c0];
    for (int i = c0 + 1; i < sz; ++i) {
      double acc = 0.0;
      bool any = false;
      for (int k = c0; k < i; ++k) {
        const double d = D[(size_t)i * sz + k], v = Dinv[(size_t)k * sz + c0];
        if (d != 0.0 && v != 0.0) {
          acc += d * v;
          any = true;
        }
      }
      Dinv[(size_t)i * sz + c0] = any ? -acc / D[(size_t)i * sz + i] : 0.0;
    }
  }
}

}  // namespace kb

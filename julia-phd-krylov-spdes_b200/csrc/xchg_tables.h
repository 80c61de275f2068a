// Send / gather tables of the owner-ordered compact interface exchange (see kernels.cuh, "OWNER-ORDERED
// COMPACT EXCHANGE").  Pure host function: every rank derives the layout of ALL ranks from the topology
// (deterministic, no communication).  Entries of rank r sit at [off_r, off_r + m_r) of the gathered
// array, ordered by interface node, a node's raw contributions by ascending subdomain.
//
// Reference order of the interface sums: EllipticPdeDomainDecomposition.jl:740-742 (ascending subdomain).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace kb {

struct XchTables {
  // this rank's send entries: entry e = (base[snd_g[e]] if snd_g[e] >= 0) + sum loc[snd_src[snd_ptr[e] .. snd_ptr[e+1])]
  std::vector<int32_t> snd_ptr, snd_src, snd_g;
  // every interface node g: y[g] = sum of gathered[fin_src[fin_ptr[g] .. fin_ptr[g+1])] in that order
  std::vector<int32_t> fin_ptr, fin_src;
  std::vector<int32_t> off, flag_off;  // per rank (+1): entry / flag offsets
  int m_total = 0;
};

// l2g[d]: 0-based interface ids of the local interface nodes of subdomain d (all subdomains).
// Subdomain d lives on rank r iff (ndom*r)/world <= d < (ndom*(r+1))/world; `loc` of a rank is the
// concatenation of its subdomains' local interface vectors.  `threads`: entries per send block.
inline XchTables build_xchg_tables(int world, int rank, const std::vector<const std::vector<int32_t>*>& l2g, int64_t ng,
                                   int threads) {
  XchTables T;
  const int64_t ndom = (int64_t)l2g.size();
  if (world < 1 || rank < 0 || rank >= world || threads < 1) throw std::runtime_error("exchange tables: bad arguments");
  std::vector<int32_t> dom_rank(ndom);
  for (int r = 0; r < world; ++r)
    for (int64_t d = (ndom * r) / world; d < (ndom * (r + 1)) / world; ++d) dom_rank[d] = r;
  const int64_t d0 = (ndom * rank) / world, d1 = (ndom * (rank + 1)) / world;
  std::vector<int64_t> gbase(d1 - d0 + 1, 0);
  for (int64_t d = d0; d < d1; ++d) gbase[d - d0 + 1] = gbase[d - d0] + (int64_t)l2g[d]->size();
  // global adjacency: contributions to each interface node in ascending subdomain order
  std::vector<int32_t> gp(ng + 1, 0);
  for (int64_t d = 0; d < ndom; ++d)
    for (int32_t g : *l2g[d]) gp[g + 1]++;
  for (int64_t g = 0; g < ng; ++g) gp[g + 1] += gp[g];
  std::vector<int32_t> gd(gp[ng]), gl(gp[ng]);
  {
    std::vector<int32_t> fill(gp.begin(), gp.end() - 1);
    for (int64_t d = 0; d < ndom; ++d)
      for (size_t l = 0; l < l2g[d]->size(); ++l) {
        const int32_t q = fill[(*l2g[d])[l]]++;
        gd[q] = (int32_t)d;
        gl[q] = (int32_t)l;
      }
  }
  auto local_src = [&](int32_t d, int32_t l) { return (int32_t)(gbase[d - d0] + l); };
  std::vector<int32_t> m(world, 0), fin_rank, fin_pos;
  T.fin_ptr.assign(ng + 1, 0);
  T.snd_ptr.assign(1, 0);
  for (int64_t g = 0; g < ng; ++g) {
    int k = gp[g];
    const int kend = gp[g + 1];
    if (k < kend) {
      // the lowest rank: ONE entry, its contributions summed in ascending order (starting from `base`)
      const int r0 = dom_rank[gd[k]];
      fin_rank.push_back(r0);
      fin_pos.push_back(m[r0]++);
      if (r0 == rank) {
        for (int q = k; q < kend && dom_rank[gd[q]] == r0; ++q) T.snd_src.push_back(local_src(gd[q], gl[q]));
        T.snd_ptr.push_back((int32_t)T.snd_src.size());
        T.snd_g.push_back((int32_t)g);
      }
      while (k < kend && dom_rank[gd[k]] == r0) ++k;
      // higher ranks: one entry per contribution
      for (; k < kend; ++k) {
        const int r = dom_rank[gd[k]];
        fin_rank.push_back(r);
        fin_pos.push_back(m[r]++);
        if (r == rank) {
          T.snd_src.push_back(local_src(gd[k], gl[k]));
          T.snd_ptr.push_back((int32_t)T.snd_src.size());
          T.snd_g.push_back(-1);
        }
      }
    }
    T.fin_ptr[g + 1] = (int32_t)fin_rank.size();
  }
  T.off.assign(world + 1, 0);
  T.flag_off.assign(world + 1, 0);
  for (int r = 0; r < world; ++r) {
    T.off[r + 1] = T.off[r] + m[r];
    T.flag_off[r + 1] = T.flag_off[r] + (m[r] + threads - 1) / threads;
  }
  T.fin_src.resize(fin_rank.size());
  for (size_t i = 0; i < fin_rank.size(); ++i) T.fin_src[i] = T.off[fin_rank[i]] + fin_pos[i];
  T.m_total = T.off[world];
  if ((int)T.snd_g.size() != m[rank]) throw std::runtime_error("exchange tables: entry count mismatch");
  return T;
}

}  // namespace kb

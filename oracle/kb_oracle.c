/* CPU restatement in C of the per-iteration hot path (TEST INFRASTRUCTURE / CPU BASELINE ONLY).
 *
 * Follows the reference line by line, with the interior solves done exactly by the uploaded
 * Cholesky factor (the reference configuration `preconds = exact factor`):
 *   ko_apply_schur  apply_local_schurs + apply_local_schur
 *                   (Fem/EllipticPdeDomainDecomposition.jl:711-747, :639-654)
 *   ko_apply_nn     apply_neumann_neumann_schur (:1361-1386)
 *   ko_pcg          pcg (RecyclingKrylovSolvers/cg.jl:67-109)
 * nthreads = 1 reproduces the reference's sequential `for idom in 1:ndom` loop (:723, :1367);
 * nthreads > 1 runs the subdomains under OpenMP and sums their contributions in subdomain order,
 * so both give bitwise-identical results.  PARITY UNPINNED (no reference golden vectors exist).
 *
 * Build: gcc -O3 -march=native -fopenmp -fPIC -shared kb_oracle.c -o libkb_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t nI, nG;
  int32_t *perm;                      /* perm[new] = old */
  int32_t *Lp, *Li; double *Lx;       /* CSC lower, diagonal first */
  int32_t *Gp, *Gi; double *Gx;       /* A_IG  nI x nG CSC */
  int32_t *Hp, *Hi; double *Hx;       /* A_GG  nG x nG CSC */
  int32_t *l2g;                       /* local interface -> interface id */
  double *PiS;                        /* nG x nG column-major, or NULL */
  double *work;                       /* nI + 3 nG doubles */
} ko_dom;

typedef struct {
  int32_t ndom, n_gamma;
  double *inv_cnt_unused;
  int32_t *cnt;
  ko_dom *dom;
  double *loc; int64_t *gbase;        /* per-subdomain local results */
  int reverse;                        /* 1: interface sums in DESCENDING subdomain order (rounding-floor probe) */
} ko_model;

static void *dup_mem(const void *src, size_t bytes) {
  void *p = malloc(bytes ? bytes : 1);
  if (bytes) memcpy(p, src, bytes);
  return p;
}

ko_model *ko_new(int32_t ndom, int32_t n_gamma, const int32_t *cnt) {
  ko_model *m = (ko_model *)calloc(1, sizeof(ko_model));
  m->ndom = ndom; m->n_gamma = n_gamma;
  m->cnt = (int32_t *)dup_mem(cnt, sizeof(int32_t) * (size_t)n_gamma);
  m->dom = (ko_dom *)calloc((size_t)ndom, sizeof(ko_dom));
  m->gbase = (int64_t *)calloc((size_t)ndom + 1, sizeof(int64_t));
  return m;
}

void ko_set_dom(ko_model *m, int32_t d, int32_t nI, int32_t nG, const int32_t *perm, const int32_t *Lp,
                const int32_t *Li, const double *Lx, const int32_t *Gp, const int32_t *Gi, const double *Gx,
                const int32_t *Hp, const int32_t *Hi, const double *Hx, const int32_t *l2g, const double *PiS) {
  ko_dom *D = &m->dom[d];
  D->nI = nI; D->nG = nG;
  D->perm = (int32_t *)dup_mem(perm, 4 * (size_t)nI);
  D->Lp = (int32_t *)dup_mem(Lp, 4 * ((size_t)nI + 1));
  D->Li = (int32_t *)dup_mem(Li, 4 * (size_t)Lp[nI]);
  D->Lx = (double *)dup_mem(Lx, 8 * (size_t)Lp[nI]);
  D->Gp = (int32_t *)dup_mem(Gp, 4 * ((size_t)nG + 1));
  D->Gi = (int32_t *)dup_mem(Gi, 4 * (size_t)Gp[nG]);
  D->Gx = (double *)dup_mem(Gx, 8 * (size_t)Gp[nG]);
  D->Hp = (int32_t *)dup_mem(Hp, 4 * ((size_t)nG + 1));
  D->Hi = (int32_t *)dup_mem(Hi, 4 * (size_t)Hp[nG]);
  D->Hx = (double *)dup_mem(Hx, 8 * (size_t)Hp[nG]);
  D->l2g = (int32_t *)dup_mem(l2g, 4 * (size_t)nG);
  D->PiS = PiS ? (double *)dup_mem(PiS, 8 * (size_t)nG * (size_t)nG) : NULL;
  D->work = (double *)malloc(8 * ((size_t)2 * nI + 3 * (size_t)nG + 8));
}

void ko_finish(ko_model *m) {
  for (int32_t d = 0; d < m->ndom; ++d) m->gbase[d + 1] = m->gbase[d] + m->dom[d].nG;
  m->loc = (double *)malloc(8 * (size_t)(m->gbase[m->ndom] + 1));
}

void ko_free(ko_model *m) {
  if (!m) return;
  for (int32_t d = 0; d < m->ndom; ++d) {
    ko_dom *D = &m->dom[d];
    free(D->perm); free(D->Lp); free(D->Li); free(D->Lx); free(D->Gp); free(D->Gi); free(D->Gx);
    free(D->Hp); free(D->Hi); free(D->Hx); free(D->l2g); free(D->PiS); free(D->work);
  }
  free(m->dom); free(m->cnt); free(m->gbase); free(m->loc); free(m);
}

/* v = A_II^{-1} t, in the permuted ordering, in place: w = L^-T L^-1 w */
static void chol_solve(const ko_dom *D, double *w) {
  const int32_t n = D->nI;
  for (int32_t j = 0; j < n; ++j) {
    const int32_t p0 = D->Lp[j], p1 = D->Lp[j + 1];
    const double xj = w[j] / D->Lx[p0];
    w[j] = xj;
    for (int32_t p = p0 + 1; p < p1; ++p) w[D->Li[p]] -= D->Lx[p] * xj;
  }
  for (int32_t j = n - 1; j >= 0; --j) {
    const int32_t p0 = D->Lp[j], p1 = D->Lp[j + 1];
    double s = w[j];
    for (int32_t p = p0 + 1; p < p1; ++p) s -= D->Lx[p] * w[D->Li[p]];
    w[j] = s / D->Lx[p0];
  }
}

/* S_d x_d  (EllipticPdeDomainDecomposition.jl:646-652) */
static void local_schur(const ko_dom *D, const double *xd, double *out) {
  const int32_t nI = D->nI, nG = D->nG;
  double *t = D->work, *w = D->work + nI;
  memset(t, 0, 8 * (size_t)nI);
  for (int32_t l = 0; l < nG; ++l) {            /* t = A_IG xd */
    const double xl = xd[l];
    for (int32_t p = D->Gp[l]; p < D->Gp[l + 1]; ++p) t[D->Gi[p]] += D->Gx[p] * xl;
  }
  for (int32_t k = 0; k < nI; ++k) w[k] = t[D->perm[k]];
  chol_solve(D, w);
  for (int32_t k = 0; k < nI; ++k) t[D->perm[k]] = w[k];
  memset(out, 0, 8 * (size_t)nG);
  for (int32_t m = 0; m < nG; ++m) {            /* out = A_GG xd */
    const double xm = xd[m];
    for (int32_t p = D->Hp[m]; p < D->Hp[m + 1]; ++p) out[D->Hi[p]] += D->Hx[p] * xm;
  }
  for (int32_t l = 0; l < nG; ++l) {            /* out -= A_IG' v */
    double s = 0.0;
    for (int32_t p = D->Gp[l]; p < D->Gp[l + 1]; ++p) s += D->Gx[p] * t[D->Gi[p]];
    out[l] -= s;
  }
}

void ko_apply_schur(ko_model *m, const double *x, double *Sx, int nthreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int32_t d = 0; d < m->ndom; ++d) {
    const ko_dom *D = &m->dom[d];
    double *xd = D->work + 2 * (size_t)D->nI;
    for (int32_t l = 0; l < D->nG; ++l) xd[l] = x[D->l2g[l]];     /* :728-730 */
    local_schur(D, xd, m->loc + m->gbase[d]);
  }
  memset(Sx, 0, 8 * (size_t)m->n_gamma);
  for (int32_t dd = 0; dd < m->ndom; ++dd) {                       /* :740-742, ascending d */
    const int32_t d = m->reverse ? m->ndom - 1 - dd : dd;
    const ko_dom *D = &m->dom[d];
    const double *o = m->loc + m->gbase[d];
    for (int32_t l = 0; l < D->nG; ++l) Sx[D->l2g[l]] += o[l];
  }
}

void ko_apply_nn(ko_model *m, const double *r, double *z, int nthreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int32_t d = 0; d < m->ndom; ++d) {
    const ko_dom *D = &m->dom[d];
    const int32_t nG = D->nG;
    double *rd = D->work + 2 * (size_t)D->nI, *o = m->loc + m->gbase[d];
    for (int32_t l = 0; l < nG; ++l) rd[l] = r[D->l2g[l]] / (double)m->cnt[D->l2g[l]];   /* :1373-1375 */
    memset(o, 0, 8 * (size_t)nG);
    for (int32_t j = 0; j < nG; ++j) {                                                  /* :1377 */
      const double rj = rd[j];
      const double *col = D->PiS + (size_t)j * nG;
      for (int32_t i = 0; i < nG; ++i) o[i] += col[i] * rj;
    }
  }
  memset(z, 0, 8 * (size_t)m->n_gamma);
  for (int32_t dd = 0; dd < m->ndom; ++dd) {                                             /* :1379-1381 */
    const int32_t d = m->reverse ? m->ndom - 1 - dd : dd;
    const ko_dom *D = &m->dom[d];
    const double *o = m->loc + m->gbase[d];
    for (int32_t l = 0; l < D->nG; ++l) z[D->l2g[l]] += o[l] / (double)m->cnt[D->l2g[l]];
  }
}

/* Rounding-floor probe: the same algorithm with the interface sums taken in descending subdomain order --
 * mathematically identical, different rounding.  The distance between the two CPU runs is the floor below
 * which no implementation of this arithmetic can be compared with another. */
void ko_set_reverse(ko_model *m, int flag) { m->reverse = flag; }

/* PiS of subdomain d (n_gd x n_gd column-major) set after construction: the constant "median"
 * preconditioner of Example07_PcgSchurStochasticEllipticPde.jl:152-154 comes from another realization. */
void ko_set_pis(ko_model *m, int32_t d, const double *PiS) {
  ko_dom *D = &m->dom[d];
  free(D->PiS);
  D->PiS = (double *)dup_mem(PiS, 8 * (size_t)D->nG * (size_t)D->nG);
}

/* NB right-hand sides at once (interleaved w[k*NB + c]): the factor is streamed once per NB columns */
#define KO_NB 8
static void chol_solve_nb(const ko_dom *D, double *w) {
  const int32_t n = D->nI;
  for (int32_t j = 0; j < n; ++j) {
    const int32_t p0 = D->Lp[j], p1 = D->Lp[j + 1];
    double xj[KO_NB];
    const double dg = D->Lx[p0];
    for (int c = 0; c < KO_NB; ++c) { xj[c] = w[(size_t)j * KO_NB + c] / dg; w[(size_t)j * KO_NB + c] = xj[c]; }
    for (int32_t p = p0 + 1; p < p1; ++p) {
      const double l = D->Lx[p];
      double *wi = w + (size_t)D->Li[p] * KO_NB;
      for (int c = 0; c < KO_NB; ++c) wi[c] -= l * xj[c];
    }
  }
  for (int32_t j = n - 1; j >= 0; --j) {
    const int32_t p0 = D->Lp[j], p1 = D->Lp[j + 1];
    double sacc[KO_NB];
    for (int c = 0; c < KO_NB; ++c) sacc[c] = w[(size_t)j * KO_NB + c];
    for (int32_t p = p0 + 1; p < p1; ++p) {
      const double l = D->Lx[p];
      const double *wi = w + (size_t)D->Li[p] * KO_NB;
      for (int c = 0; c < KO_NB; ++c) sacc[c] -= l * wi[c];
    }
    const double dg = D->Lx[p0];
    for (int c = 0; c < KO_NB; ++c) w[(size_t)j * KO_NB + c] = sacc[c] / dg;
  }
}

/* assemble_local_schurs (EllipticPdeDomainDecomposition.jl:667-695): dense S_d column by column
 * (S_d e_j by apply_local_schur, here KO_NB unit vectors per pass over the factor), then Symmetric():
 * the upper triangle mirrored (:692).  Sd: n_gd x n_gd column-major, subdomain d.  Thread-safe across
 * different subdomains (own scratch). */
void ko_assemble_local_schur(const ko_model *m, int32_t d, double *Sd) {
  const ko_dom *D = &m->dom[d];
  const int32_t nI = D->nI, nG = D->nG;
  double *w = (double *)malloc(8 * ((size_t)nI + 1) * KO_NB);
  int32_t *pinv = (int32_t *)malloc(4 * ((size_t)nI + 1));
  for (int32_t k = 0; k < nI; ++k) pinv[D->perm[k]] = k;
  for (int32_t j0 = 0; j0 < nG; j0 += KO_NB) {
    const int nb = (nG - j0 < KO_NB) ? nG - j0 : KO_NB;
    memset(w, 0, 8 * (size_t)nI * KO_NB);
    for (int c = 0; c < nb; ++c)                       /* t = A_IG e_j, rows in the factor's ordering */
      for (int32_t p = D->Gp[j0 + c]; p < D->Gp[j0 + c + 1]; ++p) w[(size_t)pinv[D->Gi[p]] * KO_NB + c] += D->Gx[p];
    chol_solve_nb(D, w);                               /* v = A_II^{-1} t */
    for (int c = 0; c < nb; ++c) {
      double *col = Sd + (size_t)(j0 + c) * nG;
      memset(col, 0, 8 * (size_t)nG);
      for (int32_t p = D->Hp[j0 + c]; p < D->Hp[j0 + c + 1]; ++p) col[D->Hi[p]] += D->Hx[p];   /* A_GG e_j */
      for (int32_t l = 0; l < nG; ++l) {               /* -= A_IG' v */
        double s = 0.0;
        for (int32_t p = D->Gp[l]; p < D->Gp[l + 1]; ++p) s += D->Gx[p] * w[(size_t)pinv[D->Gi[p]] * KO_NB + c];
        col[l] -= s;
      }
    }
  }
  for (int32_t j = 0; j < nG; ++j)                     /* Symmetric(): upper triangle wins */
    for (int32_t i = j + 1; i < nG; ++i) Sd[i + (size_t)j * nG] = Sd[j + (size_t)i * nG];
  free(w);
  free(pinv);
}

/* All subdomains under OpenMP: Sd_all = concatenation of the column-major blocks, off[d] their offsets. */
void ko_assemble_local_schurs(const ko_model *m, double *Sd_all, const int64_t *off, int nthreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int32_t d = 0; d < m->ndom; ++d) ko_assemble_local_schur(m, d, Sd_all + off[d]);
}

/* get_schur_rhs (:835-864): b_schur = b_G - sum_d R_d' A_GI,d A_II,d^{-1} b_I,d, ascending subdomains.
 * bI_cat = [b_Id[0]; b_Id[1]; ...], bI_off their offsets. */
void ko_schur_rhs(ko_model *m, const double *bI_cat, const int64_t *bI_off, const double *bG, double *out, int nthreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int32_t d = 0; d < m->ndom; ++d) {
    const ko_dom *D = &m->dom[d];
    const int32_t nI = D->nI, nG = D->nG;
    double *t = D->work, *w = D->work + nI, *o = m->loc + m->gbase[d];
    const double *b = bI_cat + bI_off[d];
    for (int32_t k = 0; k < nI; ++k) w[k] = b[D->perm[k]];
    chol_solve(D, w);                                                   /* :853-855 */
    for (int32_t k = 0; k < nI; ++k) t[D->perm[k]] = w[k];
    for (int32_t l = 0; l < nG; ++l) {                                  /* :857 */
      double s = 0.0;
      for (int32_t p = D->Gp[l]; p < D->Gp[l + 1]; ++p) s += D->Gx[p] * t[D->Gi[p]];
      o[l] = s;
    }
  }
  memcpy(out, bG, 8 * (size_t)m->n_gamma);                              /* :846-847 */
  for (int32_t d = 0; d < m->ndom; ++d) {                               /* :858-860 */
    const ko_dom *D = &m->dom[d];
    const double *o = m->loc + m->gbase[d];
    for (int32_t l = 0; l < D->nG; ++l) out[D->l2g[l]] -= o[l];
  }
}

/* KO_LONG_DOT (cpu_model.py, KB_ORACLE_VARIANT=ldot): the three inner products of an iteration accumulated in 80-bit
 * extended precision -- a rounding-floor probe (how far does the stopping index move when ONLY the PCG scalars are more
 * accurate, as the tree-shaped block sums of the GPU make them), never the default build. */
static double dot(const double *a, const double *b, int32_t n) {
#ifdef KO_LONG_DOT
  long double s = 0.0L;
  for (int32_t i = 0; i < n; ++i) s += (long double)a[i] * (long double)b[i];
  return (double)s;
#elif defined(KO_TREE_DOT)
  /* KB_ORACLE_VARIANT=tdot: plain doubles, summed as a tree (blocks of 256 entries: 32 strided lane sums combined by a
   * butterfly, then the block sums pairwise) -- the SHAPE of the GPU's block partials, not their exact order */
  const int32_t nb = (n + 255) / 256;
  double *part = (double *)malloc(8 * (size_t)(nb > 0 ? nb : 1));
  for (int32_t blk = 0; blk < nb; ++blk) {
    double lane[32];
    for (int l = 0; l < 32; ++l) lane[l] = 0.0;
    const int32_t i1 = (blk + 1) * 256 < n ? (blk + 1) * 256 : n;
    for (int32_t i = blk * 256; i < i1; ++i) lane[i & 31] += a[i] * b[i];
    for (int o = 16; o > 0; o >>= 1)
      for (int l = 0; l < o; ++l) lane[l] += lane[l + o];
    part[blk] = lane[0];
  }
  for (int32_t w = 1; w < nb; w <<= 1)
    for (int32_t k = 0; k + w < nb; k += 2 * w) part[k] += part[k + w];
  const double s = nb > 0 ? part[0] : 0.0;
  free(part);
  return s;
#elif defined(KO_BLAS_DOT)
  /* KB_ORACLE_VARIANT=bdot: the summation shape of a single-threaded SIMD BLAS ddot (OpenBLAS Haswell micro-kernel: four
   * vector accumulators of four doubles over 16 entries per step, horizontal sum, sequential tail) -- what Julia's
   * dot() does at least (OpenBLAS also splits n > 10000 over its threads) */
  double acc[16];
  for (int l = 0; l < 16; ++l) acc[l] = 0.0;
  const int32_t n16 = n & ~15;
  for (int32_t i = 0; i < n16; i += 16)
    for (int l = 0; l < 16; ++l) acc[l] += a[i + l] * b[i + l];
  for (int l = 0; l < 4; ++l) acc[l] = (acc[l] + acc[l + 4]) + (acc[l + 8] + acc[l + 12]);
  double s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  for (int32_t i = n16; i < n; ++i) s += a[i] * b[i];
  return s;
#else
  double s = 0.0;
  for (int32_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
#endif
}

/* cg.jl:67-109.  res_norm must hold maxit + 1 entries.  Returns it (= iterations + 1). */
int32_t ko_pcg(ko_model *m, const double *b, double *x, int32_t maxit, double eps, double *res_norm,
               int nthreads, int precond) {
  const int32_t n = m->n_gamma;
  double *r = (double *)malloc(8 * (size_t)n), *z = (double *)malloc(8 * (size_t)n),
         *p = (double *)malloc(8 * (size_t)n), *Ap = (double *)malloc(8 * (size_t)n);
  if (maxit == 0) maxit = n;
  int32_t it = 1;
  ko_apply_schur(m, x, Ap, nthreads);
  for (int32_t i = 0; i < n; ++i) r[i] = b[i] - Ap[i];
  double rTr = dot(r, r, n);
  if (precond) ko_apply_nn(m, r, z, nthreads); else memcpy(z, r, 8 * (size_t)n);
  double rTz = dot(r, z, n);
  memcpy(p, z, 8 * (size_t)n);
  res_norm[0] = sqrt(rTr);
  const double tol = eps * sqrt(dot(b, b, n));
  while (it < maxit && res_norm[it - 1] > tol) {
    ko_apply_schur(m, p, Ap, nthreads);
    const double d = dot(p, Ap, n);
    const double alpha = rTz / d;
    double beta = 1.0 / rTz;
    for (int32_t i = 0; i < n; ++i) x[i] += alpha * p[i];
    for (int32_t i = 0; i < n; ++i) r[i] -= alpha * Ap[i];
    rTr = dot(r, r, n);
    if (precond) ko_apply_nn(m, r, z, nthreads); else memcpy(z, r, 8 * (size_t)n);
    rTz = dot(r, z, n);
    beta *= rTz;
    for (int32_t i = 0; i < n; ++i) p[i] = beta * p[i] + z[i];
    it += 1;
    res_norm[it - 1] = sqrt(rTr);
  }
  free(r); free(z); free(p); free(Ap);
  return it;
}

/* Host-side sparse Cholesky for the interior blocks A_II,d  (stand-in for the reference's
 * `cholesky(A_IId[idom])`, SuiteSparse CHOLMOD via Fem/EllipticPdeDomainDecomposition.jl:1536;
 * CHOLMOD is not available in this image).
 *
 * Produces exactly what the Julia shim hands to libkrylov_b200.so: a fill-reducing permutation
 * p (F.p) and the simplicial lower factor L (sparse(F.L)) with L L^T = A[p,p], in CSC with
 * sorted row indices and the diagonal first in every column.
 *
 *   hp_nd_order     geometric nested dissection (coordinate bisection + graph vertex separators)
 *   hp_chol_count   elimination tree + column counts  -> nnz(L)
 *   hp_chol_factor  up-looking numeric factorisation
 *
 * All indices are 0-based int32 here; the Python wrapper converts to the ABI's Int64 1-based.
 * Build: gcc -O3 -fPIC -shared -fopenmp chol.c -o libhostprep.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t i32;

/* ------------------------------------------------------------------------------------------ */
/* nested dissection                                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const i32 *Ap, *Ai;
  const double *x, *y;
  i32 *side;   /* stamp per node: which subset call currently owns it */
  i32 *perm;   /* output: perm[new] = old */
  i32 npos;
  i32 leaf;
  i32 stamp;
  double *key; /* scratch */
} nd_ctx;

static __thread const double *g_key;
static int cmp_key(const void *a, const void *b) {
  double ka = g_key[*(const i32 *)a], kb = g_key[*(const i32 *)b];
  if (ka < kb) return -1;
  if (ka > kb) return 1;
  i32 ia = *(const i32 *)a, ib = *(const i32 *)b;
  return (ia > ib) - (ia < ib);
}

static void nd_rec(nd_ctx *c, i32 *idx, i32 n) {
  if (n <= c->leaf) {
    for (i32 k = 0; k < n; ++k) c->perm[c->npos++] = idx[k];
    return;
  }
  double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
  for (i32 k = 0; k < n; ++k) {
    double xv = c->x[idx[k]], yv = c->y[idx[k]];
    if (xv < x0) x0 = xv;
    if (xv > x1) x1 = xv;
    if (yv < y0) y0 = yv;
    if (yv > y1) y1 = yv;
  }
  g_key = (x1 - x0 >= y1 - y0) ? c->x : c->y;
  qsort(idx, (size_t)n, sizeof(i32), cmp_key);
  i32 nl = n / 2;
  i32 sl = ++c->stamp, sr = ++c->stamp;
  for (i32 k = 0; k < nl; ++k) c->side[idx[k]] = sl;
  for (i32 k = nl; k < n; ++k) c->side[idx[k]] = sr;
  /* separator = left nodes adjacent to a right node; move them to the tail of the left part */
  i32 keep = 0;
  i32 *sep = (i32 *)malloc(sizeof(i32) * (size_t)(nl > 0 ? nl : 1));
  i32 nsep = 0;
  for (i32 k = 0; k < nl; ++k) {
    i32 v = idx[k], is_sep = 0;
    for (i32 p = c->Ap[v]; p < c->Ap[v + 1]; ++p)
      if (c->side[c->Ai[p]] == sr) { is_sep = 1; break; }
    if (is_sep) sep[nsep++] = v; else idx[keep++] = v;
  }
  /* layout: [left' | right | sep] : shift right part down over the removed separator slots */
  memmove(idx + keep, idx + nl, sizeof(i32) * (size_t)(n - nl));
  memcpy(idx + keep + (n - nl), sep, sizeof(i32) * (size_t)nsep);
  free(sep);
  for (i32 k = 0; k < nsep; ++k) c->side[idx[keep + (n - nl) + k]] = 0; /* no longer in either half */
  nd_rec(c, idx, keep);
  nd_rec(c, idx + keep, n - nl);
  for (i32 k = 0; k < nsep; ++k) c->perm[c->npos++] = idx[keep + (n - nl) + k];
}

/* A: symmetric pattern (CSC or CSR, both triangles). perm[new] = old. */
int hp_nd_order(i32 n, const i32 *Ap, const i32 *Ai, const double *x, const double *y, i32 leaf,
                i32 *perm) {
  nd_ctx c;
  c.Ap = Ap; c.Ai = Ai; c.x = x; c.y = y; c.perm = perm; c.npos = 0;
  c.leaf = leaf < 1 ? 1 : leaf; c.stamp = 0;
  c.side = (i32 *)calloc((size_t)(n > 0 ? n : 1), sizeof(i32));
  i32 *idx = (i32 *)malloc(sizeof(i32) * (size_t)(n > 0 ? n : 1));
  if (!c.side || !idx) return -1;
  for (i32 k = 0; k < n; ++k) idx[k] = k;
  nd_rec(&c, idx, n);
  free(idx);
  free(c.side);
  return c.npos == n ? 0 : -2;
}

/* ------------------------------------------------------------------------------------------ */
/* Cholesky                                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* Upper triangle of C = A[p,p] in CSC: column k holds rows i <= k. */
static int permute_upper(i32 n, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *perm,
                         i32 **Cp_, i32 **Ci_, double **Cx_) {
  i32 *pinv = (i32 *)malloc(sizeof(i32) * (size_t)(n + 1));
  i32 *Cp = (i32 *)calloc((size_t)n + 2, sizeof(i32));
  if (!pinv || !Cp) return -1;
  for (i32 k = 0; k < n; ++k) pinv[perm[k]] = k;
  for (i32 j = 0; j < n; ++j)
    for (i32 p = Ap[j]; p < Ap[j + 1]; ++p) {
      i32 i2 = pinv[Ai[p]], j2 = pinv[j];
      if (i2 <= j2) Cp[j2 + 2]++;
    }
  for (i32 j = 0; j < n; ++j) Cp[j + 2] += Cp[j + 1];
  i32 nz = Cp[n + 1];
  i32 *Ci = (i32 *)malloc(sizeof(i32) * (size_t)(nz > 0 ? nz : 1));
  double *Cx = Ax ? (double *)malloc(sizeof(double) * (size_t)(nz > 0 ? nz : 1)) : NULL;
  if (!Ci || (Ax && !Cx)) return -1;
  for (i32 j = 0; j < n; ++j)
    for (i32 p = Ap[j]; p < Ap[j + 1]; ++p) {
      i32 i2 = pinv[Ai[p]], j2 = pinv[j];
      if (i2 <= j2) {
        i32 q = Cp[j2 + 1]++;
        Ci[q] = i2;
        if (Cx) Cx[q] = Ax[p];
      }
    }
  free(pinv);
  *Cp_ = Cp; *Ci_ = Ci; *Cx_ = Cx;
  return 0;
}

static void etree(i32 n, const i32 *Cp, const i32 *Ci, i32 *parent) {
  i32 *anc = (i32 *)malloc(sizeof(i32) * (size_t)(n + 1));
  for (i32 k = 0; k < n; ++k) {
    parent[k] = -1;
    anc[k] = -1;
    for (i32 p = Cp[k]; p < Cp[k + 1]; ++p) {
      i32 i = Ci[p];
      while (i != -1 && i < k) {
        i32 nxt = anc[i];
        anc[i] = k;
        if (nxt == -1) parent[i] = k;
        i = nxt;
      }
    }
  }
  free(anc);
}

/* pattern of row k of L (columns < k), topologically ordered in s[top..n-1] */
static i32 ereach(i32 n, const i32 *Cp, const i32 *Ci, i32 k, const i32 *parent, i32 *s, i32 *w) {
  i32 top = n;
  w[k] = k;
  for (i32 p = Cp[k]; p < Cp[k + 1]; ++p) {
    i32 i = Ci[p];
    if (i >= k) continue;
    i32 len = 0;
    for (; w[i] != k; i = parent[i]) {
      s[len++] = i;
      w[i] = k;
    }
    while (len > 0) s[--top] = s[--len];
  }
  return top;
}

/* colcount[j] = nnz of column j of L (diagonal included); parent = elimination tree.
 * Returns nnz(L) or <0. */
int64_t hp_chol_count(i32 n, const i32 *Ap, const i32 *Ai, const i32 *perm, i32 *parent,
                      i32 *colcount) {
  i32 *Cp, *Ci;
  double *Cx;
  if (permute_upper(n, Ap, Ai, NULL, perm, &Cp, &Ci, &Cx)) return -1;
  etree(n, Cp, Ci, parent);
  i32 *s = (i32 *)malloc(sizeof(i32) * (size_t)(n + 1));
  i32 *w = (i32 *)malloc(sizeof(i32) * (size_t)(n + 1));
  for (i32 k = 0; k < n; ++k) { w[k] = -1; colcount[k] = 1; }
  int64_t nnz = n;
  for (i32 k = 0; k < n; ++k) {
    i32 top = ereach(n, Cp, Ci, k, parent, s, w);
    for (i32 t = top; t < n; ++t) colcount[s[t]]++;
    nnz += n - top;
  }
  free(s); free(w); free(Cp); free(Ci);
  return nnz;
}

/* Numeric up-looking Cholesky. Lp must hold the column pointers implied by colcount (n+1
 * entries, prefix sums); Li/Lx sized nnz(L). Returns 0, or k+1 if the pivot of column k is
 * not positive. */
int hp_chol_factor(i32 n, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *perm,
                   const i32 *parent, const i32 *Lp, i32 *Li, double *Lx) {
  i32 *Cp, *Ci;
  double *Cx;
  if (permute_upper(n, Ap, Ai, Ax, perm, &Cp, &Ci, &Cx)) return -1;
  i32 *s = (i32 *)malloc(sizeof(i32) * (size_t)(n + 1));
  i32 *w = (i32 *)malloc(sizeof(i32) * (size_t)(n + 1));
  i32 *fill = (i32 *)malloc(sizeof(i32) * (size_t)(n + 1));
  double *xw = (double *)calloc((size_t)n + 1, sizeof(double));
  for (i32 k = 0; k < n; ++k) { w[k] = -1; fill[k] = Lp[k]; }
  int status = 0;
  for (i32 k = 0; k < n && !status; ++k) {
    i32 top = ereach(n, Cp, Ci, k, parent, s, w);
    xw[k] = 0.0;
    for (i32 p = Cp[k]; p < Cp[k + 1]; ++p)
      if (Ci[p] <= k) xw[Ci[p]] += Cx[p];
    double d = xw[k];
    xw[k] = 0.0;
    for (; top < n; ++top) {
      i32 j = s[top];
      double lkj = xw[j] / Lx[Lp[j]];
      xw[j] = 0.0;
      for (i32 p = Lp[j] + 1; p < fill[j]; ++p) xw[Li[p]] -= Lx[p] * lkj;
      d -= lkj * lkj;
      i32 q = fill[j]++;
      Li[q] = k;
      Lx[q] = lkj;
    }
    if (!(d > 0.0)) { status = k + 1; break; }
    i32 q = fill[k]++;
    Li[q] = k;
    Lx[q] = sqrt(d);
  }
  free(s); free(w); free(fill); free(xw); free(Cp); free(Ci); free(Cx);
  return status;
}

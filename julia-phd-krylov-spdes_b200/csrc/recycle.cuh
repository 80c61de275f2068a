// eigpcg / eigdefpcg of RecyclingKrylovSolvers (eigcg.jl:143-267, defcg.jl:337-473) as whole-solve entry points:
// included by api.cu inside its anonymous namespace (uses Ctx, apply_op, run_iterations, gemm_tn / gemm_nn, ...).
//
// Division of labour.  Every n_gamma-sized object (x, r, z, p, Ap, b, tvec, V, AV, W, AW) lives on the GPU.  The
// iterations between two restarts run as ONE device burst without host synchronisation (burst_core: the plain or
// deflated recurrence with the V column store and an alpha / beta history from which the host fills the tridiagonal
// part of V'AV, eigcg.jl:203-244); the restart iteration and the one after it, which touch tvec, are driven from the
// host over the same kernels.  The Rayleigh-Ritz restart (eigcg.jl:212-231: AV = A V by the multi-vector stream
// kernel, V'AV and V (Q Z) by the DMMA kernels) does its small dense algebra -- eigenvectors of the spdim x spdim
// projected matrix, SVD / rank of Y, eigen-decomposition of H -- with cuSOLVER / cuBLAS on the device, the role
// LAPACK plays in the reference (eigvecs, svd, rank, eigen at eigcg.jl:222-227).
#pragma once
// (api.cu includes <cublas_v2.h> and <cusolverDn.h> at file scope)

#define CUSOLVER_TRY(expr)                                                                                   \
  do {                                                                                                       \
    cusolverStatus_t s__ = (expr);                                                                           \
    if (s__ != CUSOLVER_STATUS_SUCCESS)                                                                      \
      throw KbError(KB200_ERR_CUDA, std::string(#expr) + ": cuSOLVER status " + std::to_string((int)s__));   \
  } while (0)
#define CUBLAS_TRY(expr)                                                                                     \
  do {                                                                                                       \
    cublasStatus_t s__ = (expr);                                                                             \
    if (s__ != CUBLAS_STATUS_SUCCESS)                                                                        \
      throw KbError(KB200_ERR_CUDA, std::string(#expr) + ": cuBLAS status " + std::to_string((int)s__));     \
  } while (0)

struct DenseLa {
  cusolverDnHandle_t cs = nullptr;
  cublasHandle_t cb = nullptr;
  DevBuf<double> work;
  DevBuf<int> info;
  explicit DenseLa(Ctx* c) {
    CUSOLVER_TRY(cusolverDnCreate(&cs));
    CUSOLVER_TRY(cusolverDnSetStream(cs, c->stream));
    CUBLAS_TRY(cublasCreate(&cb));
    CUBLAS_TRY(cublasSetStream(cb, c->stream));
    info.alloc(1);
  }
  ~DenseLa() {
    if (cs) cusolverDnDestroy(cs);
    if (cb) cublasDestroy(cb);
  }
  void check_info(Ctx* c, const char* what) {
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h != 0) throw KbError(KB200_ERR_INTERNAL, std::string(what) + " failed (info = " + std::to_string(h) + ")");
  }
  // eigen-decomposition of the symmetric m x m matrix in A (upper triangle read, like Julia's Symmetric()):
  // eigenvalues ascending in w, eigenvectors overwrite A
  void syevd(Ctx* c, int m, double* A, int lda, double* w) {
    int lwork = 0;
    CUSOLVER_TRY(cusolverDnDsyevd_bufferSize(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, m, A, lda, w, &lwork));
    if ((int64_t)work.n < lwork) work.alloc((size_t)lwork);
    CUSOLVER_TRY(cusolverDnDsyevd(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, m, A, lda, w, work.p, lwork, info.p));
    check_info(c, "syevd");
  }
  // generalized symmetric-definite problem A v = w B v (upper triangles read, like Julia's eigen(Symmetric, Symmetric) =
  // LAPACK sygvd): eigenvalues ascending in w, B-orthonormal eigenvectors overwrite A; B is overwritten by its factor
  void sygvd(Ctx* c, int m, double* A, int lda, double* B, int ldb, double* w) {
    int lwork = 0;
    CUSOLVER_TRY(cusolverDnDsygvd_bufferSize(cs, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, m, A, lda, B, ldb, w,
                                             &lwork));
    if ((int64_t)work.n < lwork) work.alloc((size_t)lwork);
    CUSOLVER_TRY(cusolverDnDsygvd(cs, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, m, A, lda, B, ldb, w, work.p,
                                  lwork, info.p));
    check_info(c, "sygvd (is the projected A-block positive definite?)");
  }
};

// Restart algebra of eigcg.jl:220-228 / defcg.jl:425-432: Tm (m x m, host, column-major, symmetric) -> Q Z (m x nev,
// device), eigenvalues of H (host), nev.
static int rayleigh_ritz(Ctx* c, DenseLa& la, const std::vector<double>& Tm, int m, int nvec, DevBuf<double>& dQZ,
                         std::vector<double>& vals) {
  cudaStream_t s = c->stream;
  const int ny = 2 * nvec;
  DevBuf<double> dT, dE1, dE2, dY, dS, dU, dVT, dQ, dTQ, dH, dw;
  dT.alloc((size_t)m * m);
  dE1.alloc((size_t)m * m);
  dE2.alloc((size_t)std::max(1, m - 1) * std::max(1, m - 1));
  dY.alloc((size_t)m * ny);
  dw.alloc((size_t)std::max(m, ny));
  CUDA_TRY(cudaMemcpyAsync(dT.p, Tm.data(), (size_t)m * m * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(dE1.p, dT.p, (size_t)m * m * 8, cudaMemcpyDeviceToDevice, s));
  la.syevd(c, m, dE1.p, m, dw.p);                                        // eigvecs(Tm)
  CUDA_TRY(cudaMemsetAsync(dY.p, 0, (size_t)m * ny * 8, s));
  CUDA_TRY(cudaMemcpyAsync(dY.p, dE1.p, (size_t)m * nvec * 8, cudaMemcpyDeviceToDevice, s));   // Y[:, 1:nvec]
  if (m > 1) {
    CUDA_TRY(cudaMemcpy2DAsync(dE2.p, (size_t)(m - 1) * 8, dT.p, (size_t)m * 8, (size_t)(m - 1) * 8, (size_t)(m - 1),
                               cudaMemcpyDeviceToDevice, s));           // Tm[1:m-1, 1:m-1]
    la.syevd(c, m - 1, dE2.p, m - 1, dw.p);
    const int take = std::min(nvec, m - 1);
    CUDA_TRY(cudaMemcpy2DAsync(dY.p + (size_t)nvec * m, (size_t)m * 8, dE2.p, (size_t)(m - 1) * 8, (size_t)(m - 1) * 8, (size_t)take,
                               cudaMemcpyDeviceToDevice, s));           // Y[1:m-1, nvec+1:end], last row zero
  }
  // svd(Y): left singular vectors and singular values; gesvd wants rows >= columns
  const int mn = std::min(m, ny);
  dS.alloc((size_t)mn);
  dQ.alloc((size_t)m * mn);
  const double one = 1.0, zero = 0.0;
  if (m >= ny) {
    dU.alloc((size_t)m * ny);
    int lwork = 0;
    CUSOLVER_TRY(cusolverDnDgesvd_bufferSize(la.cs, m, ny, &lwork));
    if ((int64_t)la.work.n < lwork) la.work.alloc((size_t)lwork);
    CUSOLVER_TRY(cusolverDnDgesvd(la.cs, 'S', 'N', m, ny, dY.p, m, dS.p, dU.p, m, nullptr, ny, la.work.p, lwork, nullptr, la.info.p));
    la.check_info(c, "gesvd");
    CUDA_TRY(cudaMemcpyAsync(dQ.p, dU.p, (size_t)m * mn * 8, cudaMemcpyDeviceToDevice, s));
  } else {
    // Y' (ny x m) = U' S V'^T  =>  left singular vectors of Y are the columns of V' = (VT)^T
    DevBuf<double> dYt;
    dYt.alloc((size_t)ny * m);
    CUBLAS_TRY(cublasDgeam(la.cb, CUBLAS_OP_T, CUBLAS_OP_N, ny, m, &one, dY.p, m, &zero, dYt.p, ny, dYt.p, ny));
    dVT.alloc((size_t)m * m);
    int lwork = 0;
    CUSOLVER_TRY(cusolverDnDgesvd_bufferSize(la.cs, ny, m, &lwork));
    if ((int64_t)la.work.n < lwork) la.work.alloc((size_t)lwork);
    CUSOLVER_TRY(cusolverDnDgesvd(la.cs, 'N', 'S', ny, m, dYt.p, ny, dS.p, nullptr, ny, dVT.p, m, la.work.p, lwork, nullptr, la.info.p));
    la.check_info(c, "gesvd");
    CUBLAS_TRY(cublasDgeam(la.cb, CUBLAS_OP_T, CUBLAS_OP_N, m, m, &one, dVT.p, m, &zero, dQ.p, m, dQ.p, m));
  }
  std::vector<double> sv((size_t)mn);
  CUDA_TRY(cudaMemcpyAsync(sv.data(), dS.p, (size_t)mn * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  // LinearAlgebra.rank default: singular values above min(size) * eps * sigma_max
  int nev = 0;
  if (mn > 0 && sv[0] > 0.0) {
    const double thr = (double)mn * 2.220446049250313e-16 * sv[0];
    for (int i = 0; i < mn; ++i) nev += sv[i] > thr;
  }
  if (nev < 1) throw KbError(KB200_ERR_INTERNAL, "Rayleigh-Ritz restart: Y has rank 0");
  if (std::getenv("KB200_DEBUG_RR"))
    std::fprintf(stderr, "[rr C] m %d nev %d sv0 %.3e sv[nev-1] %.3e sv[next] %.3e sv[last] %.3e\n", m, nev, sv[0], sv[nev - 1],
                 nev < mn ? sv[nev] : -1.0, sv[mn - 1]);
  // H = Q' (Tm Q), Q = U[:, 1:nev]; vals, Z = eigen(Symmetric(H)); Q Z
  dTQ.alloc((size_t)m * nev);
  dH.alloc((size_t)nev * nev);
  CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_N, CUBLAS_OP_N, m, nev, m, &one, dT.p, m, dQ.p, m, &zero, dTQ.p, m));
  CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_T, CUBLAS_OP_N, nev, nev, m, &one, dQ.p, m, dTQ.p, m, &zero, dH.p, nev));
  la.syevd(c, nev, dH.p, nev, dw.p);
  vals.resize((size_t)nev);
  CUDA_TRY(cudaMemcpyAsync(vals.data(), dw.p, (size_t)nev * 8, cudaMemcpyDeviceToHost, s));
  dQZ.alloc((size_t)m * nev);
  CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_N, CUBLAS_OP_N, m, nev, nev, &one, dQ.p, m, dH.p, nev, &zero, dQZ.p, m));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (std::getenv("KB200_DEBUG_RR"))
    std::fprintf(stderr, "[rr C]   vals %.6e %.6e %.6e ... %.6e\n", vals[0], vals[std::min(1, nev - 1)], vals[std::min(2, nev - 1)], vals[nev - 1]);
  return nev;
}

// The same restart for a pencil (G, B) -- the locally-optimal thick-restart variants (lotrrrdefpcg.jl:153-165,
// lotrhrdefpcg.jl:163-177): Y = [lowest nvec eigenvectors of (G, B) | those of the leading (m-1) blocks, last row zero],
// nev = rank(Y), Q = left singular vectors, eigen(Q'GQ, Q'BQ) -> Q Z (m x nev, device).  Bm empty: B = identity.
// H and Q'BQ are symmetrised (upper triangles read), see oracle/solvers.py::_lo_thick_restart.
static int rayleigh_ritz_pencil(Ctx* c, DenseLa& la, const std::vector<double>& Gm, const std::vector<double>& Bm, int m, int nvec,
                                DevBuf<double>& dQZ) {
  cudaStream_t s = c->stream;
  const bool gen = !Bm.empty();
  const int ny = 2 * nvec;
  require(m >= ny, "locally-optimal restart needs spdim >= 2 nvec");
  DevBuf<double> dG, dB, dE1, dB1, dE2, dB2, dY, dS, dU, dGQ, dH, dHB, dw;
  dG.alloc((size_t)m * m);
  dE1.alloc((size_t)m * m);
  dE2.alloc((size_t)std::max(1, m - 1) * std::max(1, m - 1));
  dY.alloc((size_t)m * ny);
  dw.alloc((size_t)m);
  CUDA_TRY(cudaMemcpyAsync(dG.p, Gm.data(), (size_t)m * m * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(dE1.p, dG.p, (size_t)m * m * 8, cudaMemcpyDeviceToDevice, s));
  if (gen) {
    dB.alloc((size_t)m * m);
    dB1.alloc((size_t)m * m);
    dB2.alloc((size_t)std::max(1, m - 1) * std::max(1, m - 1));
    CUDA_TRY(cudaMemcpyAsync(dB.p, Bm.data(), (size_t)m * m * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(dB1.p, dB.p, (size_t)m * m * 8, cudaMemcpyDeviceToDevice, s));
    la.sygvd(c, m, dE1.p, m, dB1.p, m, dw.p);
  } else {
    la.syevd(c, m, dE1.p, m, dw.p);
  }
  CUDA_TRY(cudaMemsetAsync(dY.p, 0, (size_t)m * ny * 8, s));
  CUDA_TRY(cudaMemcpyAsync(dY.p, dE1.p, (size_t)m * nvec * 8, cudaMemcpyDeviceToDevice, s));   // Y[:, 1:nvec]
  if (m > 1) {
    CUDA_TRY(cudaMemcpy2DAsync(dE2.p, (size_t)(m - 1) * 8, dG.p, (size_t)m * 8, (size_t)(m - 1) * 8, (size_t)(m - 1),
                               cudaMemcpyDeviceToDevice, s));
    if (gen) {
      CUDA_TRY(cudaMemcpy2DAsync(dB2.p, (size_t)(m - 1) * 8, dB.p, (size_t)m * 8, (size_t)(m - 1) * 8, (size_t)(m - 1),
                                 cudaMemcpyDeviceToDevice, s));
      la.sygvd(c, m - 1, dE2.p, m - 1, dB2.p, m - 1, dw.p);
    } else {
      la.syevd(c, m - 1, dE2.p, m - 1, dw.p);
    }
    const int take = std::min(nvec, m - 1);
    CUDA_TRY(cudaMemcpy2DAsync(dY.p + (size_t)nvec * m, (size_t)m * 8, dE2.p, (size_t)(m - 1) * 8, (size_t)(m - 1) * 8, (size_t)take,
                               cudaMemcpyDeviceToDevice, s));
  }
  dS.alloc((size_t)ny);
  dU.alloc((size_t)m * ny);
  int lwork = 0;
  CUSOLVER_TRY(cusolverDnDgesvd_bufferSize(la.cs, m, ny, &lwork));
  if ((int64_t)la.work.n < lwork) la.work.alloc((size_t)lwork);
  CUSOLVER_TRY(cusolverDnDgesvd(la.cs, 'S', 'N', m, ny, dY.p, m, dS.p, dU.p, m, nullptr, ny, la.work.p, lwork, nullptr, la.info.p));
  la.check_info(c, "gesvd");
  std::vector<double> sv((size_t)ny);
  CUDA_TRY(cudaMemcpyAsync(sv.data(), dS.p, (size_t)ny * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  int nev = 0;
  if (sv[0] > 0.0) {
    const double thr = (double)ny * 2.220446049250313e-16 * sv[0];   // LinearAlgebra.rank default
    for (int i = 0; i < ny; ++i) nev += sv[i] > thr;
  }
  if (nev < 1) throw KbError(KB200_ERR_INTERNAL, "locally-optimal restart: Y has rank 0");
  const double one = 1.0, zero = 0.0;
  dGQ.alloc((size_t)m * nev);
  dH.alloc((size_t)nev * nev);
  CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_N, CUBLAS_OP_N, m, nev, m, &one, dG.p, m, dU.p, m, &zero, dGQ.p, m));
  CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_T, CUBLAS_OP_N, nev, nev, m, &one, dU.p, m, dGQ.p, m, &zero, dH.p, nev));
  if (gen) {
    dHB.alloc((size_t)nev * nev);
    CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_N, CUBLAS_OP_N, m, nev, m, &one, dB.p, m, dU.p, m, &zero, dGQ.p, m));
    CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_T, CUBLAS_OP_N, nev, nev, m, &one, dU.p, m, dGQ.p, m, &zero, dHB.p, nev));
    la.sygvd(c, nev, dH.p, nev, dHB.p, nev, dw.p);
  } else {
    la.syevd(c, nev, dH.p, nev, dw.p);
  }
  dQZ.alloc((size_t)m * nev);
  CUBLAS_TRY(cublasDgemm(la.cb, CUBLAS_OP_N, CUBLAS_OP_N, m, nev, nev, &one, dU.p, m, dH.p, nev, &zero, dQZ.p, m));
  CUDA_TRY(cudaStreamSynchronize(s));
  return nev;
}

// Dense LU with partial pivoting of a small matrix (host): `WtAW \ .` of defcg.jl:272-274, :284, :301, :407
struct SmallLu {
  int k = 0;
  std::vector<double> lu;
  std::vector<int> piv;
  void factor(int kk, const std::vector<double>& A) {
    k = kk;
    lu = A;
    piv.resize((size_t)k);
    auto a = [&](int i, int j) -> double& { return lu[i + (size_t)j * k]; };
    for (int c0 = 0; c0 < k; ++c0) {
      int p = c0;
      for (int i = c0 + 1; i < k; ++i)
        if (std::fabs(a(i, c0)) > std::fabs(a(p, c0))) p = i;
      piv[c0] = p;
      if (a(p, c0) == 0.0) throw KbError(KB200_ERR_INPUT, "deflation: W' A W is singular (linearly dependent vectors)");
      if (p != c0)
        for (int j = 0; j < k; ++j) std::swap(a(p, j), a(c0, j));
      for (int i = c0 + 1; i < k; ++i) {
        const double f = a(i, c0) / a(c0, c0);
        a(i, c0) = f;
        if (f != 0.0)
          for (int j = c0 + 1; j < k; ++j) a(i, j) -= f * a(c0, j);
      }
    }
  }
  void solve(std::vector<double>& b) const {
    auto a = [&](int i, int j) { return lu[i + (size_t)j * k]; };
    for (int c0 = 0; c0 < k; ++c0) std::swap(b[c0], b[piv[c0]]);
    for (int i = 1; i < k; ++i) {
      double sacc = b[i];
      for (int j = 0; j < i; ++j) sacc -= a(i, j) * b[j];
      b[i] = sacc;
    }
    for (int i = k - 1; i >= 0; --i) {
      double sacc = b[i];
      for (int j = i + 1; j < k; ++j) sacc -= a(i, j) * b[j];
      b[i] = sacc / a(i, i);
    }
  }
};

// Work vectors of one recycling solve: columns of one device block
struct RecycleVecs {
  enum { X = 0, R, Z, P, AP, B, T, NCOL };
  Ctx* c;
  int n;
  DevBuf<double> wk;
  OpKind A;
  int precond;
  RecycleVecs(Ctx* ctx, OpKind a, int pre) : c(ctx), n((int)ctx->n_gamma), A(a), precond(pre) {
    wk.alloc((size_t)n * NCOL);
    CUDA_TRY(cudaMemsetAsync(wk.p, 0, (size_t)n * NCOL * 8, c->stream));
  }
  double* col(int j) { return wk.p + (size_t)j * n; }
  void applyA(int src, int dst) { apply_op(c, A, col(src), col(dst), nullptr, nullptr, nullptr, 0, nullptr); }
  void applyM(int src, int dst) {
    if (precond) apply_op(c, OP_NN, col(src), col(dst), nullptr, nullptr, nullptr, 0, nullptr);
    else CUDA_TRY(cudaMemcpyAsync(col(dst), col(src), (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
  }
  void axpby(double a, const double* x, double b, double* y) {
    kb::axpby_kernel<<<c->nb, c->vec_threads, 0, c->stream>>>(n, a, x, b, y);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
  }
  void axpby(double a, int x, double b, int y) { axpby(a, col(x), b, col(y)); }
  // y_host = A[:, 0:k]' x
  std::vector<double> gemv_t(const double* Am, int k, const double* x) {
    if ((int64_t)c->d_small.n < k) c->d_small.alloc((size_t)k);
    kb::gemv_t_kernel<<<(unsigned)k, c->vec_threads, 0, c->stream>>>(n, Am, n, x, c->d_small.p);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    std::vector<double> y((size_t)k);
    CUDA_TRY(cudaMemcpyAsync(y.data(), c->d_small.p, (size_t)k * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return y;
  }
  double dot(int x, int y) { return gemv_t(col(x), 1, col(y))[0]; }
  // y = b y + a A[:, 0:k] mu
  void gemv_n(double a, const double* Am, int k, const std::vector<double>& mu, double b, double* y) {
    if ((int64_t)c->d_small2.n < k) c->d_small2.alloc((size_t)k);
    CUDA_TRY(cudaMemcpyAsync(c->d_small2.p, mu.data(), (size_t)k * 8, cudaMemcpyHostToDevice, c->stream));
    kb::gemv_n_kernel<<<c->nb, c->vec_threads, (size_t)k * 8, c->stream>>>(n, k, a, Am, n, c->d_small2.p, b, y);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // mu is a host temporary
  }
  void residual() {  // r = b - A x
    applyA(X, AP);
    axpby(1.0, B, 0.0, R);
    axpby(-1.0, AP, 1.0, R);
  }
};

// C (ka x kb, host, column-major) = A[:, 0:ka]' B[:, 0:kb]
static std::vector<double> gemm_tn_host(Ctx* c, int ka, int kb_, const double* A, const double* B) {
  if ((int64_t)c->d_small.n < (int64_t)ka * kb_) c->d_small.alloc((size_t)ka * kb_);
  gemm_tn(c, ka, kb_, A, B, c->d_small.p);
  std::vector<double> C((size_t)ka * kb_);
  CUDA_TRY(cudaMemcpyAsync(C.data(), c->d_small.p, C.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return C;
}

struct BurstOut {
  int64_t done = 0;
  double rTz = 0.0;
  std::vector<double> alpha, beta, res;
};

// Up to m iterations of the (deflated) recurrence on the device without host synchronisation (kb200_pcg_burst).
static void burst_core(Ctx* c, OpKind A, int precond, double* wk5, const Defl* D, const Defl* DR, int64_t m, double* vbase,
                       double rTz, double tol, int64_t it0, int64_t maxit, BurstOut& out, double* pbase = nullptr) {
  const int n = (int)c->n_gamma;
  m = std::min<int64_t>(m, maxit - it0);
  if ((int64_t)c->d_res.n < maxit + 1) c->d_res.alloc(maxit + 1);
  IterCtx I;
  I.A = A;
  I.precond = precond;
  I.x = wk5;
  I.r = wk5 + (size_t)n;
  I.z = wk5 + (size_t)2 * n;
  I.p = wk5 + (size_t)3 * n;
  I.Ap = wk5 + (size_t)4 * n;
  if (precond) require(c->have_pinv, "kb200_upload_nn_pinv was not called");
  if (A == OP_SCHUR_DENSE) require(c->have_sd, "dense local Schur blocks were not uploaded / assembled");
  I.vbase = vbase;
  I.pbase = pbase;
  I.defl = D;
  I.deflR = DR;
  DevBuf<double> dhist;
  const size_t nh = (size_t)2 * (size_t)(it0 + m + 2);
  dhist.alloc(nh);
  CUDA_TRY(cudaMemsetAsync(dhist.p, 0, nh * 8, c->stream));
  I.hist = dhist.p;
  I.k0 = 1;
  kb::pcg_burst_init_kernel<<<1, 1, 0, c->stream>>>(c->d_state.p, rTz, tol, (int)it0, (int)maxit);
  c->launches++;
  run_iterations(c, I, 1, (int)m + 1, nullptr);
  // one synchronisation per burst: state, alpha/beta history, residual norms
  std::vector<double> hh(nh), rr((size_t)m);
  kb::PcgState hs;
  CUDA_TRY(cudaMemcpyAsync(&hs, c->d_state.p, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(hh.data(), dhist.p, nh * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(rr.data(), c->d_res.p + it0, (size_t)m * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  throw_on_fail(hs.fail);
  const int64_t done = (int64_t)hs.it - it0;
  require(done >= 0 && done <= m, "internal: burst iteration count out of range");
  out.done = done;
  out.rTz = hs.rTz;
  out.alpha.resize((size_t)done);
  out.beta.resize((size_t)done);
  out.res.resize((size_t)done);
  for (int64_t j = 0; j < done; ++j) {
    out.alpha[j] = hh[(size_t)2 * (size_t)(it0 + j + 1)];
    out.beta[j] = hh[(size_t)2 * (size_t)(it0 + j + 1) + 1];
    out.res[j] = rr[(size_t)j];
  }
}

// Symmetric(U): the upper triangle of the spdim x spdim accumulator (leading block m) as a full symmetric matrix
static std::vector<double> sym_upper(const std::vector<double>& VtAV, int ld, int m) {
  std::vector<double> T((size_t)m * m);
  for (int j = 0; j < m; ++j)
    for (int i = 0; i <= j; ++i) T[i + (size_t)j * m] = T[j + (size_t)i * m] = VtAV[i + (size_t)j * ld];
  return T;
}

// eigpcg (W == nullptr, eigcg.jl:143-267) / eigdefpcg (W given, defcg.jl:337-473).  b, x, W, V_out: host.
static void recycle_solve(Ctx* c, const double* b, double* x, const double* W, int nvec, int spdim, int64_t maxit, double rtol,
                          OpKind A, int precond, int64_t* it_out, double* res_norm, double* V_out, kb200_timings* tm) {
  const int n = (int)c->n_gamma;
  const bool defl = W != nullptr;
  require(nvec >= 1 && spdim > nvec && spdim <= 4096, "need 1 <= nvec < spdim <= 4096");
  require(!defl || spdim > nvec + 1, "eigdefpcg needs spdim > nvec + 1");
  if (maxit == 0) maxit = n;
  require(maxit >= 1 && maxit < INT_MAX, "bad maxit");
  kb200_timings T;
  std::memset(&T, 0, sizeof(T));
  const int64_t launches0 = c->launches;
  KernelTimer kt(c);
  DenseLa la(c);
  RecycleVecs v(c, A, precond);
  using RV = RecycleVecs;
  DevBuf<double> dV, dAV, dW, dAW;
  dV.alloc((size_t)n * spdim);
  CUDA_TRY(cudaMemsetAsync(dV.p, 0, (size_t)n * spdim * 8, c->stream));
  auto Vcol = [&](int j) { return dV.p + (size_t)j * n; };
  h2d(c, v.col(RV::B), b, n, &T);
  h2d(c, v.col(RV::X), x, n, &T);
  CUDA_TRY(cudaEventRecord(c->ev_a, c->stream));
  std::vector<double> VtAV((size_t)spdim * spdim, 0.0);
  auto VA = [&](int i, int j) -> double& { return VtAV[i + (size_t)j * spdim]; };
  std::vector<double> WtAW, WtW, WtAW_inv, WtW_inv;
  SmallLu luA, luW;
  DevBuf<double> dM, dMinv, dM2, dM2inv, dy, dmu, dy2, dmu2;
  Defl D, DR;
  if (defl) {
    dW.alloc((size_t)n * nvec);
    dAW.alloc((size_t)n * nvec);
    h2d(c, dW.p, W, (int64_t)n * nvec, &T);
    apply_op_multi(c, A, dW.p, dAW.p, nvec, MultiCtx(), &kt);                      // WtA' = A W        defcg.jl:357-360
    WtAW = gemm_tn_host(c, nvec, nvec, dAW.p, dW.p);
    WtW = gemm_tn_host(c, nvec, nvec, dW.p, dW.p);
    luA.factor(nvec, WtAW);
    luW.factor(nvec, WtW);
    invert_small(nvec, WtAW, WtAW_inv);
    invert_small(nvec, WtW, WtW_inv);
    auto up = [&](DevBuf<double>& buf, const std::vector<double>& src) {
      buf.alloc(src.size());
      CUDA_TRY(cudaMemcpyAsync(buf.p, src.data(), src.size() * 8, cudaMemcpyHostToDevice, c->stream));
    };
    up(dM, WtAW);
    up(dMinv, WtAW_inv);
    up(dM2, WtW);
    up(dM2inv, WtW_inv);
    dy.alloc(nvec);
    dmu.alloc(nvec);
    dy2.alloc(nvec);
    dmu2.alloc(nvec);
    D.k = nvec;
    D.W = dW.p;
    D.AW = dAW.p;
    D.M = dM.p;
    D.Minv = dMinv.p;
    D.y = dy.p;
    D.mu = dmu.p;
    DR = D;
    DR.M = dM2.p;
    DR.Minv = dM2inv.p;
    DR.y = dy2.p;
    DR.mu = dmu2.p;
    v.residual();                                                                   // x += W (WtAW \ W'r)  :362-366
    std::vector<double> mu = v.gemv_t(dW.p, nvec, v.col(RV::R));
    luA.solve(mu);
    v.gemv_n(1.0, dW.p, nvec, mu, 1.0, v.col(RV::X));
  } else {
    dAV.alloc((size_t)n * spdim);
  }
  std::vector<double> res((size_t)maxit + 1);
  int64_t it = 1;
  v.residual();
  double rTr = v.dot(RV::R, RV::R);
  v.applyM(RV::R, RV::Z);
  double rTz = v.dot(RV::R, RV::Z);
  v.axpby(1.0, RV::Z, 0.0, RV::P);
  if (defl) {                                                                       // p = z - W (WtAW \ WtA z)   :380-382
    std::vector<double> mu = v.gemv_t(dAW.p, nvec, v.col(RV::Z));
    luA.solve(mu);
    v.gemv_n(-1.0, dW.p, nvec, mu, 1.0, v.col(RV::P));
  }
  res[0] = std::sqrt(rTr);
  int ivec = 0;
  bool first_restart = true, just_restarted = false;
  if (defl) {
    for (int j = 0; j < nvec; ++j)
      for (int i = 0; i < nvec; ++i) VA(i, j) = WtAW[i + (size_t)j * nvec];        // :390
    CUDA_TRY(cudaMemcpyAsync(dV.p, dW.p, (size_t)n * nvec * 8, cudaMemcpyDeviceToDevice, c->stream));   // V[:, 1:nvec] = W
    ivec = nvec;
  }
  v.axpby(1.0 / std::sqrt(rTz), v.col(RV::Z), 0.0, Vcol(ivec));                     // V[:, ivec] = z / sqrt(rTz)
  double hlpr = 0.0;
  double bb = 0.0;
  for (int i = 0; i < n; ++i) bb += b[i] * b[i];
  const double tol = rtol * std::sqrt(bb);                                          // eps * norm2(b)
  auto restart = [&](int m) -> int {   // Rayleigh-Ritz on the leading m x m block; V[:, 1:nev] = V (Q Z)
    DevBuf<double> dQZ;
    std::vector<double> vals;
    const int nev = rayleigh_ritz(c, la, sym_upper(VtAV, spdim, m), m, nvec, dQZ, vals);
    gemm_nn(c, m, nev, dV.p, dQZ.p, dV.p);
    std::fill(VtAV.begin(), VtAV.end(), 0.0);
    for (int j = 0; j < nev; ++j) VA(j, j) = vals[j];
    return nev;
  };
  while (it < maxit && res[it - 1] > tol) {                                         // eigcg.jl:185 / defcg.jl:400
    const bool can_burst = defl ? (ivec < spdim - 1) : (!just_restarted && ivec < spdim - 1);
    if (can_burst) {
      BurstOut bo;
      burst_core(c, A, precond, v.col(RV::X), defl ? &D : nullptr, defl ? &DR : nullptr, spdim - 1 - ivec, Vcol(ivec + 1), rTz, tol, it,
                 maxit, bo);
      for (int64_t j = 0; j < bo.done; ++j) {
        it += 1;
        res[it - 1] = bo.res[j];
        VA(ivec, ivec) += 1.0 / bo.alpha[j];                                        // :206 / :417
        ivec += 1;                                                                  // V[:, ivec] stored by the device
        VA(ivec - 1, ivec) = -std::sqrt(bo.beta[j]) / bo.alpha[j];
        VA(ivec, ivec) = bo.beta[j] / bo.alpha[j];
      }
      rTz = bo.rTz;
      if (bo.done) just_restarted = false;
      continue;
    }
    v.applyA(RV::P, RV::AP);
    const double d = v.dot(RV::P, RV::AP);
    if (!(d > 0.0)) throw KbError(KB200_ERR_NOT_SPD, "p'Ap <= 0 encountered: operator is not positive definite along p");
    const double alpha = rTz / d;
    double beta = 1.0 / rTz;
    v.axpby(alpha, RV::P, 1.0, RV::X);
    v.axpby(-alpha, RV::AP, 1.0, RV::R);
    if (defl) {                                                                     // r -= W (WtW \ W'r)   defcg.jl:407
      std::vector<double> mu2 = v.gemv_t(dW.p, nvec, v.col(RV::R));
      luW.solve(mu2);
      v.gemv_n(-1.0, dW.p, nvec, mu2, 1.0, v.col(RV::R));
    }
    rTr = v.dot(RV::R, RV::R);
    v.applyM(RV::R, RV::Z);
    if (!defl && just_restarted) hlpr = std::sqrt(rTz);                             // eigcg.jl:194-196
    rTz = v.dot(RV::R, RV::Z);
    beta *= rTz;
    if (!defl && ivec == spdim - 1) v.axpby(-beta, RV::AP, 1.0, RV::T);             // tvec -= beta Ap   :199-201
    v.axpby(1.0, RV::Z, beta, RV::P);
    if (defl) {
      std::vector<double> mu = v.gemv_t(dAW.p, nvec, v.col(RV::Z));
      luA.solve(mu);
      v.gemv_n(-1.0, dW.p, nvec, mu, 1.0, v.col(RV::P));
    }
    it += 1;
    res[it - 1] = std::sqrt(rTr);
    VA(ivec, ivec) += 1.0 / alpha;                                                  // :206 / :417
    if (!defl && just_restarted) {
      v.axpby(1.0, RV::AP, 1.0, RV::T);                                             // tvec += Ap        :208
      const int nev = ivec;
      std::vector<double> col = v.gemv_t(dV.p, nev, v.col(RV::T));                  // :210
      for (int i = 0; i < nev; ++i) VA(i, ivec) = col[i] / hlpr;
      just_restarted = false;
    }
    if (ivec == spdim - 1) {                                                        // restart        :214 / :419
      if (!defl) {
        apply_op_multi(c, A, dV.p, dAV.p, spdim, MultiCtx(), &kt);                  // AV = A V        :215-218
        VtAV = gemm_tn_host(c, spdim, spdim, dV.p, dAV.p);                          // V'AV            :219
      } else if (first_restart) {
        std::vector<double> blk = gemm_tn_host(c, nvec, spdim - nvec, dAW.p, Vcol(nvec));   // WtA V   defcg.jl:421-423
        for (int j = 0; j < spdim - nvec; ++j)
          for (int i = 0; i < nvec; ++i) VA(i, nvec + j) = blk[i + (size_t)j * nvec];
        first_restart = false;
      }
      const int nev = restart(spdim);
      ivec = nev;
      v.axpby(1.0 / std::sqrt(rTz), v.col(RV::Z), 0.0, Vcol(ivec));
      VA(ivec, ivec) = beta / alpha;
      if (!defl) v.axpby(-beta, RV::AP, 0.0, RV::T);                                // tvec = -beta Ap   :240
      just_restarted = true;
    } else {
      just_restarted = false;
      ivec += 1;
      v.axpby(1.0 / std::sqrt(rTz), v.col(RV::Z), 0.0, Vcol(ivec));
      VA(ivec - 1, ivec) = -std::sqrt(beta) / alpha;
      VA(ivec, ivec) = beta / alpha;
    }
  }
  if (!just_restarted && ivec + 1 > nvec) {                                         // eigcg.jl:251-264 / defcg.jl:450-466
    const int m = ivec;
    if (defl && first_restart && m > nvec) {
      std::vector<double> blk = gemm_tn_host(c, nvec, m - nvec, dAW.p, Vcol(nvec));
      for (int j = 0; j < m - nvec; ++j)
        for (int i = 0; i < nvec; ++i) VA(i, nvec + j) = blk[i + (size_t)j * nvec];
    }
    if (m >= 2) restart(m);
  }
  CUDA_TRY(cudaEventRecord(c->ev_b, c->stream));
  d2h(c, x, v.col(RV::X), n, &T);
  d2h(c, V_out, dV.p, (int64_t)n * nvec, &T);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (!c->mute) std::copy(res.begin(), res.begin() + it, res_norm);
  *it_out = it;
  finish_timings(c, kt, T, launches0, tm);
}


// Post-hoc recycling: rrpcg / rrdefpcg (rrdefpcg.jl:204-275, :48-151) and hrpcg / hrdefpcg
// (hrdefpcg.jl:217-295, :48-164).  The (deflated) recurrence is the one of eigdefpcg's bursts -- p projected against
// W in the A-inner product, r re-orthogonalised against W (:111) -- with the first `cap` columns z / sqrt(rTz)
// (Rayleigh-Ritz) or p (harmonic Ritz) captured ON the device as they are produced; after convergence ONE projection:
//   RR:  eigen(Symmetric([W Z]' A [W Z]), I)                                   -> W2 = [W Z] vecs[:, 1:nvec]
//   HR:  eigen(Symmetric((A V)' M^-1 (A V)), blockdiag(W'AW, diag(P'AP)))      -> W2 = [W P] vecs[:, 1:nvec]
// A V by the multi-vector stream kernel, M^-1 (A V) by the batched DMMA Neumann-Neumann product, the Gram matrices
// and V vecs by the DMMA tall-skinny kernels, the small (generalized) eigenproblem with cuSOLVER syevd / sygvd.
// The reference projects with all cap columns even when the solve stopped before filling them (uninitialised
// memory); here, as in oracle/solvers.py, the projection uses the m = min(it, cap) columns that exist.
static void posthoc_solve(Ctx* c, const double* b, double* x, const double* W, int nvec, int spdim, int64_t maxit, double rtol, OpKind A,
                          int precond, bool harmonic, int64_t* it_out, double* res_norm, double* W2_out, kb200_timings* tm) {
  const int n = (int)c->n_gamma;
  const bool defl = W != nullptr;
  require(nvec >= 1 && spdim > nvec && spdim <= 4096, "need 1 <= nvec < spdim <= 4096");
  if (maxit == 0) maxit = n;
  require(maxit >= 1 && maxit < INT_MAX, "bad maxit");
  const int k0 = defl ? nvec : 0, cap = spdim - k0;
  kb200_timings T;
  std::memset(&T, 0, sizeof(T));
  const int64_t launches0 = c->launches;
  KernelTimer kt(c);
  DenseLa la(c);
  RecycleVecs v(c, A, precond);
  using RV = RecycleVecs;
  DevBuf<double> dV, dAV, dAW;                       // V = [W | captured columns]
  dV.alloc((size_t)n * spdim);
  dAV.alloc((size_t)n * spdim);
  CUDA_TRY(cudaMemsetAsync(dV.p, 0, (size_t)n * spdim * 8, c->stream));
  auto Ccol = [&](int j) { return dV.p + (size_t)(k0 + j) * n; };
  h2d(c, v.col(RV::B), b, n, &T);
  h2d(c, v.col(RV::X), x, n, &T);
  CUDA_TRY(cudaEventRecord(c->ev_a, c->stream));
  std::vector<double> WtAW, WtW, WtAW_inv, WtW_inv;
  SmallLu luA;
  DevBuf<double> dM, dMinv, dM2, dM2inv, dy, dmu, dy2, dmu2;
  Defl D, DR;
  if (defl) {
    h2d(c, dV.p, W, (int64_t)n * nvec, &T);                                          // V[:, 1:nvec] = W
    apply_op_multi(c, A, dV.p, dAV.p, nvec, MultiCtx(), &kt);                         // WtA' = A W      rrdefpcg.jl:73-79
    WtAW = gemm_tn_host(c, nvec, nvec, dAV.p, dV.p);
    WtW = gemm_tn_host(c, nvec, nvec, dV.p, dV.p);
    WtAW = sym_upper(WtAW, nvec, nvec);                                               // Symmetric()     :81, :83
    WtW = sym_upper(WtW, nvec, nvec);
    luA.factor(nvec, WtAW);
    invert_small(nvec, WtAW, WtAW_inv);
    invert_small(nvec, WtW, WtW_inv);
    auto up = [&](DevBuf<double>& buf, const std::vector<double>& src) {
      buf.alloc(src.size());
      CUDA_TRY(cudaMemcpyAsync(buf.p, src.data(), src.size() * 8, cudaMemcpyHostToDevice, c->stream));
    };
    up(dM, WtAW);
    up(dMinv, WtAW_inv);
    up(dM2, WtW);
    up(dM2inv, WtW_inv);
    dy.alloc(nvec);
    dmu.alloc(nvec);
    dy2.alloc(nvec);
    dmu2.alloc(nvec);
    D.k = nvec;
    D.W = dV.p;
    D.AW = dAV.p;
    D.M = dM.p;
    D.Minv = dMinv.p;
    D.y = dy.p;
    D.mu = dmu.p;
    DR = D;
    DR.M = dM2.p;
    DR.Minv = dM2inv.p;
    DR.y = dy2.p;
    DR.mu = dmu2.p;
    v.residual();                                                                     // x += W (WtAW \ W'r)   :86-89
    std::vector<double> mu = v.gemv_t(dV.p, nvec, v.col(RV::R));
    luA.solve(mu);
    v.gemv_n(1.0, dV.p, nvec, mu, 1.0, v.col(RV::X));
  }
  std::vector<double> res((size_t)maxit + 1);
  int64_t it = 1;
  v.residual();
  const double rTr = v.dot(RV::R, RV::R);
  v.applyM(RV::R, RV::Z);
  double rTz = v.dot(RV::R, RV::Z);
  v.axpby(1.0, RV::Z, 0.0, RV::P);
  if (defl) {                                                                         // p = z - W (WtAW \ WtA z)   :97-98
    std::vector<double> mu = v.gemv_t(dAV.p, nvec, v.col(RV::Z));
    luA.solve(mu);
    v.gemv_n(-1.0, dV.p, nvec, mu, 1.0, v.col(RV::P));
  }
  res[0] = std::sqrt(rTr);
  if (harmonic) v.axpby(1.0, v.col(RV::P), 0.0, Ccol(0));                             // P[:, 1] = p            hrdefpcg.jl:101
  else v.axpby(1.0 / std::sqrt(rTz), v.col(RV::Z), 0.0, Ccol(0));                     // Z[:, 1] = z / sqrt(rTz) rrdefpcg.jl:99
  double bb = 0.0;
  for (int i = 0; i < n; ++i) bb += b[i] * b[i];
  const double tol = rtol * std::sqrt(bb);
  while (it < maxit && res[it - 1] > tol) {
    // iterations it .. : columns it (0-based) .. cap-1 are still to be captured
    const int64_t left = cap - it;                       // iteration k stores column k (0-based) while k <= cap - 1
    const int64_t m = left > 0 ? left : (int64_t)c->check_every * 4;
    BurstOut bo;
    double* cbase = left > 0 ? Ccol((int)it) : nullptr;
    burst_core(c, A, precond, v.col(RV::X), defl ? &D : nullptr, defl ? &DR : nullptr, m, harmonic ? nullptr : cbase, rTz, tol, it, maxit, bo,
               harmonic ? cbase : nullptr);
    for (int64_t j = 0; j < bo.done; ++j) {
      it += 1;
      res[it - 1] = bo.res[j];
    }
    rTz = bo.rTz;
    if (bo.done == 0) break;
  }
  // ---- projection ---------------------------------------------------------------------------------------------------
  const int mc = (int)std::min<int64_t>(it, cap), mt = k0 + mc;
  apply_op_multi(c, A, Ccol(0), dAV.p + (size_t)k0 * n, mc, MultiCtx(), &kt);          // AZ / AP          :127-133
  std::vector<double> Tm = gemm_tn_host(c, mt, mt, dV.p, dAV.p);                       // V' A V
  DevBuf<double> dG, dB, dw, dvec;
  dw.alloc((size_t)mt);
  if (!harmonic) {
    std::vector<double> Ts = sym_upper(Tm, mt, mt);
    dG.alloc((size_t)mt * mt);
    CUDA_TRY(cudaMemcpyAsync(dG.p, Ts.data(), Ts.size() * 8, cudaMemcpyHostToDevice, c->stream));
    la.syevd(c, mt, dG.p, mt, dw.p);                                                  // eigen(VtAV, I)   :141-146
  } else {
    DevBuf<double> dMAV;
    dMAV.alloc((size_t)n * mt);
    if (precond) apply_op_multi(c, OP_NN, dAV.p, dMAV.p, mt, MultiCtx(), &kt);         // M \ (A V)        hrdefpcg.jl:128-141
    else CUDA_TRY(cudaMemcpyAsync(dMAV.p, dAV.p, (size_t)n * mt * 8, cudaMemcpyDeviceToDevice, c->stream));
    std::vector<double> G = gemm_tn_host(c, mt, mt, dAV.p, dMAV.p);                    // (A V)' M^-1 (A V)   :142-147
    std::vector<double> Gs = sym_upper(G, mt, mt), Bm((size_t)mt * mt, 0.0);
    for (int j = 0; j < k0; ++j)
      for (int i = 0; i < k0; ++i) Bm[i + (size_t)j * mt] = WtAW[i + (size_t)j * k0];  // VtAV = blockdiag(WtAW, diag(P'AP))   :149-156
    for (int j = k0; j < mt; ++j) Bm[j + (size_t)j * mt] = Tm[j + (size_t)j * mt];
    dG.alloc((size_t)mt * mt);
    dB.alloc((size_t)mt * mt);
    CUDA_TRY(cudaMemcpyAsync(dG.p, Gs.data(), Gs.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(dB.p, Bm.data(), Bm.size() * 8, cudaMemcpyHostToDevice, c->stream));
    la.sygvd(c, mt, dG.p, mt, dB.p, mt, dw.p);                                        // eigen(VtAinvMAV, VtAV)   :158
  }
  const int nret = std::min(nvec, mt);
  DevBuf<double> dW2;
  dW2.alloc((size_t)n * nvec);
  CUDA_TRY(cudaMemsetAsync(dW2.p, 0, (size_t)n * nvec * 8, c->stream));
  gemm_nn(c, mt, nret, dV.p, dG.p, dW2.p);                                            // W2 = [W C] vecs[:, 1:nvec]   :147-150
  CUDA_TRY(cudaEventRecord(c->ev_b, c->stream));
  d2h(c, x, v.col(RV::X), n, &T);
  d2h(c, W2_out, dW2.p, (int64_t)n * nvec, &T);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (!c->mute) std::copy(res.begin(), res.begin() + it, res_norm);
  *it_out = it;
  finish_timings(c, kt, T, launches0, tm);
}

// Thick-restart variants (trrrpcg / trrrdefpcg, trrrdefpcg.jl:234-355 / :48-199; trhrpcg / trhrdefpcg,
// trhrdefpcg.jl:247-389 / :48-208): the recurrence of posthoc_solve (deflated by the ORIGINAL basis W throughout, device
// bursts capturing Z = z / sqrt(r'z) or P = p columns), but the window of captured columns is cyclic and EVERY time it
// is full the recycle basis W2 is replaced by the nvec lowest (harmonic) Ritz vectors of [W2 | window].  Without a
// deflation basis the first window has spdim columns and projects onto them alone; later windows have spdim - nvec.
// Device layout: V = [W2 (nvec columns) | window (spdim columns)], AV likewise.  A solve that stops before its first
// window is full returns W2 = W (deflated variants) or zeros (the reference returns uninitialised memory there).  The
// harmonic variants keep the reference's structural zeros in V'AV (blockdiag(W2'AW2, diag(P'AP)), trhrdefpcg.jl:170-177).
//
// lo = true: the locally-optimal variants (lotrrrpcg / lotrrrdefpcg, lotrrrdefpcg.jl:252-389 / :48-214; lotrhrpcg /
// lotrhrdefpcg, lotrhrdefpcg.jl:267-428 / :48-229): the projection is the eigCG-type restart of rayleigh_ritz_pencil,
// W2 carries nev in [nvec, 2 nvec] columns (stored right in front of the window so that [W2 | window] stays
// contiguous), the window has spdim - nev columns and the first nvec columns of W2 are returned.
static void thick_restart_solve(Ctx* c, const double* b, double* x, const double* W, int nvec, int spdim, int64_t maxit, double rtol, OpKind A,
                                int precond, bool harmonic, bool lo, int64_t* it_out, double* res_norm, double* W2_out, kb200_timings* tm) {
  const int n = (int)c->n_gamma;
  const bool defl = W != nullptr;
  require(nvec >= 1 && spdim > nvec && spdim <= 4096, "need 1 <= nvec < spdim <= 4096");
  if (maxit == 0) maxit = n;
  require(maxit >= 1 && maxit < INT_MAX, "bad maxit");
  kb200_timings T;
  std::memset(&T, 0, sizeof(T));
  const int64_t launches0 = c->launches;
  KernelTimer kt(c);
  DenseLa la(c);
  RecycleVecs v(c, A, precond);
  using RV = RecycleVecs;
  const int w2cap = lo ? 2 * nvec : nvec;     // columns reserved for W2 in front of the window
  const int ncols = w2cap + spdim;
  if (lo) require(spdim >= 2 * nvec + 2, "locally-optimal thick restart needs spdim >= 2 nvec + 2");
  DevBuf<double> dV, dAV, dMAV, dW, dAW, dW2new;
  dV.alloc((size_t)n * ncols);
  dAV.alloc((size_t)n * ncols);
  dW2new.alloc((size_t)n * w2cap);
  CUDA_TRY(cudaMemsetAsync(dV.p, 0, (size_t)n * ncols * 8, c->stream));
  int nev = defl ? nvec : 0;                  // columns of W2 inside the search space
  auto Ccol = [&](int j) { return dV.p + (size_t)(w2cap + j) * n; };
  auto W2p = [&]() { return dV.p + (size_t)(w2cap - nev) * n; };
  auto AW2p = [&]() { return dAV.p + (size_t)(w2cap - nev) * n; };
  h2d(c, v.col(RV::B), b, n, &T);
  h2d(c, v.col(RV::X), x, n, &T);
  CUDA_TRY(cudaEventRecord(c->ev_a, c->stream));
  std::vector<double> WtAW, WtW, WtAW_inv, WtW_inv, W2tAW2;
  SmallLu luA;
  DevBuf<double> dM, dMinv, dM2, dM2inv, dy, dmu, dy2, dmu2;
  Defl D, DR;
  if (defl) {
    dW.alloc((size_t)n * nvec);
    dAW.alloc((size_t)n * nvec);
    h2d(c, dW.p, W, (int64_t)n * nvec, &T);
    apply_op_multi(c, A, dW.p, dAW.p, nvec, MultiCtx(), &kt);                          // WtA' = A W          trrrdefpcg.jl:80-86
    WtAW = sym_upper(gemm_tn_host(c, nvec, nvec, dAW.p, dW.p), nvec, nvec);            // Symmetric()         :87-90
    WtW = sym_upper(gemm_tn_host(c, nvec, nvec, dW.p, dW.p), nvec, nvec);
    W2tAW2 = WtAW;                                                                     // W2 = W              :91-93
    CUDA_TRY(cudaMemcpyAsync(W2p(), dW.p, (size_t)n * nvec * 8, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(AW2p(), dAW.p, (size_t)n * nvec * 8, cudaMemcpyDeviceToDevice, c->stream));
    luA.factor(nvec, WtAW);
    invert_small(nvec, WtAW, WtAW_inv);
    invert_small(nvec, WtW, WtW_inv);
    auto up = [&](DevBuf<double>& buf, const std::vector<double>& src) {
      buf.alloc(src.size());
      CUDA_TRY(cudaMemcpyAsync(buf.p, src.data(), src.size() * 8, cudaMemcpyHostToDevice, c->stream));
    };
    up(dM, WtAW);
    up(dMinv, WtAW_inv);
    up(dM2, WtW);
    up(dM2inv, WtW_inv);
    dy.alloc(nvec);
    dmu.alloc(nvec);
    dy2.alloc(nvec);
    dmu2.alloc(nvec);
    D.k = nvec;
    D.W = dW.p;
    D.AW = dAW.p;
    D.M = dM.p;
    D.Minv = dMinv.p;
    D.y = dy.p;
    D.mu = dmu.p;
    DR = D;
    DR.M = dM2.p;
    DR.Minv = dM2inv.p;
    DR.y = dy2.p;
    DR.mu = dmu2.p;
    v.residual();                                                                      // x += W (WtAW \ W'r)   :96-99
    std::vector<double> mu = v.gemv_t(dW.p, nvec, v.col(RV::R));
    luA.solve(mu);
    v.gemv_n(1.0, dW.p, nvec, mu, 1.0, v.col(RV::X));
  }
  std::vector<double> res((size_t)maxit + 1);
  int64_t it = 1;
  v.residual();
  const double rTr = v.dot(RV::R, RV::R);
  v.applyM(RV::R, RV::Z);
  double rTz = v.dot(RV::R, RV::Z);
  v.axpby(1.0, RV::Z, 0.0, RV::P);
  if (defl) {                                                                          // p = z - W (WtAW \ WtA z)   :108-109
    std::vector<double> mu = v.gemv_t(dAW.p, nvec, v.col(RV::Z));
    luA.solve(mu);
    v.gemv_n(-1.0, dW.p, nvec, mu, 1.0, v.col(RV::P));
  }
  res[0] = std::sqrt(rTr);
  if (harmonic) v.axpby(1.0, v.col(RV::P), 0.0, Ccol(0));                              // P[:, 1] = p
  else v.axpby(1.0 / std::sqrt(rTz), v.col(RV::Z), 0.0, Ccol(0));                      // Z[:, 1] = z / sqrt(rTz)   :110
  double bb = 0.0;
  for (int i = 0; i < n; ++i) bb += b[i] * b[i];
  const double tol = rtol * std::sqrt(bb);
  bool restarted = false;
  int wcap = spdim - nev;                     // columns of the current window
  int col = 1;                                // columns of the window filled so far
  DevBuf<double> dG, dB, dw;
  while (it < maxit && res[it - 1] > tol) {
    BurstOut bo;
    double* cbase = Ccol(col);
    burst_core(c, A, precond, v.col(RV::X), defl ? &D : nullptr, defl ? &DR : nullptr, wcap - col, harmonic ? nullptr : cbase, rTz, tol, it,
               maxit, bo, harmonic ? cbase : nullptr);
    for (int64_t j = 0; j < bo.done; ++j) {
      it += 1;
      res[it - 1] = bo.res[j];
    }
    rTz = bo.rTz;
    col += (int)bo.done;
    if (bo.done == 0) break;
    if (col < wcap) continue;
    // ---- the window is full: project onto [W2 | window]                                  :127-171 / trhrdefpcg.jl:127-190
    const int k0 = nev, mt = k0 + wcap;      // mt == spdim
    double* Vp = W2p();
    double* AVp = AW2p();
    if (k0 > 0 && restarted) apply_op_multi(c, A, Vp, AVp, k0, MultiCtx(), &kt);             // W2tA = A W2 (W2 changed)
    apply_op_multi(c, A, Ccol(0), dAV.p + (size_t)w2cap * n, wcap, MultiCtx(), &kt);         // AZ / AP
    std::vector<double> Tm = gemm_tn_host(c, mt, mt, Vp, AVp);                               // V' A V
    if (k0 > 0 && restarted) W2tAW2 = sym_upper(gemm_tn_host(c, k0, k0, AVp, Vp), k0, k0);   // Symmetric(W2tA * W2)
    dw.alloc((size_t)mt);
    dG.alloc((size_t)mt * mt);
    std::vector<double> Gs, Bm;
    if (!harmonic) {
      Gs = sym_upper(Tm, mt, mt);
      for (int j = 0; j < k0; ++j)
        for (int i2 = 0; i2 < k0; ++i2) Gs[i2 + (size_t)j * mt] = W2tAW2[i2 + (size_t)j * k0];
    } else {
      if ((int64_t)dMAV.n < (int64_t)n * mt) dMAV.alloc((size_t)n * ncols);
      if (precond) apply_op_multi(c, OP_NN, AVp, dMAV.p, mt, MultiCtx(), &kt);               // M \ (A V)
      else CUDA_TRY(cudaMemcpyAsync(dMAV.p, AVp, (size_t)n * mt * 8, cudaMemcpyDeviceToDevice, c->stream));
      Gs = sym_upper(gemm_tn_host(c, mt, mt, AVp, dMAV.p), mt, mt);                          // (A V)' M^-1 (A V)
      Bm.assign((size_t)mt * mt, 0.0);
      for (int j = 0; j < k0; ++j)
        for (int i2 = 0; i2 < k0; ++i2) Bm[i2 + (size_t)j * mt] = W2tAW2[i2 + (size_t)j * k0];
      for (int j = k0; j < mt; ++j) Bm[j + (size_t)j * mt] = Tm[j + (size_t)j * mt];         // diag(P' A P)
    }
    int nev_new = nvec;
    CUDA_TRY(cudaMemsetAsync(dW2new.p, 0, (size_t)n * w2cap * 8, c->stream));
    if (lo) {
      DevBuf<double> dQZ;
      nev_new = rayleigh_ritz_pencil(c, la, Gs, Bm, mt, nvec, dQZ);
      gemm_nn(c, mt, nev_new, Vp, dQZ.p, dW2new.p);                                          // W2 = [W2 window] Q Z
    } else {
      CUDA_TRY(cudaMemcpyAsync(dG.p, Gs.data(), Gs.size() * 8, cudaMemcpyHostToDevice, c->stream));
      if (!harmonic) {
        la.syevd(c, mt, dG.p, mt, dw.p);                                                     // eigen(VtAV, I)
      } else {
        dB.alloc((size_t)mt * mt);
        CUDA_TRY(cudaMemcpyAsync(dB.p, Bm.data(), Bm.size() * 8, cudaMemcpyHostToDevice, c->stream));
        la.sygvd(c, mt, dG.p, mt, dB.p, mt, dw.p);                                           // eigen(VtAinvMAV, VtAV)
      }
      gemm_nn(c, mt, std::min(nvec, mt), Vp, dG.p, dW2new.p);                                // W2 = [W2 window] vecs[:, 1:nvec]
    }
    nev = nev_new;
    CUDA_TRY(cudaMemcpyAsync(W2p(), dW2new.p, (size_t)n * nev * 8, cudaMemcpyDeviceToDevice, c->stream));
    restarted = true;
    wcap = spdim - nev;
    col = 0;
  }
  CUDA_TRY(cudaEventRecord(c->ev_b, c->stream));
  d2h(c, x, v.col(RV::X), n, &T);
  if (defl || restarted) d2h(c, W2_out, W2p(), (int64_t)n * nvec, &T);   // the first nvec columns of W2
  else std::fill(W2_out, W2_out + (size_t)n * nvec, 0.0);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (!c->mute) std::copy(res.begin(), res.begin() + it, res_norm);
  *it_out = it;
  finish_timings(c, kt, T, launches0, tm);
}

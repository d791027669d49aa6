// KKT system solve on the device:  H x + A'w = -q,  A x = b.
//   kkt_enqueue        KKTSystem.solvePD + blockSolve + solveWithCholFactor   KKTSystem.scala:99-246
//   kkt_solve_device   KKTSystem.solve fallback chain                          KKTSystem.scala:43-66
//   chol_enqueue       MatrixUtils.choleskySolve                               MatrixUtils.scala:468-516
// One attempt is a fixed kernel sequence with no host decision inside: numerical trouble (failed
// pivot, tiny diagonal, Schur complement not PD, residual above tolerance) is recorded in device
// flags and summarised in F_BAD; the host looks once per attempt (once per Newton step).
//
// Flop-minimal block elimination (same quantities as the reference, to rounding):
//   Y  = L^-1 [D A', D q]                      one TRSM   (reference: two dtrtrs, KKTSystem.scala:116-124)
//   S  = Yp' Yp  (= A H^-1 A', exactly symmetric by mirroring; reference symmetrises (R+R')/2, :139)
//   z  = -(b + Yp' yq),  w = K^-T K^-1 z,  x = -L^-T (yq + Yp w)
#include "kkt.cuh"
#include "vecops.cuh"

namespace cvxb {
namespace {

// qs = d o q -> also column p of Y; norms of qs and b
__global__ void __launch_bounds__(VT) kkt_rhs_kernel(int n, int p, const double* __restrict__ dr,
                                                     const double* __restrict__ q, const double* __restrict__ b,
                                                     double* __restrict__ qs, double* __restrict__ ycol, double* scal,
                                                     int* flag) {
  __shared__ double buf[33];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += VT) {
    double v = dr[i] * q[i];
    qs[i] = v;
    ycol[i] = v;
    s = fma(v, v, s);
  }
  s = block_sum(s, buf);
  double sb = 0;
  for (int i = threadIdx.x; i < p; i += VT) sb = fma(b[i], b[i], sb);
  sb = block_sum(sb, buf);
  if (threadIdx.x == 0) {
    scal[S_NORM_Q] = sqrt(s);
    scal[S_NORM_B] = sqrt(sb);
    scal[S_ERR1] = 0.0;
    scal[S_ERR2] = 0.0;
    flag[F_BAD] = 0;
    flag[F_ZERO_DIAG] = 0;
  }
}

// z = -(b + t)
__global__ void __launch_bounds__(VT) kkt_z_kernel(int p, const double* __restrict__ b, const double* __restrict__ t,
                                                   double* __restrict__ z) {
  for (int i = threadIdx.x; i < p; i += VT) z[i] = -(b[i] + t[i]);
}

// xs = sign * v ;  x = d o xs
__global__ void __launch_bounds__(VT) kkt_unscale_kernel(int n, double sign, const double* __restrict__ v,
                                                         const double* __restrict__ dr, double* __restrict__ xs,
                                                         double* __restrict__ x) {
  for (int i = threadIdx.x; i < n; i += VT) {
    double u = sign * v[i];
    xs[i] = u;
    x[i] = dr[i] * u;
  }
}

// err1 = ||L L'xs + d o (A'w) + qs|| / relsize(qs);  err2 = ||A x - b|| / relsize(b)   KKTSystem.scala:148-165
// and the verdict of the attempt.
__global__ void __launch_bounds__(VT) kkt_resid_kernel(int n, int p, const double* __restrict__ t2,
                                                       const double* __restrict__ t3, const double* __restrict__ dr,
                                                       const double* __restrict__ qs, const double* __restrict__ ax,
                                                       const double* __restrict__ b, double tol, int regularized,
                                                       double min_diag_tol, double* scal, int* flag) {
  __shared__ double buf[33];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += VT) {
    double r = t2[i] + dr[i] * t3[i] + qs[i];
    s = fma(r, r, s);
  }
  s = block_sum(s, buf);
  double s2 = 0;
  for (int i = threadIdx.x; i < p; i += VT) {
    double r = ax[i] - b[i];
    s2 = fma(r, r, s2);
  }
  s2 = block_sum(s2, buf);
  if (threadIdx.x == 0) {
    double e1 = relative_size(sqrt(s), scal[S_NORM_Q], tol);
    double e2 = relative_size(sqrt(s2), scal[S_NORM_B], tol);
    scal[S_ERR1] = e1;
    scal[S_ERR2] = e2;
    int bad = 0;
    if (flag[F_CHOL_H]) bad |= 1;
    if (!regularized && !(scal[S_MINDIAG_H] > min_diag_tol)) bad |= 2;
    if (flag[F_CHOL_S]) bad |= 4;
    if (!(e1 <= tol) || !(e2 <= tol)) bad |= 8;      // NaN counts as failure
    flag[F_BAD] = bad;
  }
}

// choleskySolve: ws = d o (sign*b) ; norm b
__global__ void __launch_bounds__(VT) chol_rhs_kernel(int n, double sign, const double* __restrict__ dr,
                                                      const double* __restrict__ b, double* __restrict__ ws,
                                                      double* scal, int* flag) {
  __shared__ double buf[33];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += VT) {
    double bi = sign * b[i];
    ws[i] = dr[i] * bi;
    s = fma(bi, bi, s);
  }
  s = block_sum(s, buf);
  if (threadIdx.x == 0) {
    scal[S_NORM_B] = sqrt(s);
    scal[S_NORM_Q] = 0.0;
    scal[S_ERR1] = 0.0;
    scal[S_ERR2] = 0.0;
    flag[F_BAD] = 0;
    flag[F_CHOL_S] = 0;
    flag[F_ZERO_DIAG] = 0;
  }
}

// relErr = ||H x - sign*b|| / relsize(b)   MatrixUtils.scala:496
__global__ void __launch_bounds__(VT) chol_resid_kernel(int n, double sign, const double* __restrict__ hx,
                                                        const double* __restrict__ b, double tol, int regularized,
                                                        double min_diag_tol, double* scal, int* flag) {
  __shared__ double buf[33];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += VT) {
    double r = hx[i] - sign * b[i];
    s = fma(r, r, s);
  }
  s = block_sum(s, buf);
  if (threadIdx.x == 0) {
    double e1 = relative_size(sqrt(s), scal[S_NORM_B], tol);
    scal[S_ERR1] = e1;
    int bad = 0;
    if (flag[F_CHOL_H]) bad |= 1;
    if (!regularized && !(scal[S_MINDIAG_H] > min_diag_tol)) bad |= 2;
    if (!(e1 <= tol)) bad |= 8;
    flag[F_BAD] = bad;
  }
}

__global__ void prefactored_reset_kernel(int* flag, double* scal) {
  flag[F_CHOL_H] = 0;
  scal[S_MINDIAG_H] = 1e300;
}

// qk = q - t   (t = A'b)
__global__ void __launch_bounds__(VT) sub_kernel(int n, const double* __restrict__ a, const double* __restrict__ b,
                                                 double* __restrict__ out) {
  for (int i = threadIdx.x; i < n; i += VT) out[i] = a[i] - b[i];
}

// S := (R + R') / 2 in place, exactly symmetric   (KKTSystem.scala:139)
__global__ void symmetrize_kernel(int p, double* __restrict__ S, int lds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p) return;
  for (int j = blockIdx.y; j < i; j += gridDim.y) {
    const double v = (S[(size_t)j * lds + i] + S[(size_t)i * lds + j]) * 0.5;
    S[(size_t)j * lds + i] = v;
    S[(size_t)i * lds + j] = v;
  }
}

template <typename T>
int dev_alloc(KktWork& W, T** ptr, size_t count) {
  void* q = W.arena ? W.arena->take((count ? count : 1) * sizeof(T)) : nullptr;
  if (q) { *ptr = (T*)q; return CVXB_OK; }
  CVXB_CUDA_OK(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
  W.owned.push_back(q);
  *ptr = (T*)q;
  return CVXB_OK;
}

}  // namespace

size_t kkt_work_bytes(int n, int p) {
  const size_t ldn = pad_ld(n), ldp = pad_ld(p);
  const size_t nblk = (n + NB - 1) / NB, pblk = (p + NB - 1) / NB;
  size_t d = ldn * n + nblk * NB * NB + ldn * (p + 1) + ldp * (p > 0 ? p : 1) + (pblk ? pblk : 1) * NB * NB + 9 * ldn + 3 * ldp;
  if (p > 0) d += 2 * ldn * (p + 1) + 128;       // Xlit, Blit
  return d * sizeof(double) + 32 * 256;
}

int kkt_work_alloc(Handle& h, KktWork& W, int n, int p, Arena* arena) {
  (void)h;
  W.arena = arena;
  W.n = n;
  W.p = p;
  W.ldn = pad_ld(n);
  W.ldp = pad_ld(p);
  int nblk = (n + NB - 1) / NB, pblk = (p + NB - 1) / NB;
  CVXB_TRY(dev_alloc(W, &W.L, (size_t)W.ldn * n));
  CVXB_TRY(dev_alloc(W, &W.invD, (size_t)nblk * NB * NB));
  CVXB_TRY(dev_alloc(W, &W.Y, (size_t)W.ldn * (p + 1)));
  if (p > 0) {
    CVXB_TRY(dev_alloc(W, &W.Xlit, (size_t)W.ldn * (p + 1)));
    CVXB_TRY(dev_alloc(W, &W.Blit, (size_t)W.ldn * (p + 1)));
  }
  CVXB_TRY(dev_alloc(W, &W.S, (size_t)W.ldp * (p > 0 ? p : 1)));
  CVXB_TRY(dev_alloc(W, &W.invDs, (size_t)(pblk ? pblk : 1) * NB * NB));
  double** nv[] = {&W.dr, &W.colsq, &W.qs, &W.xs, &W.t1, &W.t2, &W.t3, &W.qk, &W.dr2};
  for (double** v : nv) CVXB_TRY(dev_alloc(W, v, (size_t)W.ldn));
  double** pv[] = {&W.z, &W.ax, &W.tp};
  for (double** v : pv) CVXB_TRY(dev_alloc(W, v, (size_t)W.ldp));
  return CVXB_OK;
}

void kkt_work_free(KktWork& W) {
  for (void* q : W.owned) cudaFree(q);
  W.owned.clear();
  W = KktWork();
}

int fetch_status(Handle& h) {
  CVXB_CUDA_OK(cudaMemcpyAsync(h.h_flag, h.d_flag, NFLAG * sizeof(int), cudaMemcpyDeviceToHost, h.stream));
  CVXB_CUDA_OK(cudaMemcpyAsync(h.h_scal, h.d_scal, NSCAL * sizeof(double), cudaMemcpyDeviceToHost, h.stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(h.stream));
  h.status_reads++;
  if (h.h_flag[F_WAVE_ABORT]) {
    // a wavefront triangular solve gave up waiting for a predecessor block (its CTAs were not co-resident):
    // never expected under a cooperative launch; switch to the per-block kernels and report
    cudaMemsetAsync(h.d_flag + F_WAVE_ABORT, 0, sizeof(int), h.stream);
    if (h.wave_ready) { cudaFree(h.wave_ready); h.wave_ready = nullptr; }
    if (h.sk_ws) { cudaFree(h.sk_ws); h.sk_ws = nullptr; }       // the stream-K SYRK waits on peers the same way
    set_last_error("a cooperative kernel (wavefront triangular solve / stream-K SYRK) gave up waiting for a peer block "
                   "(blocks not co-resident); falling back to the non-cooperative kernels, retry the call");
    return CVXB_ECUDA;
  }
  return CVXB_OK;
}

void fill_info(Handle& h, cvxb_kkt_info* info, int path, int regularized) {
  if (!info) return;
  info->path = path;
  info->regularized = regularized;
  info->ruiz_sweeps = h.h_flag[F_RUIZ_SWEEPS];
  info->chol_info = h.h_flag[F_CHOL_H];
  info->min_diag = h.h_scal[S_MINDIAG_H];
  info->err1 = h.h_scal[S_ERR1];
  info->err2 = h.h_scal[S_ERR2];
}

int kkt_enqueue(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* A, int lda,
                const double* q, const double* b, double tol, bool regularize, bool skip_ruiz, double* x, double* w,
                bool prefactored) {
  NvtxRange nvtx("cvxb KKTSystem.solvePD (Ruiz, Cholesky + forward substitution, Schur complement)");
  const int n = W.n, p = W.p;
  if (prefactored) {
    // KKTSystem.solveWithCholFactor (KKTSystem.scala:99-167): the caller's factor L is used as is (d = 1, no
    // equilibration, no factorisation); only its diagonal blocks are inverted for the triangular solves
    CVXB_TRY(ruiz_equilibrate(h, n, Hm, ldh, W.dr, W.colsq, 0, P.ruizTol));     // d := 1
    CVXB_TRY(scaled_lower(h, n, Hm, ldh, nullptr, 0.0, W.L, W.ldn));
    CVXB_LAUNCH(h, prefactored_reset_kernel, 1, 1, 0, h.d_flag, h.d_scal);
    CVXB_TRY(invert_diag_blocks(h, n, W.L, W.ldn, W.invD));
  } else {
    if (!skip_ruiz) {
      CVXB_TRY(prof_begin(h, PROF_RUIZ));
      CVXB_TRY(ruiz_equilibrate(h, n, Hm, ldh, W.dr, W.colsq, P.ruizMaxSweeps, P.ruizTol, W.L, (size_t)W.ldn * n));
      CVXB_TRY(prof_end(h, PROF_RUIZ, 8.0 * n * n));       // bytes per sweep; sweeps taken are in F_RUIZ_SWEEPS
    }
    CVXB_TRY(scaled_lower(h, n, Hm, ldh, W.dr, regularize ? P.cholRegDelta : 0.0, W.L, W.ldn));
  }
  // right-hand sides  [D A', D q]
  CVXB_TRY(transpose_scale(h, p, n, A, lda, W.dr, W.Y, W.ldn));
  double* yq = W.Y + (size_t)p * W.ldn;
  CVXB_LAUNCH(h, kkt_rhs_kernel, 1, VT, 0, n, p, W.dr, q, b, W.qs, yq, h.d_scal, h.d_flag);
  // bugCompat & 2: the block elimination literally as the reference writes it (KKTSystem.scala:116-139): X = H^-1 [DA', Dq]
  // by two triangular solves, R = (DA')' X_p by a full GEMM, S = (R + R')/2 -- instead of S = Y'Y.  The two agree to rounding
  // until cond(H) ~ 1e20 (last barrier stages of LPs), where R loses positive definiteness and the reference gives up
  // (DESIGN.md section 2); this switch reproduces that.
  const bool literal = (P.bugCompat & 2) != 0 && p > 0 && W.Xlit && W.Blit;
  if (literal) CVXB_TRY(copy_matrix(h, n, p + 1, W.Y, W.ldn, W.Blit, W.ldn));
  if (prefactored) CVXB_TRY(trsm_lower(h, n, p + 1, W.L, W.ldn, W.invD, W.Y, W.ldn, false));
  else {  // factorisation with Y = L^-1 [DA', Dq] riding along
    CVXB_TRY(prof_begin(h, PROF_FACTOR));
    CVXB_TRY(potrf_lower_rhs(h, n, W.L, W.ldn, W.invD, F_CHOL_H, S_MINDIAG_H, W.Y, W.ldn, p + 1));
    CVXB_TRY(prof_end(h, PROF_FACTOR, (double)n * n * n / 3.0 + (double)n * n * (p + 1)));
  }
  if (literal) {
    double* xq = W.Xlit + (size_t)p * W.ldn;
    CVXB_TRY(copy_matrix(h, n, p + 1, W.Y, W.ldn, W.Xlit, W.ldn));
    CVXB_TRY(trsm_lower(h, n, p + 1, W.L, W.ldn, W.invD, W.Xlit, W.ldn, true));            // X = L^-T L^-1 [DA', Dq]
    GemmArgs gr{p, p, n, W.Blit, W.ldn, true, W.Xlit, W.ldn, true, W.S, W.ldp, 1.0, 0.0, 0};   // R = (DA')' X_p
    CVXB_TRY(gemm_dmma(h, gr));
    CVXB_LAUNCH(h, symmetrize_kernel, dim3((p + 127) / 128, p > 1024 ? 1024 : p), 128, 0, p, W.S, W.ldp);
    CVXB_TRY(gemv_t(h, n, p, 1.0, W.Blit, W.ldn, xq, 0.0, W.tp));                            // (DA')' H^-1 Dq
    CVXB_LAUNCH(h, kkt_z_kernel, 1, VT, 0, p, b, W.tp, w);
    CVXB_TRY(potrf_lower_rhs(h, p, W.S, W.ldp, W.invDs, F_CHOL_S, S_MINDIAG_S, w, W.ldp, 1));
    CVXB_TRY(trsm_lower(h, p, 1, W.S, W.ldp, W.invDs, w, W.ldp, true));
    CVXB_TRY(gemv_n(h, n, p, 1.0, W.Xlit, W.ldn, w, 1.0, xq));                               // H^-1 Dq + H^-1 DA' w
    CVXB_LAUNCH(h, kkt_unscale_kernel, 1, VT, 0, n, -1.0, xq, W.dr, W.xs, x);
  } else {
  // Schur complement S = Yp'Yp and its (plain, unregularised) Cholesky  KKTSystem.scala:126-140
  GemmArgs g{p, p, n, W.Y, W.ldn, true, W.Y, W.ldn, true, W.S, W.ldp, 1.0, 0.0, 2};
  g.streamk = true;
  CVXB_TRY(prof_begin(h, PROF_SCHUR));
  CVXB_TRY(gemm_dmma(h, g));
  CVXB_TRY(prof_end(h, PROF_SCHUR, (double)n * p * ((double)p + 1.0)));
  // z = -(b + A H^-1 q) = -(b + Yp' yq) ; w = K^-T K^-1 z  (the forward half rides along with the factorisation of S)
  CVXB_TRY(gemv_t(h, n, p, 1.0, W.Y, W.ldn, yq, 0.0, W.tp));
  CVXB_LAUNCH(h, kkt_z_kernel, 1, VT, 0, p, b, W.tp, w);
  CVXB_TRY(potrf_lower_rhs(h, p, W.S, W.ldp, W.invDs, F_CHOL_S, S_MINDIAG_S, w, W.ldp, 1));
  CVXB_TRY(trsm_lower(h, p, 1, W.S, W.ldp, W.invDs, w, W.ldp, true));
  // x = -L^-T (yq + Yp w)
  CVXB_TRY(gemv_n(h, n, p, 1.0, W.Y, W.ldn, w, 1.0, yq));
  CVXB_TRY(trsm_lower(h, n, 1, W.L, W.ldn, W.invD, yq, W.ldn, true));
  CVXB_LAUNCH(h, kkt_unscale_kernel, 1, VT, 0, n, -1.0, yq, W.dr, W.xs, x);
  }
  // residuals on the equilibrated system, with L L' in place of Q  (KKTSystem.scala:148-154)
  CVXB_TRY(gemv_t(h, n, n, 1.0, W.L, W.ldn, W.xs, 0.0, W.t1));
  CVXB_TRY(gemv_n(h, n, n, 1.0, W.L, W.ldn, W.t1, 0.0, W.t2));
  CVXB_TRY(gemv_t(h, p, n, 1.0, A, lda, w, 0.0, W.t3));
  CVXB_TRY(gemv_n(h, p, n, 1.0, A, lda, x, 0.0, W.ax));
  CVXB_LAUNCH(h, kkt_resid_kernel, 1, VT, 0, n, p, W.t2, W.t3, W.dr, W.qs, W.ax, b, tol, regularize ? 1 : 0,
              P.cholMinDiag, h.d_scal, h.d_flag);
  return CVXB_OK;
}

namespace {

// one path of KKTSystem.solve: plain attempt, then the regularised retry of regularizedCholesky
// returns CVXB_OK with *ok = whether the path produced an accepted solution
int kkt_try_path(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* A, int lda,
                 const double* q, const double* b, double tol, double* x, double* w, bool first_attempt_done,
                 int* regularized, bool* ok) {
  *regularized = 0;
  if (!first_attempt_done) {
    CVXB_TRY(kkt_enqueue(h, W, P, Hm, ldh, A, lda, q, b, tol, false, false, x, w));
    CVXB_TRY(fetch_status(h));
  }
  int bad = h.h_flag[F_BAD];
  if (bad & 3) {   // pivot failure or min diag <= 1e-7: second attempt on Q + delta I  (MatrixUtils.scala:452-461)
    *regularized = 1;
    CVXB_TRY(kkt_enqueue(h, W, P, Hm, ldh, A, lda, q, b, tol, true, true, x, w));
    CVXB_TRY(fetch_status(h));
    bad = h.h_flag[F_BAD];
  }
  *ok = (bad == 0);
  return CVXB_OK;
}

}  // namespace

int kkt_solve_fallbacks(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* A,
                        int lda, const double* q, const double* b, double tol, double* x, double* w,
                        cvxb_kkt_info* info) {
  const int n = W.n, p = W.p;
  bool ok = false;
  int reg = 0;
  CVXB_TRY(kkt_try_path(h, W, P, Hm, ldh, A, lda, q, b, tol, x, w, true, &reg, &ok));
  if (ok) { fill_info(h, info, 0, reg); return CVXB_OK; }
  // path 1: K = H + A'A, z = q - A'b   (KKTSystem.scala:57-59)
  if (!W.Hk) {
    void* ptr = nullptr;
    CVXB_CUDA_OK(cudaMalloc(&ptr, (size_t)W.ldn * n * sizeof(double)));
    W.owned.push_back(ptr);
    W.Hk = (double*)ptr;
  }
  CVXB_TRY(copy_matrix(h, n, n, Hm, ldh, W.Hk, W.ldn));
  GemmArgs g{n, n, p, A, lda, true, A, lda, true, W.Hk, W.ldn, 1.0, 1.0, 2};
  g.streamk = true;
  CVXB_TRY(gemm_dmma(h, g));
  CVXB_TRY(gemv_t(h, p, n, 1.0, A, lda, b, 0.0, W.t1));
  CVXB_LAUNCH(h, sub_kernel, 1, VT, 0, n, q, W.t1, W.qk);
  CVXB_TRY(kkt_try_path(h, W, P, W.Hk, W.ldn, A, lda, W.qk, b, tol, x, w, false, &reg, &ok));
  if (ok) { fill_info(h, info, 1, reg); return CVXB_OK; }
  fill_info(h, info, 2, reg);
  // path 2: KKTSystem.kktSymSolve -- decomposition of the full (n+p)^2 KKT matrix  (KKTSystem.scala:63, 283-310)
  return kkt_sym_solve_device(h, n, p, Hm, ldh, A, lda, q, b, tol, x, w);
}

int kkt_solve_device(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* A, int lda,
                     const double* q, const double* b, double tol, double* x, double* w, cvxb_kkt_info* info) {
  CVXB_TRY(kkt_enqueue(h, W, P, Hm, ldh, A, lda, q, b, tol, false, false, x, w));
  CVXB_TRY(fetch_status(h));
  if (h.h_flag[F_BAD] == 0) { fill_info(h, info, 0, 0); return CVXB_OK; }
  return kkt_solve_fallbacks(h, W, P, Hm, ldh, A, lda, q, b, tol, x, w, info);
}

int chol_enqueue(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* b,
                 double rhs_sign, double tol, bool regularize, bool skip_ruiz, double* x) {
  const int n = W.n;
  if (!skip_ruiz) CVXB_TRY(ruiz_equilibrate(h, n, Hm, ldh, W.dr, W.colsq, P.ruizMaxSweeps, P.ruizTol, W.L, (size_t)W.ldn * n));
  CVXB_TRY(scaled_lower(h, n, Hm, ldh, W.dr, regularize ? P.cholRegDelta : 0.0, W.L, W.ldn));
  CVXB_LAUNCH(h, chol_rhs_kernel, 1, VT, 0, n, rhs_sign, W.dr, b, W.qs, h.d_scal, h.d_flag);
  CVXB_TRY(potrf_lower_rhs(h, n, W.L, W.ldn, W.invD, F_CHOL_H, S_MINDIAG_H, W.qs, W.ldn, 1));     // w = L^-1 (d o b) rides along
  CVXB_TRY(trsm_lower(h, n, 1, W.L, W.ldn, W.invD, W.qs, W.ldn, true));
  CVXB_LAUNCH(h, kkt_unscale_kernel, 1, VT, 0, n, 1.0, W.qs, W.dr, W.xs, x);
  CVXB_TRY(gemv_n(h, n, n, 1.0, Hm, ldh, x, 0.0, W.t1));
  CVXB_LAUNCH(h, chol_resid_kernel, 1, VT, 0, n, rhs_sign, W.t1, b, tol, regularize ? 1 : 0, P.cholMinDiag, h.d_scal,
              h.d_flag);
  return CVXB_OK;
}

int chol_solve_retry(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* b,
                     double rhs_sign, double tol, double* x, cvxb_kkt_info* info) {
  int bad = h.h_flag[F_BAD], reg = 0;
  if (bad & 3) {
    reg = 1;
    CVXB_TRY(chol_enqueue(h, W, P, Hm, ldh, b, rhs_sign, tol, true, true, x));
    CVXB_TRY(fetch_status(h));
    bad = h.h_flag[F_BAD];
  }
  fill_info(h, info, 0, reg);
  if (bad == 0) return CVXB_OK;
  set_last_error("choleskySolve: %s (chol info %d, min diag %.3g, relative error %.3g > tol %.3g)",
                 (bad & 1) ? "matrix not positive definite" : "error exceeds tolerance", h.h_flag[F_CHOL_H],
                 h.h_scal[S_MINDIAG_H], h.h_scal[S_ERR1], tol);
  return CVXB_ELINSOLVE;
}

int chol_solve_device(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* b,
                      double rhs_sign, double tol, double* x, cvxb_kkt_info* info) {
  CVXB_TRY(chol_enqueue(h, W, P, Hm, ldh, b, rhs_sign, tol, false, false, x));
  CVXB_TRY(fetch_status(h));
  if (h.h_flag[F_BAD] == 0) { fill_info(h, info, 0, 0); return CVXB_OK; }
  return chol_solve_retry(h, W, P, Hm, ldh, b, rhs_sign, tol, x, info);
}

}  // namespace cvxb

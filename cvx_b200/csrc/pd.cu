// Primal-dual interior-point solver, device resident (B&V 11.7).
//   kktMatrix_noEqs   H_pd = hess f + sum_i -(lam_i/f_i) grad g_i grad g_i'     PrimalDualSolver.scala:216-240
//   rhs1 / deltaLambda / residuals / surrogate gap                              :63-144,162-209,289-297
//   kktSystem_noEqs / _withEqs                                                  :254-285
//   lineSearch_noEQs / _withEQs                                                 :311-374,478-543
//   solve_noEQs / solve_withEQs / solve                                         :381-460,550-641
// Defects D1 (line search restarts from the initial iterate) and D2 (sign of the gradient term when
// equalities are present) of solve_withEQs are reproduced only when params.bugCompat != 0.
//
// As in the barrier path the line search costs O(m+n+p) per trial: every residual is affine in the step
// along the ray except grad f(x + s dx) for the KL objective, which is elementwise.
#include "solver.cuh"
#include "vecops.cuh"

using namespace cvxb;

namespace cvxb {

int problem_alloc(Handle& h, int n, int m, int p, int objective, cvxb_problem_s** out, int mq, int kd);
int quad_refresh(cvxb_problem_s* P);
int quad_direction(cvxb_problem_s* P, const double* dir);
int quad_hessian_terms(cvxb_problem_s* P, const double* c);
int composed_refresh(cvxb_problem_s* P);
int composed_hessian(cvxb_problem_s* P, double t, const double* t_dev);
int dual_refresh(cvxb_problem_s* P);
int dual_hessian(cvxb_problem_s* P, double t, const double* t_dev);
int upload_vec(Handle& h, double* dst, const double* src, int n);
int download_vec(Handle& h, double* dst, const double* src, int n);

namespace {

constexpr double IN_SET_FACTOR = 1.0 + 3e-16;

__device__ __forceinline__ double grad_f0(int kind, int n, double xj, double aj, double pxj, double pw = 2.0) {
  if (kind == CVXB_OBJ_PNORM) {
    const double sg = fabs(xj) < 1e-14 ? 0.0 : (xj > 0 ? 1.0 : -1.0);
    return sg * pw * pow(sg * xj, pw - 1.0);
  }
  if (kind == CVXB_OBJ_LINEAR) return aj;
  if (kind == CVXB_OBJ_QUADRATIC) return aj + pxj;
  if (kind == CVXB_OBJ_COMPOSED) return pxj;          // F' grad f_inner(z0 + F u)
  if (kind == CVXB_OBJ_KLDUAL) return aj - pxj;       // w - B y
  return 1.0 + log(xj) + log((double)n);
}

// lam0 = -1/(g(x)-ub)   ConstraintSet.scala:116-120 ; gx holds G x
__global__ void __launch_bounds__(VT) pd_lambda0_kernel(int m, const double* __restrict__ gx, const double* __restrict__ gr,
                                                        const double* __restrict__ ub, const double* __restrict__ qcorr,
                                                        double* __restrict__ lam) {
  for (int i = threadIdx.x; i < m; i += VT) lam[i] = -1.0 / ((gr[i] + gx[i] - (qcorr ? qcorr[i] : 0.0)) - ub[i]);
}

// f = g(x) - ub ; weights -lam/f ; 1/(t f) ; surrogate gap -f.lam ; checks f < 0, lam > 0
__global__ void __launch_bounds__(VT) pd_cnt_kernel(int m, double t, const double* __restrict__ gr, const double* __restrict__ ub,
                                                    double* __restrict__ gx, const double* __restrict__ lam,
                                                    double* __restrict__ fvec, double* __restrict__ wts,
                                                    double* __restrict__ inv, const double* __restrict__ qcorr,
                                                    double* scal, int* flag) {
  __shared__ double buf[33];
  __shared__ int ibuf[33];
  int bad = 0;
  double gap = 0.0;
  for (int i = threadIdx.x; i < m; i += VT) {
    double g = gr[i] + gx[i] - (qcorr ? qcorr[i] : 0.0);
    gx[i] = g;
    double f = g - ub[i];
    fvec[i] = f;
    if (!(f < 0.0)) bad |= 1;
    if (!(lam[i] > 0.0)) bad |= 2;
    wts[i] = -(lam[i] / f);
    inv[i] = 1.0 / (t * f);
    gap = fma(-f, lam[i], gap);
  }
  gap = block_sum(gap, buf);
  bad = block_or(bad, ibuf);
  if (threadIdx.x == 0) {
    scal[S_PD_GAP] = gap;
    flag[F_PD_NOTNEG] = bad & 1;
    flag[F_PD_LAMNEG] = (bad >> 1) & 1;
  }
}

// v = -grad f + G'(1/(t f)) ; rd0 = G'lam + A'nu ; q = (+-)v + A'nu ; pres = Ax - b ; negpres = -(Ax-b)
__global__ void __launch_bounds__(VT) pd_rhs_kernel(int n, int p, int kind, int bug, const double* __restrict__ x,
                                                    const double* __restrict__ a, const double* __restrict__ Px,
                                                    const double* __restrict__ gt, const double* __restrict__ Atnu,
                                                    double* __restrict__ rd0, double* __restrict__ v, double* __restrict__ q,
                                                    const double* __restrict__ ax, const double* __restrict__ b,
                                                    double* __restrict__ pres, double* __restrict__ negpres, double pw) {
  for (int j = threadIdx.x; j < n; j += VT) {
    double gf = grad_f0(kind, n, x[j], a ? a[j] : 0.0, Px ? Px[j] : 0.0, pw);
    double vj = -gf + gt[j];
    double an = p > 0 ? Atnu[j] : 0.0;
    v[j] = vj;
    rd0[j] = rd0[j] + an;
    q[j] = (bug ? vj : -vj) + an;
  }
  for (int i = threadIdx.x; i < p; i += VT) {
    double r = ax[i] - b[i];
    pres[i] = r;
    negpres[i] = -r;
  }
}

// dlam = (-lam (G dx) + r_cent) / f,  r_cent = -lam f - 1/t      PrimalDualSolver.scala:184-209
__global__ void __launch_bounds__(VT) pd_dlam_kernel(int m, double t, const double* __restrict__ lam,
                                                     const double* __restrict__ fvec, const double* __restrict__ Gd,
                                                     double* __restrict__ dlam) {
  for (int i = threadIdx.x; i < m; i += VT) {
    double f = fvec[i];
    double rc = -lam[i] * f - 1.0 / t;
    dlam[i] = (-lam[i] * Gd[i] + rc) / f;
  }
}

struct PdLs {
  int m, n, p, kind, withEqs;
  double t, alpha, beta, frac, pw;
  // base point of the search (current iterate, or the initial one under bugCompat) and direction
  const double *gx, *ub, *lam, *x, *nu, *rd0, *pres, *Px, *a;
  const double *Gd, *dlam, *dx, *dnu, *rd1, *Adx, *Pd;
  const double *qq, *rd2;   // quadratic constraints: d'P_k d / 2 per row and the s^2 term of the dual residual (or NULL)
  // iterate to write
  double *xo, *lamo, *nuo;
  // objectives whose gradient at a trial point needs a matrix-vector product (one CTA does it: these are the reduced
  // and dual problems, small next to the primal ones):
  //   CVXB_OBJ_COMPOSED  grad = F' grad f_inner(xf + s dxf)      M = F (mrows x n), vin = xf, dvin = dxf, inner family mkind
  //   CVXB_OBJ_KLDUAL    grad = w - B (y o exp(-s v))            M = B (n x mcols), vin = y,  dvin = v
  const double *M, *vin, *dvin;
  int ldM, mrows, mcols, mkind;
  double *tmp, *gf;         // scratch: length max(mrows, mcols) and n
};

// gf := gradient of the objective at the trial point for the two matrix-backed families (all threads of the CTA)
__device__ void pd_trial_gradient(const PdLs& A, double s) {
  if (A.kind == CVXB_OBJ_COMPOSED) {
    for (int i = threadIdx.x; i < A.mrows; i += VT) {
      const double xi = A.vin[i] + s * A.dvin[i];
      A.tmp[i] = grad_f0(A.mkind, A.mrows, xi, 0.0, 0.0, A.pw);      // NaN for KL outside its domain: the trial then fails
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < A.n; j += VT / 32) {                       // one warp per column of F, coalesced down the column
      double acc = 0.0;
      for (int i = lane; i < A.mrows; i += 32) acc = fma(A.M[(size_t)j * A.ldM + i], A.tmp[i], acc);
      acc = warp_sum(acc);
      if (lane == 0) A.gf[j] = acc;
    }
  } else {
    for (int i = threadIdx.x; i < A.mcols; i += VT) A.tmp[i] = A.vin[i] * exp(-s * A.dvin[i]);
    __syncthreads();
    for (int j = threadIdx.x; j < A.n; j += VT) {                     // thread j owns row j of B: consecutive threads, consecutive rows
      double acc = 0.0;
      for (int i = 0; i < A.mcols; ++i) acc = fma(A.M[(size_t)i * A.ldM + j], A.tmp[i], acc);
      A.gf[j] = A.a[j] - acc;
    }
  }
  __syncthreads();
}


// ||r(t, u + s du)||^2 and strict feasibility of x_s; lam_s > 0 asserted via *lamneg
__device__ double pd_trial(const PdLs& A, double s, double* buf, int* ibuf, int* feas, int* lamneg, double* rdual2,
                           double* rpri2, double* gap) {
  double sc = 0.0, g_ = 0.0;
  int infeas = 0, ln = 0;
  for (int i = threadIdx.x; i < A.m; i += VT) {
    const double lin = s * A.Gd[i], qd = A.qq ? s * s * A.qq[i] : 0.0;
    double g = A.gx[i] + (lin + qd);
    const double margin = 3.6e-15 * (fabs(A.gx[i]) + fabs(lin) + fabs(qd) + fabs(A.ub[i]));   // see solver.cu: ls_in_set
    if (!(g * IN_SET_FACTOR + margin < A.ub[i])) infeas = 1;
    double f = g - A.ub[i];
    double l = A.lam[i] + s * A.dlam[i];
    if (!(l > 0.0)) ln = 1;
    double rc = -l * f - 1.0 / A.t;
    sc = fma(rc, rc, sc);
    g_ = fma(-f, l, g_);
  }
  sc = block_sum(sc, buf);
  g_ = block_sum(g_, buf);
  infeas = block_or(infeas, ibuf);
  ln = block_or(ln, ibuf);
  double sd = 0.0;
  const bool matrix_backed = A.kind == CVXB_OBJ_COMPOSED || A.kind == CVXB_OBJ_KLDUAL;
  if (matrix_backed) pd_trial_gradient(A, s);
  for (int j = threadIdx.x; j < A.n; j += VT) {
    double gf;
    if (matrix_backed) gf = A.gf[j];
    else if (A.kind == CVXB_OBJ_LINEAR) gf = A.a[j];
    else if (A.kind == CVXB_OBJ_QUADRATIC) gf = A.a[j] + A.Px[j] + s * A.Pd[j];
    else if (A.kind == CVXB_OBJ_PNORM) gf = grad_f0(A.kind, A.n, A.x[j] + s * A.dx[j], 0.0, 0.0, A.pw);
    else {
      double xj = A.x[j] + s * A.dx[j];
      gf = 1.0 + log(xj) + log((double)A.n);        // NaN for x_s <= 0: the comparison below then fails
    }
    double r = gf + A.rd0[j] + s * (A.rd1[j] + (A.rd2 ? s * A.rd2[j] : 0.0));
    sd = fma(r, r, sd);
  }
  sd = block_sum(sd, buf);
  double sp = 0.0;
  if (A.withEqs) {
    for (int i = threadIdx.x; i < A.p; i += VT) {
      double r = A.pres[i] + s * A.Adx[i];
      sp = fma(r, r, sp);
    }
    sp = block_sum(sp, buf);
  }
  *feas = !infeas;
  *lamneg = ln;
  *rdual2 = sd;
  *rpri2 = sp;
  *gap = g_;
  return sd + sc + sp;
}

__global__ void __launch_bounds__(VT) pd_linesearch_kernel(PdLs A, double* scal, int* flag) {
  __shared__ double buf[33];
  __shared__ int ibuf[33];
  const int upstream_bad = flag[F_BAD] | flag[F_PD_NOTNEG] | flag[F_PD_LAMNEG];
  int status = 0, it = 0;
  double s = 0.0, rd2 = 0, rp2 = 0, gap = 0, nrm2 = 0;
  if (!upstream_bad) {
    // s_max keeps lam + s dlam > 0   (:332-339, :502-509)
    double s0 = 1.0;
    for (int i = threadIdx.x; i < A.m; i += VT)
      if (A.dlam[i] < 0.0) s0 = fmin(s0, -A.lam[i] / A.dlam[i]);
    s0 = block_min(s0, buf);
    s = A.frac * s0;
    int feas, ln;
    double n0 = sqrt(pd_trial(A, 0.0, buf, ibuf, &feas, &ln, &rd2, &rp2, &gap));   // ||r_t(u)||
    nrm2 = pd_trial(A, s, buf, ibuf, &feas, &ln, &rd2, &rp2, &gap);
    bool ok = feas && (sqrt(nrm2) < (1.0 - A.alpha * s) * n0);
    const double maxIter = -30.0 / log(A.beta);
    if (ln) status = 4;
    while (!status && !ok && it <= maxIter) {
      s *= A.beta;
      nrm2 = pd_trial(A, s, buf, ibuf, &feas, &ln, &rd2, &rp2, &gap);
      if (ln) { status = 4; break; }
      ok = feas && (sqrt(nrm2) < (1.0 - A.alpha * s) * n0);
      ++it;
    }
    if (!status && it >= maxIter) status = 1;      // LineSearchFailedException (:369-372, :538-541)
    if (!status) {
      for (int i = threadIdx.x; i < A.m; i += VT) A.lamo[i] = A.lam[i] + s * A.dlam[i];
      for (int j = threadIdx.x; j < A.n; j += VT) A.xo[j] = A.x[j] + s * A.dx[j];
      if (A.withEqs)
        for (int i = threadIdx.x; i < A.p; i += VT) A.nuo[i] = A.nu[i] + s * A.dnu[i];
    }
  }
  if (threadIdx.x == 0) {
    flag[F_PD_LS_FAIL] = status;
    flag[F_LS_TRIALS] = it;
    flag[F_STEP_TAKEN] = (!upstream_bad && !status) ? 1 : 0;
    scal[S_STEP] = s;
    scal[S_PD_GAP] = gap;                                     // -g(x).lam at the accepted point
    scal[S_PD_EQGAP] = sqrt(rp2);                             // ||Ax-b||
    scal[S_PD_RNORM] = A.withEqs ? sqrt(nrm2) : sqrt(rd2);    // :600-612 / :440-452
  }
}

// objective value at x
__global__ void __launch_bounds__(VT) pd_objective_kernel(int n, int kind, double obj_r, const double* __restrict__ x,
                                                          const double* __restrict__ a, const double* __restrict__ Px,
                                                          double pw, double* scal, const double* __restrict__ dual_y = nullptr,
                                                          int kd = 0) {
  __shared__ double buf[33];
  double f0 = 0.0;
  if (kind == CVXB_OBJ_KLDUAL)
    for (int j = threadIdx.x; j < kd; j += VT) f0 += dual_y[j];          // + R'exp(-B'z)   (Dist_KL.scala:143-147)
  for (int j = threadIdx.x; j < n; j += VT) {
    double xj = x[j];
    if (kind == CVXB_OBJ_PNORM) f0 += pow(fabs(xj), pw);
    else if (kind == CVXB_OBJ_LINEAR || kind == CVXB_OBJ_KLDUAL) f0 += a[j] * xj;
    else if (kind == CVXB_OBJ_QUADRATIC) f0 += a[j] * xj + 0.5 * xj * Px[j];
    else f0 += xj * log(xj * (double)n);
  }
  f0 = block_sum(f0, buf) + obj_r;
  if (threadIdx.x == 0) scal[S_OBJ] = f0;
}

__global__ void __launch_bounds__(VT) mul_kernel(int n, const double* __restrict__ a, const double* __restrict__ b,
                                                 double* __restrict__ out) {
  for (int i = threadIdx.x; i < n; i += VT) out[i] = a[i] * b[i];
}

template <typename T>
int palloc2(cvxb_problem_s* P, T** ptr, size_t count) {
  void* q = P->arena.take((count ? count : 1) * sizeof(T));
  if (q) {
    CVXB_CUDA_OK(cudaMemsetAsync(q, 0, (count ? count : 1) * sizeof(T), P->h->stream));
    *ptr = (T*)q;
    return CVXB_OK;
  }
  CVXB_CUDA_OK(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
  CVXB_CUDA_OK(cudaMemsetAsync(q, 0, (count ? count : 1) * sizeof(T), P->h->stream));
  P->owned.push_back(q);
  *ptr = (T*)q;
  return CVXB_OK;
}

int pd_alloc(cvxb_problem_s* P) {
  if (P->lam) return CVXB_OK;
  double** mv[] = {&P->lam, &P->dlam, &P->wts, &P->tmpm, &P->lam0s, &P->gx0s};
  for (double** v : mv) CVXB_TRY(palloc2(P, v, (size_t)P->ldm));
  double** nv[] = {&P->rd0, &P->rd1, &P->x0s, &P->rd00s, &P->Px0s, &P->vvec, &P->qvec, &P->atnu};
  for (double** v : nv) CVXB_TRY(palloc2(P, v, (size_t)P->ldn));
  double** pv[] = {&P->dnu, &P->Adx, &P->pres, &P->nu0s, &P->pres0s, &P->negpres};
  for (double** v : pv) CVXB_TRY(palloc2(P, v, (size_t)P->ldp));
  // the n x n scratch of SymmetricLinearSystem.solve is only needed without equality constraints
  if (P->p == 0 && !P->Hreg) CVXB_TRY(palloc2(P, &P->Hreg, (size_t)P->ldn * P->n));
  return CVXB_OK;
}

// everything of one iteration that precedes the linear solve, at (P->x, P->lam, P->nu), parameter t
int pd_assemble(cvxb_problem_s* P, const cvxb_params& pars, double t) {
  Handle& h = *P->h;
  const int n = P->n, m = P->m, p = P->p;
  CVXB_TRY(quad_refresh(P));
  CVXB_TRY(prof_begin(h, PROF_GEMV));
  CVXB_TRY(gemv_n(h, m, n, 1.0, P->G, P->ldm, P->x, 0.0, P->gx));
  CVXB_TRY(prof_end(h, PROF_GEMV, 8.0 * m * n));
  if (P->objective == CVXB_OBJ_QUADRATIC) CVXB_TRY(gemv_n(h, n, n, 1.0, P->obj_P, P->ldn, P->x, 0.0, P->Px));
  if (P->objective == CVXB_OBJ_COMPOSED) CVXB_TRY(composed_refresh(P));
  if (P->objective == CVXB_OBJ_KLDUAL) CVXB_TRY(dual_refresh(P));
  CVXB_LAUNCH(h, pd_cnt_kernel, 1, VT, 0, m, t, P->gr, P->ub, P->gx, P->lam, P->tmpm, P->wts, P->inv, P->qcorr, h.d_scal, h.d_flag);
  CVXB_TRY(prof_begin(h, PROF_GEMV));
  CVXB_TRY(gemv_t(h, m, n, 1.0, P->G, P->ldm, P->inv, 0.0, P->gt));
  CVXB_TRY(prof_end(h, PROF_GEMV, 8.0 * m * n));
  CVXB_TRY(gemv_t(h, m, n, 1.0, P->G, P->ldm, P->lam, 0.0, P->rd0));
  if (p > 0) {
    CVXB_TRY(gemv_t(h, p, n, 1.0, P->A, P->ldp, P->nu, 0.0, P->atnu));
    CVXB_TRY(gemv_n(h, p, n, 1.0, P->A, P->ldp, P->x, 0.0, P->axv));
  }
  CVXB_LAUNCH(h, pd_rhs_kernel, 1, VT, 0, n, p, P->objective, (pars.bugCompat & 1) && p > 0 ? 1 : 0, P->x, P->obj_a, P->Px, P->gt,
              P->atnu, P->rd0, P->vvec, P->qvec, P->axv, P->b, P->pres, P->negpres, P->obj_pow);
  // H_pd = hess f + G' diag(-lam/f) G
  CVXB_TRY(scale_rows(h, m, n, P->G, P->ldm, P->wts, P->Gs, P->ldm, true));
  if (P->objective == CVXB_OBJ_QUADRATIC) CVXB_TRY(fill_matrix(h, n, 1.0, P->obj_P, P->ldn, nullptr, 0.0, P->H, P->ldn));
  else if (P->objective == CVXB_OBJ_KL) CVXB_TRY(fill_matrix(h, n, 0.0, nullptr, 0, P->x, 1.0, P->H, P->ldn));
  else if (P->objective == CVXB_OBJ_PNORM)
    CVXB_TRY(fill_matrix(h, n, 0.0, nullptr, 0, P->x, P->obj_pow * (P->obj_pow - 1.0), P->H, P->ldn, nullptr, P->obj_pow - 2.0));
  else if (P->objective == CVXB_OBJ_COMPOSED) CVXB_TRY(composed_hessian(P, 1.0, nullptr));
  else if (P->objective == CVXB_OBJ_KLDUAL) CVXB_TRY(dual_hessian(P, 1.0, nullptr));
  else CVXB_TRY(fill_matrix(h, n, 0.0, nullptr, 0, nullptr, 0.0, P->H, P->ldn));
  CVXB_TRY(quad_hessian_terms(P, P->lam));     // + lam_k hess g_k   (PrimalDualSolver.scala:230-236)
  GemmArgs g{n, n, m, P->Gs, P->ldm, true, P->Gs, P->ldm, true, P->H, P->ldn, 1.0, 1.0, 2};
  g.streamk = true;
  return gemm_dmma_timed(h, g, (double)m * n * ((double)n + 1.0));
}

// SymmetricLinearSystem(H, v).solve: Ruiz, then choleskySolve(Q, d o v) (which equilibrates again, D6), x = d o u
int pd_symmetric_enqueue(cvxb_problem_s* P, const cvxb_params& pars, bool regularize, bool skip_outer) {
  Handle& h = *P->h;
  const int n = P->n;
  if (!skip_outer) {
    CVXB_TRY(ruiz_equilibrate(h, n, P->H, P->ldn, P->kw.dr2, P->kw.colsq, pars.ruizMaxSweeps, pars.ruizTol, P->kw.L,
                              (size_t)P->kw.ldn * n));
    CVXB_TRY(scaled_full(h, n, P->H, P->ldn, P->kw.dr2, P->Hreg, P->ldn));
    CVXB_LAUNCH(h, mul_kernel, 1, VT, 0, n, P->kw.dr2, P->vvec, P->kw.qk);
  }
  CVXB_TRY(chol_enqueue(h, P->kw, pars, P->Hreg, P->ldn, P->kw.qk, 1.0, pars.tolEqSolve, regularize, false, P->kw.t2));
  CVXB_LAUNCH(h, mul_kernel, 1, VT, 0, n, P->kw.dr2, P->kw.t2, P->dir);
  return CVXB_OK;
}

// symSolve(Q, d o v) then x = d o u   (SymmetricLinearSystem.scala:33, 51-55); Q and d o v are still in place
int pd_symmetric_eigen(cvxb_problem_s* P, const cvxb_params& pars) {
  Handle& h = *P->h;
  CVXB_TRY(svd_solve_device(h, P->n, P->Hreg, P->ldn, P->kw.qk, 1.0, pars.tolEqSolve, P->kw.t2, nullptr, true));
  CVXB_LAUNCH(h, mul_kernel, 1, VT, 0, P->n, P->kw.dr2, P->kw.t2, P->dir);
  return CVXB_OK;
}

int pd_after_solve(cvxb_problem_s* P, const cvxb_params& pars, double t, bool use_base0) {
  Handle& h = *P->h;
  const int n = P->n, m = P->m, p = P->p;
  CVXB_TRY(gemv_n(h, m, n, 1.0, P->G, P->ldm, P->dir, 0.0, P->Gd));
  if (P->objective == CVXB_OBJ_QUADRATIC) CVXB_TRY(gemv_n(h, n, n, 1.0, P->obj_P, P->ldn, P->dir, 0.0, P->Pd));
  if (P->objective == CVXB_OBJ_COMPOSED) CVXB_TRY(gemv_n(h, P->cmp_n, n, 1.0, P->cmpF, P->cmp_ld, P->dir, 0.0, P->cmpd));   // F du
  if (P->objective == CVXB_OBJ_KLDUAL) CVXB_TRY(gemv_t(h, n, P->kd, 1.0, P->obj_P, P->ldn, P->dir, 0.0, P->dv));           // B'dz
  CVXB_LAUNCH(h, pd_dlam_kernel, 1, VT, 0, m, t, P->lam, P->tmpm, P->Gd, P->dlam);
  CVXB_TRY(gemv_t(h, m, n, 1.0, P->G, P->ldm, P->dlam, 0.0, P->rd1));
  if (p > 0) {
    CVXB_TRY(gemv_t(h, p, n, 1.0, P->A, P->ldp, P->dnu, 1.0, P->rd1));
    CVXB_TRY(gemv_n(h, p, n, 1.0, P->A, P->ldp, P->dir, 0.0, P->Adx));
  }
  if (P->mq > 0) {
    // grad g_k(x + s dx) = (a_k + P_k x) + s P_k dx: the dual residual gains  s * sum lam_k P_k dx  +  s^2 * sum dlam_k P_k dx
    CVXB_TRY(quad_direction(P, P->dir));
    if (!P->rd2) CVXB_TRY(palloc2(P, &P->rd2, (size_t)P->ldn));
    CVXB_TRY(gemv_n(h, n, P->mq, 1.0, P->PDv, P->ldq, (use_base0 ? P->lam0s : P->lam) + P->mlin, 1.0, P->rd1));
    CVXB_TRY(gemv_n(h, n, P->mq, 1.0, P->PDv, P->ldq, P->dlam + P->mlin, 0.0, P->rd2));
  }
  PdLs A;
  A.m = m; A.n = n; A.p = p; A.kind = P->objective; A.withEqs = p > 0;
  A.t = t; A.alpha = pars.alpha; A.beta = pars.beta; A.frac = pars.pdStepFraction; A.pw = P->obj_pow;
  if (use_base0) {     // defect D1: the search starts from the initial iterate u0 every time
    A.gx = P->gx0s; A.lam = P->lam0s; A.x = P->x0s; A.nu = P->nu0s; A.rd0 = P->rd00s; A.pres = P->pres0s; A.Px = P->Px0s;
  } else {
    A.gx = P->gx; A.lam = P->lam; A.x = P->x; A.nu = P->nu; A.rd0 = P->rd0; A.pres = P->pres; A.Px = P->Px;
  }
  A.ub = P->ub; A.a = P->obj_a;
  A.Gd = P->Gd; A.dlam = P->dlam; A.dx = P->dir; A.dnu = P->dnu; A.rd1 = P->rd1; A.Adx = P->Adx; A.Pd = P->Pd;
  A.xo = P->x; A.lamo = P->lam; A.nuo = P->nu;
  A.qq = P->mq > 0 ? P->qq : nullptr; A.rd2 = P->mq > 0 ? P->rd2 : nullptr;
  A.M = nullptr; A.vin = A.dvin = nullptr; A.ldM = A.mrows = A.mcols = A.mkind = 0; A.tmp = A.gf = nullptr;
  if (P->objective == CVXB_OBJ_COMPOSED) {
    A.M = P->cmpF; A.ldM = P->cmp_ld; A.mrows = P->cmp_n; A.mcols = n; A.mkind = P->cmp_kind;
    A.vin = P->cmpx; A.dvin = P->cmpd; A.tmp = P->cmpg; A.gf = P->cmpgf;
  } else if (P->objective == CVXB_OBJ_KLDUAL) {
    A.M = P->obj_P; A.ldM = P->ldn; A.mrows = n; A.mcols = P->kd; A.vin = P->dy; A.dvin = P->dv; A.tmp = P->du; A.gf = P->Pd;
  }
  CVXB_LAUNCH(h, pd_linesearch_kernel, 1, VT, 0, A, h.d_scal, h.d_flag);
  return CVXB_OK;
}

int pd_check_flags(Handle& h) {
  if (h.h_flag[F_PD_NOTNEG]) { set_last_error("PrimalDualSolver: assertion failed: constraint values g_i(x)-ub_i not < 0"); return CVXB_ENOTFEASIBLE; }
  if (h.h_flag[F_PD_LAMNEG]) { set_last_error("PrimalDualSolver: assertion failed: lambda not positive"); return CVXB_ELINESEARCH; }
  return CVXB_OK;
}

// one search direction with the host-decided fallbacks; leaves dx in P->dir, dnu in P->dnu
int pd_solve_direction(cvxb_problem_s* P, const cvxb_params& pars, cvxb_kkt_info* info) {
  Handle& h = *P->h;
  if (P->p > 0) {
    CVXB_TRY(kkt_enqueue(h, P->kw, pars, P->H, P->ldn, P->A, P->ldp, P->qvec, P->negpres, pars.tolEqSolve, false, false,
                         P->dir, P->dnu));
    CVXB_TRY(fetch_status(h));
    CVXB_TRY(pd_check_flags(h));
    if (h.h_flag[F_BAD] == 0) { fill_info(h, info, 0, 0); return CVXB_OK; }
    return kkt_solve_fallbacks(h, P->kw, pars, P->H, P->ldn, P->A, P->ldp, P->qvec, P->negpres, pars.tolEqSolve, P->dir,
                               P->dnu, info);
  }
  CVXB_TRY(pd_symmetric_enqueue(P, pars, false, false));
  CVXB_TRY(fetch_status(h));
  CVXB_TRY(pd_check_flags(h));
  if (h.h_flag[F_BAD] == 0) { fill_info(h, info, 0, 0); return CVXB_OK; }
  if (h.h_flag[F_BAD] & 3) {
    CVXB_TRY(pd_symmetric_enqueue(P, pars, true, true));
    CVXB_TRY(fetch_status(h));
    if (h.h_flag[F_BAD] == 0) { fill_info(h, info, 0, 1); return CVXB_OK; }
  }
  fill_info(h, info, 2, 0);
  return pd_symmetric_eigen(P, pars);
}

}  // namespace

int pd_loop(cvxb_problem_s* P, const cvxb_params& pars, cvxb_solution* out) {
  NvtxRange nvtx("cvxb PrimalDualSolver.solve");
  Handle& h = *P->h;
  const int n = P->n, m = P->m, p = P->p;
  const bool withEqs = p > 0;
  const bool bug = (pars.bugCompat & 1) && withEqs;
  const double mu = pars.mu, tol = pars.tolSolver;
  CVXB_TRY(pd_alloc(P));
  // lam0 = -1/(g(x0)-ub), nu0 = 0
  CVXB_TRY(quad_refresh(P));
  CVXB_TRY(gemv_n(h, m, n, 1.0, P->G, P->ldm, P->x, 0.0, P->gx));
  CVXB_LAUNCH(h, pd_lambda0_kernel, 1, VT, 0, m, P->gx, P->gr, P->ub, P->qcorr, P->lam);
  CVXB_CUDA_OK(cudaMemsetAsync(P->nu, 0, (size_t)P->ldp * sizeof(double), h.stream));
  CVXB_CUDA_OK(cudaMemsetAsync(P->dnu, 0, (size_t)P->ldp * sizeof(double), h.stream));
  // surrogate gap at the start (t is irrelevant for it)
  CVXB_LAUNCH(h, pd_cnt_kernel, 1, VT, 0, m, 1.0, P->gr, P->ub, P->gx, P->lam, P->tmpm, P->wts, P->inv, P->qcorr, h.d_scal, h.d_flag);
  CVXB_TRY(fetch_status(h));
  if (h.h_flag[F_PD_NOTNEG]) { set_last_error("PrimalDualSolver: starting point not strictly feasible"); return CVXB_ENOTFEASIBLE; }
  double gap = h.h_scal[S_PD_GAP];
  double eqGap = withEqs ? 1.7976931348623157e308 : 0.0, rnorm = 1.7976931348623157e308;
  double t = mu * m / gap;
  const double maxIter = (withEqs ? 1500.0 : 2000.0) / mu;
  long long limit = pars.stepLimit > 0 ? pars.stepLimit : (long long)1 << 60;
  int it = 0;
  long long trials = 0;
  int fallbacks = 0, regularized = 0;
  if (bug) {   // u0 and everything the line search needs at u0 (defect D1)
    CVXB_CUDA_OK(cudaMemcpyAsync(P->x0s, P->x, n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
    CVXB_CUDA_OK(cudaMemcpyAsync(P->lam0s, P->lam, m * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
    CVXB_CUDA_OK(cudaMemsetAsync(P->nu0s, 0, (size_t)P->ldp * sizeof(double), h.stream));
  }
  while (!(gap < tol && rnorm < tol) && it < maxIter && it < limit) {
    NvtxRange nvtx_step("cvxb primal-dual iteration");
    CVXB_TRY(pd_assemble(P, pars, t));
    if (bug && it == 0) {
      CVXB_CUDA_OK(cudaMemcpyAsync(P->gx0s, P->gx, m * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
      CVXB_CUDA_OK(cudaMemcpyAsync(P->rd00s, P->rd0, n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
      CVXB_CUDA_OK(cudaMemcpyAsync(P->pres0s, P->pres, (size_t)P->ldp * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
      CVXB_CUDA_OK(cudaMemcpyAsync(P->Px0s, P->Px, n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
    }
    // optimistic: direction + line search enqueued back to back, one status read per iteration
    if (withEqs)
      CVXB_TRY(kkt_enqueue(h, P->kw, pars, P->H, P->ldn, P->A, P->ldp, P->qvec, P->negpres, pars.tolEqSolve, false, false,
                           P->dir, P->dnu));
    else
      CVXB_TRY(pd_symmetric_enqueue(P, pars, false, false));
    CVXB_TRY(pd_after_solve(P, pars, t, bug));
    CVXB_TRY(fetch_status(h));
    CVXB_TRY(pd_check_flags(h));
    if (h.h_flag[F_BAD]) {
      cvxb_kkt_info info;
      if (withEqs) {
        CVXB_TRY(kkt_solve_fallbacks(h, P->kw, pars, P->H, P->ldn, P->A, P->ldp, P->qvec, P->negpres, pars.tolEqSolve,
                                     P->dir, P->dnu, &info));
      } else {
        int ok = 0;
        if (h.h_flag[F_BAD] & 3) {
          CVXB_TRY(pd_symmetric_enqueue(P, pars, true, true));
          CVXB_TRY(fetch_status(h));
          ok = h.h_flag[F_BAD] == 0;
          info.path = 0; info.regularized = 1;
        }
        if (!ok) {
          CVXB_TRY(pd_symmetric_eigen(P, pars));
          info.path = 2; info.regularized = 0;
        }
      }
      if (info.path) fallbacks++;
      if (info.regularized) regularized++;
      CVXB_TRY(pd_after_solve(P, pars, t, bug));
      CVXB_TRY(fetch_status(h));
      CVXB_TRY(pd_check_flags(h));
    }
    if (h.h_flag[F_PD_LS_FAIL]) {
      if (h.h_flag[F_PD_LS_FAIL] == 4) set_last_error("PrimalDualSolver line search: assertion failed: lambda_s not positive");
      else set_last_error("PrimalDualSolver: Line search unsuccessful (LineSearchFailedException).");
      return CVXB_ELINESEARCH;
    }
    trials += h.h_flag[F_LS_TRIALS];
    gap = h.h_scal[S_PD_GAP];
    rnorm = h.h_scal[S_PD_RNORM];
    if (withEqs) eqGap = h.h_scal[S_PD_EQGAP];
    t = mu * m / gap;
    it++;
  }
  if (P->objective == CVXB_OBJ_QUADRATIC) CVXB_TRY(gemv_n(h, n, n, 1.0, P->obj_P, P->ldn, P->x, 0.0, P->Px));
  if (P->objective == CVXB_OBJ_COMPOSED) {       // f(u) = f_inner(z0 + F u)
    CVXB_TRY(composed_refresh(P));
    CVXB_LAUNCH(h, pd_objective_kernel, 1, VT, 0, P->cmp_n, P->cmp_kind, P->obj_r, P->cmpx, nullptr, nullptr, P->obj_pow, h.d_scal);
  } else {
    if (P->objective == CVXB_OBJ_KLDUAL) CVXB_TRY(dual_refresh(P));
    CVXB_LAUNCH(h, pd_objective_kernel, 1, VT, 0, n, P->objective, P->obj_r, P->x, P->obj_a, P->Px, P->obj_pow, h.d_scal, P->dy,
                P->kd);
  }
  CVXB_TRY(fetch_status(h));
  out->has_lambda = 1; out->has_nu = withEqs ? 1 : 0;
  out->has_newtonDecrement = 0; out->newtonDecrement = 0;
  out->dualityGap = gap; out->has_dualityGap = 1;
  out->equalityGap = eqGap; out->has_equalityGap = withEqs ? 1 : 0;
  out->has_normGrad = 0; out->normGrad = 0;
  out->normDualResidual = rnorm; out->has_normDualResidual = 1;
  out->iter = it - 1;
  out->maxedOut = (it == (int)maxIter) ? 1 : 0;
  out->objective = h.h_scal[S_OBJ];
  out->outer_stages = it;
  out->newton_steps = it;
  out->executed_newton_steps = it;
  out->linesearch_trials = trials;
  out->kkt_fallbacks = fallbacks;
  out->kkt_regularized = regularized;
  return CVXB_OK;
}

int run_phase1_public(cvxb_problem_s* P, const cvxb_params& pars, long long* budget, bool* limited, cvxb_solution* ph);

}  // namespace cvxb

extern "C" {

int cvxb_pd_solve(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, cvxb_solution* out) {
  if (!h || !prob || !out) { cvxb::set_last_error("cvxb_pd_solve: null argument"); return CVXB_EINVAL; }
  if (prob->h != h) { cvxb::set_last_error("problem belongs to another handle"); return CVXB_EINVAL; }
  cvxb::DeviceGuard guard(h->device);
  cvxb_params dp, lp;
  if (!pars) { cvxb_default_params(&dp); pars = &dp; }
  double *ox = out->x, *ol = out->lambda, *onu = out->nu;
  memset(out, 0, sizeof(*out));
  out->x = ox; out->lambda = ol; out->nu = onu;
  CVXB_CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  if (!prob->has_feasible) {
    long long budget = pars->stepLimit;
    bool limited = pars->stepLimit > 0;
    cvxb_solution ph;
    memset(&ph, 0, sizeof(ph));
    int st = run_phase1_public(prob, *pars, &budget, &limited, &ph);
    out->phase1_newton_steps = ph.newton_steps; out->phase1_executed_steps = ph.executed_newton_steps;
    out->phase1_stages = ph.outer_stages; out->phase1_s = ph.phase1_s;
    if (st != CVXB_OK) return st;
    if (limited) {
      if (budget <= 0 || !prob->has_feasible) {
        // step budget used up inside phase I (no feasible point yet): report the phase-I iterate, as cvxb_barrier_solve does
        cudaEventRecord(h->ev1, h->stream);
        cudaEventSynchronize(h->ev1);
        float ms1 = 0;
        cudaEventElapsedTime(&ms1, h->ev0, h->ev1);
        out->solve_ms = ms1;
        CVXB_TRY(download_vec(*h, out->x, prob->phase1->x, prob->n));
        CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
        return CVXB_OK;
      }
      lp = *pars;
      lp.stepLimit = budget;           // the primal-dual loop gets what phase I left over
      pars = &lp;
    }
  }
  CVXB_CUDA_OK(cudaMemcpyAsync(prob->x, prob->x_feas, prob->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  int st = pd_loop(prob, *pars, out);
  cudaEventRecord(h->ev1, h->stream);
  cudaEventSynchronize(h->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  out->solve_ms = ms;
  if (st != CVXB_OK) return st;
  CVXB_TRY(download_vec(*h, out->x, prob->x, prob->n));
  CVXB_TRY(download_vec(*h, out->lambda, prob->lam, prob->m));
  if (prob->p > 0) CVXB_TRY(download_vec(*h, out->nu, prob->nu, prob->p));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_pd_newton_direction(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, const double* x,
                             const double* lambda, const double* nu, double t, double* H_out, double* dx, double* dlambda,
                             double* dnu, cvxb_kkt_info* info) {
  if (!h || !prob || !x || !lambda) { cvxb::set_last_error("cvxb_pd_newton_direction: null argument"); return CVXB_EINVAL; }
  if (prob->h != h) { cvxb::set_last_error("problem belongs to another handle"); return CVXB_EINVAL; }
  cvxb::DeviceGuard guard(h->device);
  cvxb_params dp;
  if (!pars) { cvxb_default_params(&dp); pars = &dp; }
  cvxb_problem_s* P = prob;
  CVXB_TRY(pd_alloc(P));
  CVXB_TRY(upload_vec(*h, P->x, x, P->n));
  CVXB_TRY(upload_vec(*h, P->lam, lambda, P->m));
  if (P->p > 0) {
    if (nu) CVXB_TRY(upload_vec(*h, P->nu, nu, P->p));
    else CVXB_CUDA_OK(cudaMemsetAsync(P->nu, 0, (size_t)P->ldp * sizeof(double), h->stream));
  }
  CVXB_TRY(pd_assemble(P, *pars, t));
  int st = pd_solve_direction(P, *pars, info);
  if (H_out) {
    bool dev = (h->flags & CVXB_FLAG_DEVICE_PTRS) != 0;
    CVXB_CUDA_OK(cudaMemcpy2DAsync(H_out, (size_t)P->n * sizeof(double), P->H, (size_t)P->ldn * sizeof(double),
                                   (size_t)P->n * sizeof(double), P->n, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                   h->stream));
  }
  if (st == CVXB_OK) {
    Handle& hh = *h;
    CVXB_TRY(gemv_n(hh, P->m, P->n, 1.0, P->G, P->ldm, P->dir, 0.0, P->Gd));
    CVXB_LAUNCH(hh, pd_dlam_kernel, 1, VT, 0, P->m, t, P->lam, P->tmpm, P->Gd, P->dlam);
    CVXB_TRY(download_vec(*h, dx, P->dir, P->n));
    CVXB_TRY(download_vec(*h, dlambda, P->dlam, P->m));
    if (P->p > 0) CVXB_TRY(download_vec(*h, dnu, P->dnu, P->p));
  }
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return st;
}

}  // extern "C"

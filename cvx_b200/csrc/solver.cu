// Seam A: device-resident barrier solves.
//   barrier function family     BarrierSolver.scala:280-315   (value, gradient, Hessian)
//   EqualityConstrainedSolver   EqualityConstrainedSolver.scala:37-107
//   UnconstrainedSolver         UnconstrainedSolver.scala:34-125   (with defect D4 reproduced)
//   BarrierSolver outer loop    BarrierSolver.scala:70-188
//   phase I                     ConstraintSet.scala:131-168,310-395,556-575, Constraint.scala:64-89,
//                               EqualityConstraint.scala:84-100, CvxUtils.scala:61-87
// The iterate never leaves the device.  One Newton step is a fixed kernel sequence (assembly, KKT
// solve, line search, re-evaluation); the host reads one block of status words per step.
//
// Line search in O(m+n) per trial: the reference re-evaluates every constraint object at every trial
// point (O(mn) each); along the ray x + s*d the constraint values are g(x) + s*(G d), so after ONE
// extra GEMV (G d) a whole backtracking search is a loop inside a single CTA.
#include "solver.cuh"
#include <vector>
#include "vecops.cuh"

using namespace cvxb;

namespace cvxb {
namespace {

constexpr double IN_SET_FACTOR = 1.0 + 3e-16;   // Constraint.isSatisfiedStrictly, Constraint.scala:23

// ---- objective families (device) -------------------------------------------------------------------
// value needs a.x / x.Px/2 / sum x log(n x): partial term of entry j
__device__ __forceinline__ double pnorm_sgn(double u) { return fabs(u) < 1e-14 ? 0.0 : (u > 0 ? 1.0 : -1.0); }

__device__ __forceinline__ double obj_term(int kind, int n, int j, double xj, const double* a, const double* Px,
                                           double pw = 2.0) {
  if (kind == CVXB_OBJ_PNORM) return pow(fabs(xj), pw);                          // ObjectiveFunctions.scala:76
  if (kind == CVXB_OBJ_LINEAR || kind == CVXB_OBJ_KLDUAL) return a[j] * xj;    // dual: w'z here, sum_j y_j added by the caller
  if (kind == CVXB_OBJ_QUADRATIC) return a[j] * xj + 0.5 * xj * Px[j];
  return xj * log(xj * (double)n);   // Dist_KL.scala:225-227
}
__device__ __forceinline__ double obj_grad(int kind, int n, int j, double xj, const double* a, const double* Px,
                                           double pw = 2.0) {
  if (kind == CVXB_OBJ_PNORM) { const double sg = pnorm_sgn(xj); return sg * pw * pow(sg * xj, pw - 1.0); }   // :77-78
  if (kind == CVXB_OBJ_COMPOSED) return Px[j];           // F' grad f_inner(z0 + F u), computed by barrier_eval
  if (kind == CVXB_OBJ_LINEAR) return a[j];
  if (kind == CVXB_OBJ_KLDUAL) return a[j] - Px[j];      // w - B y   (Px holds B y)
  if (kind == CVXB_OBJ_QUADRATIC) return a[j] + Px[j];
  return 1.0 + log(xj) + log((double)n);   // Dist_KL.scala:229-233
}

// E3: gx = r + Gx ; slack, 1/slack, log-sum, feasibility ; objective value f0 and barrier value
__global__ void __launch_bounds__(VT) eval_cnt_kernel(int m, int n, int kind, double obj_r, double t,
                                                      const double* __restrict__ t_dev, const double* __restrict__ gr, const double* __restrict__ ub,
                                                      double* __restrict__ gx, double* __restrict__ inv,
                                                      const double* __restrict__ x, const double* __restrict__ a,
                                                      const double* __restrict__ Px, const double* __restrict__ qcorr,
                                                      const double* __restrict__ dual_y, int kd, double pw, double* scal,
                                                      int* flag) {
  __shared__ double buf[33];
  __shared__ int ibuf[33];
  if (t_dev) t = *t_dev;        // barrier parameter kept on the device when the step is replayed from a CUDA graph
  double ls = 0.0, mn = 1e308;
  int bad = 0;
  for (int i = threadIdx.x; i < m; i += VT) {
    // quadratic rows: G_i = a + P x, so G_i.x counts x'Px twice; g = r + a.x + x'Px/2
    double g = (gr ? gr[i] : 0.0) + gx[i] - (qcorr ? qcorr[i] : 0.0);
    gx[i] = g;
    double d = ub[i] - g;
    if (!(d > 0.0)) bad = 1;
    inv[i] = 1.0 / d;
    ls += log(d);
    mn = fmin(mn, d);
  }
  ls = block_sum(ls, buf);
  mn = block_min(mn, buf);
  bad = block_or(bad, ibuf);
  double f0 = 0.0;
  for (int j = threadIdx.x; j < n; j += VT) f0 += obj_term(kind, n, j, x[j], a, Px, pw);
  if (kind == CVXB_OBJ_KLDUAL)
    for (int j = threadIdx.x; j < kd; j += VT) f0 += dual_y[j];      // + R'exp(-B'z)
  f0 = block_sum(f0, buf) + obj_r;
  if (threadIdx.x == 0) {
    scal[S_LOGSUM] = ls;
    scal[S_F0] = f0;
    scal[S_FVAL] = t * f0 - ls;      // BarrierSolver.scala:280-289
    scal[S_MINSLACK] = mn;
    flag[F_INFEAS] = bad;
  }
}

// E6: y = t*grad f0 + G'(1/d) ; ||y|| ; eqdiff = b - Ax ; ||eqdiff||     BarrierSolver.scala:291-301
__global__ void __launch_bounds__(VT) eval_grad_kernel(int n, int p, int kind, double t, const double* __restrict__ t_dev,
                                                       const double* __restrict__ x,
                                                       const double* __restrict__ a, const double* __restrict__ Px,
                                                       const double* __restrict__ gt, double* __restrict__ y,
                                                       const double* __restrict__ b, const double* __restrict__ ax,
                                                       double* __restrict__ eqdiff, double pw, double* scal) {
  __shared__ double buf[33];
  if (t_dev) t = *t_dev;
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += VT) {
    double v = t * obj_grad(kind, n, j, x[j], a, Px, pw) + gt[j];
    y[j] = v;
    s = fma(v, v, s);
  }
  s = block_sum(s, buf);
  double e = 0.0;
  for (int i = threadIdx.x; i < p; i += VT) {
    double v = b[i] - ax[i];
    eqdiff[i] = v;
    e = fma(v, v, e);
  }
  e = block_sum(e, buf);
  if (threadIdx.x == 0) {
    scal[S_NORMGRAD] = sqrt(s);
    scal[S_EQNORM] = sqrt(e);
  }
}

// ---- line search ---------------------------------------------------------------------------------
struct LsArgs {
  int m, n, kind, mode /*0 equality-constrained, 1 unconstrained*/, iter0;
  double t, alpha, beta, tol;
  const double *gx, *ub, *Gd, *a, *Px, *Pd, *y;
  const double* qq;      // d'P_k d / 2 on quadratic rows (NULL without quadratic constraints)
  const double* t_dev;   // non-NULL: t and the first-step flag live on the device (CUDA-graph replay)
  const double *dual_y, *dual_v;   // CVXB_OBJ_KLDUAL: y = R o exp(-B'z), v = B'd ; f(z + s d) = w'z + s w'd + sum y_j exp(-s v_j)
  int kd;
  double pw;             // CVXB_OBJ_PNORM exponent
  double *x, *dir;
  // the vectors the KL / p-norm objective is evaluated on: (x, dir, n), or (z0 + F u, F du, dim x) for a composed objective
  const double *xobj, *dobj;
  int nobj;
};

__device__ bool ls_in_set(const LsArgs& A, double s, int* ibuf) {
  int out = 0;
  for (int i = threadIdx.x; i < A.m; i += VT) {
    const double lin = s * A.Gd[i], qd = A.qq ? s * s * A.qq[i] : 0.0;
    const double g = A.gx[i] + (lin + qd);
    // The ray value g(x) + s G d (+ s^2 d'Pd/2) and a fresh evaluation at x + s d differ by rounding; a few ulps of
    // margin keep the accepted point strictly feasible under BOTH, so the next evaluation can never find a
    // non-positive slack that this test let through (the reference evaluates both with the same function).
    const double margin = 3.6e-15 * (fabs(A.gx[i]) + fabs(lin) + fabs(qd) + fabs(A.ub[i]));
    if (!(g * IN_SET_FACTOR + margin < A.ub[i])) out = 1;
  }
  return block_or(out, ibuf) == 0;
}

// barrier value at x + s*dir; *throws = 1 when some slack <= 0 (IllegalArgumentException in the reference)
__device__ double ls_value(const LsArgs& A, double s, double f0, double c1, double c2, double* buf, int* ibuf,
                           int* throws) {
  double ls = 0.0;
  int bad = 0;
  for (int i = threadIdx.x; i < A.m; i += VT) {
    double d = A.ub[i] - (A.gx[i] + s * (A.Gd[i] + (A.qq ? s * A.qq[i] : 0.0)));
    if (!(d > 0.0)) bad = 1;
    ls += log(d);
  }
  ls = block_sum(ls, buf);
  bad = block_or(bad, ibuf);
  *throws = bad;
  double f0s;
  if (A.kind == CVXB_OBJ_KL) {
    double v = 0.0;
    for (int j = threadIdx.x; j < A.nobj; j += VT) {
      double xj = A.xobj[j] + s * A.dobj[j];
      v += xj * log(xj * (double)A.nobj);
    }
    f0s = block_sum(v, buf);
  } else if (A.kind == CVXB_OBJ_PNORM) {
    double v = 0.0;
    for (int j = threadIdx.x; j < A.nobj; j += VT) v += pow(fabs(A.xobj[j] + s * A.dobj[j]), A.pw);
    f0s = block_sum(v, buf);
  } else if (A.kind == CVXB_OBJ_KLDUAL) {
    double v = 0.0;
    for (int j = threadIdx.x; j < A.kd; j += VT) v += A.dual_y[j] * exp(-s * A.dual_v[j]);
    f0s = f0 + s * c1 + block_sum(v, buf) - c2;     // c1 = w'd, c2 = sum_j y_j (so that f0 + ... replaces the exp term)
  } else {
    f0s = f0 + s * c1 + 0.5 * s * s * c2;   // exact along the ray for linear / quadratic objectives
  }
  return A.t * f0s - ls;
}

__global__ void __launch_bounds__(VT) linesearch_kernel(LsArgs A, double* scal, int* flag) {
  __shared__ double buf[33];
  __shared__ int ibuf[33];
  if (A.t_dev) { A.t = *A.t_dev; A.iter0 = flag[F_ITER0]; }
  // q = d . grad ;  objective line coefficients
  double q = 0.0, c1 = 0.0, c2 = 0.0;
  for (int j = threadIdx.x; j < A.n; j += VT) {
    double dj = A.dir[j];
    q = fma(dj, A.y[j], q);
    if (A.kind == CVXB_OBJ_LINEAR || A.kind == CVXB_OBJ_KLDUAL) c1 = fma(A.a[j], dj, c1);
    else if (A.kind == CVXB_OBJ_QUADRATIC) { c1 = fma(A.a[j] + A.Px[j], dj, c1); c2 = fma(dj, A.Pd[j], c2); }
  }
  if (A.kind == CVXB_OBJ_KLDUAL)
    for (int j = threadIdx.x; j < A.kd; j += VT) c2 += A.dual_y[j];
  q = block_sum(q, buf);
  c1 = block_sum(c1, buf);
  c2 = block_sum(c2, buf);
  const double nd = -q / 2;
  const double f = scal[S_FVAL], f0 = scal[S_F0];
  const int upstream_bad = flag[F_BAD] | flag[F_INFEAS];
  int status = 0, it = 0, taken = 0;
  double step = 0.0;
  if (!upstream_bad && nd > A.tol) {
    if (A.mode == 0) {
      // EqualityConstrainedSolver.scala:79-92 (one counter shared by both loops, defect D7)
      double s = 1.0;
      while (!ls_in_set(A, s, ibuf) && it < 100) { s *= A.beta; ++it; }
      if (it == 100) status = 1;
      else {
        int thr = 0;
        while (it < 100) {
          double v = ls_value(A, s, f0, c1, c2, buf, ibuf, &thr);
          if (thr) { status = 3; break; }
          if (!(v > f + A.alpha * s * q)) break;
          s *= A.beta; ++it;
        }
        if (!status && it == 100) status = 2;
      }
      step = s;
    } else {
      // UnconstrainedSolver.scala:85-115; rho = 1+1/4 == 1 in integer arithmetic (defect D4), so the
      // trust radius stays at its first value; loops bounded by 200 but the failure test is it == 100.
      const double hnorm = sqrt(-q);
      double trust = A.iter0 ? hnorm : scal[S_TRUST];
      const double sc = (A.iter0 || hnorm <= trust) ? 1.0 : trust / hnorm;
      double tt = 1.0;
      while (!ls_in_set(A, sc * tt, ibuf) && it < 200) { tt *= A.beta; ++it; }
      if (it == 100) status = 1;
      else {
        int thr = 0;
        if (ls_in_set(A, sc, ibuf)) {   // else-branch of :100-105 evaluates objF.valueAt(x + s*t)
          (void)ls_value(A, sc * tt, f0, c1, c2, buf, ibuf, &thr);
          if (thr) status = 3;
        }
        while (!status && it < 200) {
          double v = ls_value(A, sc * tt, f0, c1, c2, buf, ibuf, &thr);
          if (thr) { status = 3; break; }
          if (!(v > f + A.alpha * tt * q)) break;
          tt *= A.beta; ++it;
        }
        if (!status && it == 100) status = 2;
      }
      step = sc * tt;
      if (threadIdx.x == 0) { scal[S_TRUST] = trust; scal[S_HNORM] = hnorm; }
    }
    if (!status) {
      for (int j = threadIdx.x; j < A.n; j += VT) A.x[j] = A.x[j] + A.dir[j] * step;
      taken = 1;
    }
  }
  if (threadIdx.x == 0) {
    scal[S_Q] = q;
    scal[S_ND] = nd;
    scal[S_STEP] = step;
    flag[F_LS_STATUS] = status;
    flag[F_LS_TRIALS] = it;
    flag[F_STEP_TAKEN] = taken;
    if (!upstream_bad) flag[F_ITER0] = 0;       // the next step of this stage is not the first any more
  }
}

// ---- dual KL objective ------------------------------------------------------------------------------
// y_j = R_j exp(-u_j), u = B'z        (Dist_KL.scala:143-147, primalOptimum :163)
__global__ void dual_y_kernel(int kd, const double* __restrict__ R, const double* __restrict__ u, double* __restrict__ y) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < kd) y[j] = R[j] * exp(-u[j]);
}
// Bs(:, j) = B(:, j) * sqrt(t y_j): then t * hess = Bs Bs'   (Dist_KL.scala:152-159)
__global__ void dual_scale_cols_kernel(int D, int kd, const double* __restrict__ B, int ldb, const double* __restrict__ y,
                                       double t, const double* __restrict__ t_dev, double* __restrict__ Bs, int ldbs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  if (t_dev) t = *t_dev;
  for (int j = blockIdx.y; j < kd; j += gridDim.y) Bs[(size_t)j * ldbs + i] = B[(size_t)j * ldb + i] * sqrt(t * y[j]);
}

// ---- objective composed with x = z0 + F u (reduced KL / p-norm problems) ----------------------------------
// g = grad f_inner(x) elementwise (Dist_KL.scala:229-233, ObjectiveFunctions.scala:77-78)
__global__ void cmp_grad_kernel(int n, int kind, double pw, const double* __restrict__ x, double* __restrict__ g) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) g[j] = obj_grad(kind, n, j, x[j], nullptr, nullptr, pw);
}
// w = t * diag hess f_inner(x): t / x (Dist_KL.scala:236-239), t p (p-1) |x|^(p-2) (ObjectiveFunctions.scala:79-82)
__global__ void cmp_weights_kernel(int n, int kind, double pw, double t, const double* __restrict__ t_dev,
                                   const double* __restrict__ x, double* __restrict__ w) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (t_dev) t = *t_dev;
  w[j] = kind == CVXB_OBJ_KL ? t / x[j] : t * pw * (pw - 1.0) * pow(fabs(x[j]), pw - 2.0);
}

// ---- quadratic constraints ------------------------------------------------------------------------
// row mlin+k of G := a_k + P_k x (the gradient, QuadraticConstraint.scala:34-36); qcorr := x'P_k x / 2
__global__ void __launch_bounds__(256) quad_rows_kernel(int n, int mlin, int ldq, const double* __restrict__ qa,
                                                        const double* __restrict__ PX, const double* __restrict__ x,
                                                        double* __restrict__ G, int ldg, double* __restrict__ qcorr) {
  __shared__ double buf[33];
  const int k = blockIdx.x;
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += 256) {
    const double px = PX[(size_t)k * ldq + j];
    G[(size_t)j * ldg + mlin + k] = qa[(size_t)k * ldq + j] + px;
    s = fma(x[j], px, s);
  }
  s = block_sum(s, buf);
  if (threadIdx.x == 0) qcorr[mlin + k] = 0.5 * s;
}
// out[mlin+k] := v' (P_k d) / 2
__global__ void __launch_bounds__(256) quad_dot_kernel(int n, int mlin, int ldq, const double* __restrict__ PDv,
                                                       const double* __restrict__ v, double* __restrict__ out) {
  __shared__ double buf[33];
  const int k = blockIdx.x;
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += 256) s = fma(v[j], PDv[(size_t)k * ldq + j], s);
  s = block_sum(s, buf);
  if (threadIdx.x == 0) out[mlin + k] = 0.5 * s;
}
// H += sum_k c[mlin+k] P_k     (hess g_k / d_k in the barrier, lam_k hess g_k in the primal-dual matrix)
__global__ void quad_hess_kernel(int n, int mq, int mlin, int ldq, const double* __restrict__ Pq, int ldP,
                                 const double* __restrict__ c, double* __restrict__ Hm, int ldh) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    double s = 0.0;
    for (int k = 0; k < mq; ++k) s = fma(c[mlin + k], Pq[(size_t)j * ldP + (size_t)k * ldq + i], s);
    Hm[(size_t)j * ldh + i] += s;
  }
}
// phase I: child blocks P1_k = [P_k 0; 0 0], a1_k = [a_k; -1]
__global__ void phase1_quad_kernel(int n, int mq, int ldq, const double* __restrict__ Pq, int ldP,
                                   const double* __restrict__ qa, int ldq1, double* __restrict__ Pq1, int ldP1,
                                   double* __restrict__ qa1) {
  const int k = blockIdx.z;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  for (int j = blockIdx.y; j <= n; j += gridDim.y)
    Pq1[(size_t)j * ldP1 + (size_t)k * ldq1 + i] = (i < n && j < n) ? Pq[(size_t)j * ldP + (size_t)k * ldq + i] : 0.0;
  if (blockIdx.y == 0) qa1[(size_t)k * ldq1 + i] = (i < n) ? qa[(size_t)k * ldq + i] : -1.0;
}

// ---- phase I construction -----------------------------------------------------------------------
// G1 = [G, -1 ; Ae, -1] with Ae rows interleaved (a_i, -a_i), ub1 = [ub ; b_i + tol, -b_i + tol], r1 = [r ; 0]
__global__ void phase1_build_kernel(int n, int m, int p, const double* __restrict__ G, int ldg,
                                    const double* __restrict__ gr, const double* __restrict__ ub,
                                    const double* __restrict__ A, int lda, const double* __restrict__ b, double eqtol,
                                    double* __restrict__ G1, int ldg1, double* __restrict__ gr1,
                                    double* __restrict__ ub1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int m1 = m + 2 * p;
  if (i >= m1) return;
  for (int j = blockIdx.y; j <= n; j += gridDim.y) {
    double v;
    if (j == n) v = -1.0;
    else if (i < m) v = G[(size_t)j * ldg + i];
    else {
      int k = (i - m) >> 1;
      double a = A[(size_t)j * lda + k];
      v = ((i - m) & 1) ? -a : a;
    }
    G1[(size_t)j * ldg1 + i] = v;
  }
  if (blockIdx.y == 0) {
    if (i < m) { gr1[i] = gr ? gr[i] : 0.0; ub1[i] = ub[i]; }
    else {
      int k = (i - m) >> 1;
      gr1[i] = 0.0;
      ub1[i] = ((i - m) & 1) ? (-b[k] + eqtol) : (b[k] + eqtol);
    }
  }
}

// start of phase I: (x0, 1 + max_i (g_i(x0) - ub_i))   ConstraintSet.scala:161-163 ; gx holds G1[:, :n] x0 (s column = 0)
__global__ void __launch_bounds__(VT) phase1_start_kernel(int m1, int n, const double* __restrict__ gx,
                                                          const double* __restrict__ gr1, const double* __restrict__ ub1,
                                                          const double* __restrict__ qcorr, double* __restrict__ x1) {
  __shared__ double buf[33];
  double mx = -1e308;
  for (int i = threadIdx.x; i < m1; i += VT) mx = fmax(mx, (gr1[i] + gx[i] - (qcorr ? qcorr[i] : 0.0)) - ub1[i]);
  mx = -block_min(-mx, buf);
  if (threadIdx.x == 0) x1[n] = 1.0 + mx;
}

__global__ void set_stage_kernel(double t, double* scal, int* flag) {
  scal[S_T] = t;
  flag[F_ITER0] = 1;
}

// ---- device-driven centering stage ----------------------------------------------------------------------------------
struct LoopSeed {
  int iter, budget, limited, maxIter, mode;
  double nd, normGrad, eqGap, tol;
};
__global__ void loop_seed_kernel(LoopSeed v, double* scal, int* flag) {
  flag[F_L_ITER] = v.iter; flag[F_L_EXEC] = 0; flag[F_L_TRIALS] = 0; flag[F_L_BUDGET] = v.budget; flag[F_L_LIMITED] = v.limited;
  flag[F_L_MAXITER] = v.maxIter; flag[F_L_MODE] = v.mode; flag[F_L_REASON] = 0;
  scal[S_L_ND] = v.nd; scal[S_L_NORMGRAD] = v.normGrad; scal[S_L_EQGAP] = v.eqGap; scal[S_L_TOL] = v.tol;
}
enum LoopReason { LOOP_DONE = 0, LOOP_NEEDS_HOST = 1, LOOP_LS_FAILED = 2, LOOP_INFEASIBLE = 3, LOOP_BUDGET = 4, LOOP_SPIN = 5 };
// After one Newton step: the bookkeeping and the loop test of EqualityConstrainedSolver.solve (:49, mode 0) /
// UnconstrainedSolver.solve (:45, mode 1), exactly as inner_solve_eq / inner_solve_uncon apply them on the host.
// Anything the device cannot finish by itself (a refused linear solve: the host walks the fallback chain; a failed
// line search; an infeasible iterate) stops the loop with the step left uncounted.
__global__ void loop_decide_kernel(cudaGraphConditionalHandle handle, double* scal, int* flag) {
  const double tol = scal[S_L_TOL];
  int go = 0, reason = LOOP_DONE;
  if (flag[F_BAD]) reason = LOOP_NEEDS_HOST;
  else if (flag[F_LS_STATUS]) reason = LOOP_LS_FAILED;
  else if (flag[F_INFEAS]) reason = LOOP_INFEASIBLE;
  else {
    const int mode = flag[F_L_MODE];
    const double nd = scal[S_ND];
    double normGrad = scal[S_L_NORMGRAD], eqGap = scal[S_L_EQGAP];
    int iter = flag[F_L_ITER] + 1;
    flag[F_L_EXEC] += 1;
    if (flag[F_L_LIMITED]) flag[F_L_BUDGET] -= 1;
    bool spin = false;
    if (flag[F_STEP_TAKEN]) {
      flag[F_L_TRIALS] += flag[F_LS_TRIALS];
      normGrad = scal[S_NORMGRAD];
      if (mode == 0) eqGap = scal[S_EQNORM];
    } else if (mode == 0 && ((nd > tol && normGrad > tol) || eqGap > tol) && !(nd > tol)) {
      // no step taken, so x, H and d repeat exactly: the reference spins here until maxIter (inner_solve_eq)
      iter = flag[F_L_MAXITER];
      spin = true;
    }
    const bool cond = mode == 0 ? ((nd > tol && normGrad > tol) || eqGap > tol) : (nd > tol && normGrad > tol);
    const bool budget_out = flag[F_L_LIMITED] && flag[F_L_BUDGET] <= 0;
    go = (!spin && iter < flag[F_L_MAXITER] && cond && !budget_out) ? 1 : 0;
    if (spin) reason = LOOP_SPIN;
    else if (!go && budget_out && cond && iter < flag[F_L_MAXITER]) reason = LOOP_BUDGET;
    flag[F_L_ITER] = iter;
    scal[S_L_ND] = nd; scal[S_L_NORMGRAD] = normGrad; scal[S_L_EQGAP] = eqGap;
  }
  flag[F_L_REASON] = reason;
  cudaGraphSetConditional(handle, go ? 1u : 0u);
}

__global__ void set_unit_kernel(int n, int k, double* a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (i == k) ? 1.0 : 0.0;
}

template <typename T>
int palloc(cvxb_problem_s* P, T** ptr, size_t count) {
  void* q = P->arena.take((count ? count : 1) * sizeof(T));     // arena memory is zeroed once at creation
  if (q) { *ptr = (T*)q; return CVXB_OK; }
  CVXB_CUDA_OK(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
  CVXB_CUDA_OK(cudaMemsetAsync(q, 0, (count ? count : 1) * sizeof(T), P->h->stream));
  P->owned.push_back(q);
  *ptr = (T*)q;
  return CVXB_OK;
}

}  // namespace

// --------------------------------------------------------------------------------- problem objects
// m = number of LINEAR constraints; the problem carries m + mq rows in every per-constraint vector
int problem_alloc(Handle& h, int n, int m, int p, int objective, cvxb_problem_s** out, int mq, int kd) {
  cvxb_problem_s* P = new cvxb_problem_s();
  P->h = &h;
  P->mlin = m; P->mq = mq; P->ldq = pad_ld(n);
  m += mq;
  P->n = n; P->m = m; P->p = p; P->objective = objective;
  P->ldm = pad_ld(m); P->ldn = pad_ld(n); P->ldp = pad_ld(p);
  int st = CVXB_OK;
  {   // one allocation for everything below (+ the KKT workspace)
    const size_t ldm = P->ldm, ldn = P->ldn, ldp = P->ldp;
    size_t d = 2 * ldm * n + ldn * n * (objective == CVXB_OBJ_QUADRATIC ? 2 : 1) + ldp * n + 7 * ldm + 10 * ldn + 4 * ldp;
    if (mq > 0) d += (size_t)mq * P->ldq * (n + 3) + 4 * 32;
    if (objective == CVXB_OBJ_KLDUAL) d += 2 * ldn * (size_t)kd + 4 * (size_t)pad_ld(kd) + 6 * 32;
    d += 6 * ldm + 9 * ldn + 6 * ldp;      // primal-dual work vectors (pd_alloc): no allocation inside the first PD solve
    size_t bytes = d * sizeof(double) + 96 * 256 + kkt_work_bytes(n, p);
    void* base = nullptr;
    // stream-ordered pool of the handle's device (release threshold = keep everything): the second problem of a
    // similar size re-uses the first one's memory instead of paying a fresh cudaMalloc of hundreds of MB
    if (cudaMallocAsync(&base, bytes, h.stream) == cudaSuccess) {
      P->arena_async = true;
      P->owned.push_back(base);
      P->arena.base = (char*)base; P->arena.size = bytes; P->arena.used = 0;
      if (cudaMemsetAsync(base, 0, bytes, h.stream) != cudaSuccess) st = CVXB_ECUDA;
    } else {
      cudaGetLastError();   // fall back to per-buffer allocations
    }
  }
  auto A = [&](double** ptr, size_t c) { if (st == CVXB_OK) st = palloc(P, ptr, c); };
  A(&P->G, (size_t)P->ldm * n); A(&P->gr, P->ldm); A(&P->ub, P->ldm);
  A(&P->A, (size_t)P->ldp * n); A(&P->b, P->ldp);
  A(&P->obj_a, P->ldn);
  if (objective == CVXB_OBJ_QUADRATIC) A(&P->obj_P, (size_t)P->ldn * n);
  A(&P->x_feas, P->ldn); A(&P->x_def, P->ldn);
  A(&P->x, P->ldn); A(&P->gx, P->ldm); A(&P->inv, P->ldm); A(&P->Gd, P->ldm); A(&P->y, P->ldn); A(&P->gt, P->ldn);
  A(&P->dir, P->ldn); A(&P->nu, P->ldp); A(&P->eqdiff, P->ldp); A(&P->Px, P->ldn); A(&P->Pd, P->ldn); A(&P->axv, P->ldp);
  A(&P->Gs, (size_t)P->ldm * n); A(&P->H, (size_t)P->ldn * n);
  if (objective == CVXB_OBJ_KLDUAL) {
    P->kd = kd; P->ldk = pad_ld(kd);
    A(&P->obj_P, (size_t)P->ldn * kd); A(&P->Bs, (size_t)P->ldn * kd);
    A(&P->objR, P->ldk); A(&P->du, P->ldk); A(&P->dy, P->ldk); A(&P->dv, P->ldk);
  }
  if (mq > 0) {
    A(&P->Pq, (size_t)mq * P->ldq * n); A(&P->qa, (size_t)mq * P->ldq); A(&P->PX, (size_t)mq * P->ldq);
    A(&P->PDv, (size_t)mq * P->ldq); A(&P->qcorr, P->ldm); A(&P->qq, P->ldm);
  }
  if (st == CVXB_OK) st = kkt_work_alloc(h, P->kw, n, p, &P->arena);
  if (st != CVXB_OK) {
    for (size_t i = 0; i < P->owned.size(); ++i) {
      if (i == 0 && P->arena_async) cudaFreeAsync(P->owned[i], h.stream); else cudaFree(P->owned[i]);
    }
    kkt_work_free(P->kw);
    delete P;
    return st;
  }
  *out = P;
  return CVXB_OK;
}

void problem_free(cvxb_problem_s* P) {
  if (!P) return;
  if (P->phase1) problem_free(P->phase1);
  for (int k = 0; k < 2; ++k) {
    if (P->step_graph[k]) cudaGraphExecDestroy(P->step_graph[k]);
    if (P->loop_graph[k]) cudaGraphExecDestroy(P->loop_graph[k]);
  }
  for (size_t i = 0; i < P->owned.size(); ++i) {
    if (i == 0 && P->arena_async) cudaFreeAsync(P->owned[i], P->h->stream); else cudaFree(P->owned[i]);
  }
  kkt_work_free(P->kw);
  delete P;
}

// ------------------------------------------------------------------------------- barrier function
// value / gradient / constraint state at P->x for barrier parameter t (E1-E6); no host sync
// quadratic constraints at P->x: PX = P_k x, gradient rows of G, value corrections
int quad_refresh(cvxb_problem_s* P) {
  if (P->mq <= 0) return CVXB_OK;
  Handle& h = *P->h;
  CVXB_TRY(gemv_n(h, P->mq * P->ldq, P->n, 1.0, P->Pq, P->mq * P->ldq, P->x, 0.0, P->PX));
  CVXB_LAUNCH(h, quad_rows_kernel, P->mq, 256, 0, P->n, P->mlin, P->ldq, P->qa, P->PX, P->x, P->G, P->ldm, P->qcorr);
  return CVXB_OK;
}
// d'P_k d / 2 for the direction `dir` -> P->qq (line searches); PDv keeps P_k d
int quad_direction(cvxb_problem_s* P, const double* dir) {
  if (P->mq <= 0) return CVXB_OK;
  Handle& h = *P->h;
  CVXB_TRY(gemv_n(h, P->mq * P->ldq, P->n, 1.0, P->Pq, P->mq * P->ldq, dir, 0.0, P->PDv));
  CVXB_LAUNCH(h, quad_dot_kernel, P->mq, 256, 0, P->n, P->mlin, P->ldq, P->PDv, dir, P->qq);
  return CVXB_OK;
}
// H += sum_k c_k P_k
int quad_hessian_terms(cvxb_problem_s* P, const double* c) {
  if (P->mq <= 0) return CVXB_OK;
  Handle& h = *P->h;
  const int n = P->n;
  CVXB_LAUNCH(h, quad_hess_kernel, dim3((n + 127) / 128, n > 1024 ? 1024 : n), 128, 0, n, P->mq, P->mlin, P->ldq, P->Pq,
              P->mq * P->ldq, c, P->H, P->ldn);
  return CVXB_OK;
}

// composed objective at P->x: x_full = z0 + F u, g = grad f_inner(x_full), Px := F'g   (ObjectiveFunction.scala:26-40)
int composed_refresh(cvxb_problem_s* P) {
  Handle& h = *P->h;
  CVXB_CUDA_OK(cudaMemcpyAsync(P->cmpx, P->cmpz0, (size_t)P->cmp_n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
  CVXB_TRY(gemv_n(h, P->cmp_n, P->n, 1.0, P->cmpF, P->cmp_ld, P->x, 1.0, P->cmpx));
  CVXB_LAUNCH(h, cmp_grad_kernel, (P->cmp_n + 255) / 256, 256, 0, P->cmp_n, P->cmp_kind, P->obj_pow, P->cmpx, P->cmpg);
  CVXB_TRY(gemv_t(h, P->cmp_n, P->n, 1.0, P->cmpF, P->cmp_ld, P->cmpg, 0.0, P->Px));
  return CVXB_OK;
}
// H := t F' diag(hess f_inner(x_full)) F as one weighted SYRK, mirrored (exactly symmetric)
int composed_hessian(cvxb_problem_s* P, double t, const double* t_dev) {
  Handle& h = *P->h;
  CVXB_LAUNCH(h, cmp_weights_kernel, (P->cmp_n + 255) / 256, 256, 0, P->cmp_n, P->cmp_kind, P->obj_pow, t, t_dev, P->cmpx, P->cmpw);
  CVXB_TRY(scale_rows(h, P->cmp_n, P->n, P->cmpF, P->cmp_ld, P->cmpw, P->cmpFs, P->cmp_ld, true));
  GemmArgs g{P->n, P->n, P->cmp_n, P->cmpFs, P->cmp_ld, true, P->cmpFs, P->cmp_ld, true, P->H, P->ldn, 1.0, 0.0, 2};
  g.streamk = true;
  return gemm_dmma(h, g);
}

// dual KL objective at P->x = z: u = B'z, y = R o exp(-u), Px := B y   (Dist_KL.scala:143-151)
int dual_refresh(cvxb_problem_s* P) {
  Handle& h = *P->h;
  CVXB_TRY(gemv_t(h, P->n, P->kd, 1.0, P->obj_P, P->ldn, P->x, 0.0, P->du));
  CVXB_LAUNCH(h, dual_y_kernel, (P->kd + 255) / 256, 256, 0, P->kd, P->objR, P->du, P->dy);
  return gemv_n(h, P->n, P->kd, 1.0, P->obj_P, P->ldn, P->dy, 0.0, P->Px);
}
// H := t * B diag(y) B' as one more weighted SYRK (contraction length = the primal dimension), mirrored: exactly
// symmetric   (Dist_KL.scala:152-159)
int dual_hessian(cvxb_problem_s* P, double t, const double* t_dev) {
  Handle& h = *P->h;
  const int n = P->n;
  CVXB_LAUNCH(h, dual_scale_cols_kernel, dim3((n + 127) / 128, P->kd > 1024 ? 1024 : P->kd), 128, 0, n, P->kd, P->obj_P,
              P->ldn, P->dy, t, t_dev, P->Bs, P->ldn);
  GemmArgs gd{n, n, P->kd, P->Bs, P->ldn, false, P->Bs, P->ldn, false, P->H, P->ldn, 1.0, 0.0, 2};
  gd.streamk = true;
  return gemm_dmma(h, gd);
}

int barrier_eval(cvxb_problem_s* P, double t, const double* t_dev = nullptr) {
  Handle& h = *P->h;
  const int n = P->n, m = P->m, p = P->p;
  CVXB_TRY(quad_refresh(P));
  CVXB_TRY(gemv_n(h, m, n, 1.0, P->G, P->ldm, P->x, 0.0, P->gx));
  if (P->objective == CVXB_OBJ_QUADRATIC) CVXB_TRY(gemv_n(h, n, n, 1.0, P->obj_P, P->ldn, P->x, 0.0, P->Px));
  if (P->objective == CVXB_OBJ_KLDUAL) CVXB_TRY(dual_refresh(P));           // u = B'z, y = R o exp(-u), Px := B y
  if (P->objective == CVXB_OBJ_COMPOSED) CVXB_TRY(composed_refresh(P));     // x_full = z0 + F u, Px := F' grad f(x_full)
  const bool cmp = P->objective == CVXB_OBJ_COMPOSED;
  CVXB_LAUNCH(h, eval_cnt_kernel, 1, VT, 0, m, cmp ? P->cmp_n : n, cmp ? P->cmp_kind : P->objective, P->obj_r, t, t_dev, P->gr,
              P->ub, P->gx, P->inv, cmp ? P->cmpx : P->x, P->obj_a, P->Px, P->qcorr, P->dy, P->kd, P->obj_pow, h.d_scal,
              h.d_flag);
  CVXB_TRY(gemv_t(h, m, n, 1.0, P->G, P->ldm, P->inv, 0.0, P->gt));
  if (p > 0) CVXB_TRY(gemv_n(h, p, n, 1.0, P->A, P->ldp, P->x, 0.0, P->axv));
  CVXB_LAUNCH(h, eval_grad_kernel, 1, VT, 0, n, p, P->objective, t, t_dev, P->x, P->obj_a, P->Px, P->gt, P->y, P->b, P->axv,
              P->eqdiff, P->obj_pow, h.d_scal);
  return CVXB_OK;
}

// H = t*hess f0 + G' diag(1/d^2) G     BarrierSolver.scala:303-315 (weighted outer-product sum as one SYRK)
int barrier_hessian(cvxb_problem_s* P, double t, const double* t_dev = nullptr) {
  Handle& h = *P->h;
  const int n = P->n, m = P->m;
  const double tv = t_dev ? 1.0 : t;      // with t_dev the kernel multiplies by *t_dev itself
  CVXB_TRY(scale_rows(h, m, n, P->G, P->ldm, P->inv, P->Gs, P->ldm, false));
  if (P->objective == CVXB_OBJ_QUADRATIC) CVXB_TRY(fill_matrix(h, n, tv, P->obj_P, P->ldn, nullptr, 0.0, P->H, P->ldn, t_dev));
  else if (P->objective == CVXB_OBJ_KL) CVXB_TRY(fill_matrix(h, n, 0.0, nullptr, 0, P->x, tv, P->H, P->ldn, t_dev));
  else if (P->objective == CVXB_OBJ_PNORM)      // t p (p-1) |x|^(p-2) on the diagonal
    CVXB_TRY(fill_matrix(h, n, 0.0, nullptr, 0, P->x, tv * P->obj_pow * (P->obj_pow - 1.0), P->H, P->ldn, t_dev, P->obj_pow - 2.0));
  else if (P->objective == CVXB_OBJ_KLDUAL) CVXB_TRY(dual_hessian(P, t, t_dev));
  else if (P->objective == CVXB_OBJ_COMPOSED) CVXB_TRY(composed_hessian(P, t, t_dev));
  else CVXB_TRY(fill_matrix(h, n, 0.0, nullptr, 0, nullptr, 0.0, P->H, P->ldn));
  CVXB_TRY(quad_hessian_terms(P, P->inv));     // + hess g_k / d_k   (BarrierSolver.scala:313)
  GemmArgs g{n, n, m, P->Gs, P->ldm, true, P->Gs, P->ldm, true, P->H, P->ldn, 1.0, 1.0, 2};
  g.streamk = true;
  return gemm_dmma_timed(h, g, (double)m * n * ((double)n + 1.0));   // lower triangle, mul + add
}

int enqueue_linesearch(cvxb_problem_s* P, const cvxb_params& pars, double t, int mode, int iter0,
                       const double* t_dev = nullptr) {
  Handle& h = *P->h;
  const int n = P->n, m = P->m;
  CVXB_TRY(gemv_n(h, m, n, 1.0, P->G, P->ldm, P->dir, 0.0, P->Gd));
  if (P->objective == CVXB_OBJ_QUADRATIC) CVXB_TRY(gemv_n(h, n, n, 1.0, P->obj_P, P->ldn, P->dir, 0.0, P->Pd));
  CVXB_TRY(quad_direction(P, P->dir));
  if (P->objective == CVXB_OBJ_KLDUAL) CVXB_TRY(gemv_t(h, n, P->kd, 1.0, P->obj_P, P->ldn, P->dir, 0.0, P->dv));
  const bool cmp = P->objective == CVXB_OBJ_COMPOSED;
  if (cmp) CVXB_TRY(gemv_n(h, P->cmp_n, n, 1.0, P->cmpF, P->cmp_ld, P->dir, 0.0, P->cmpd));     // F du
  LsArgs A;
  A.m = m; A.n = n; A.kind = cmp ? P->cmp_kind : P->objective; A.mode = mode; A.iter0 = iter0;
  A.xobj = cmp ? P->cmpx : P->x; A.dobj = cmp ? P->cmpd : P->dir; A.nobj = cmp ? P->cmp_n : n;
  A.t = t; A.alpha = pars.alpha; A.beta = pars.beta; A.tol = pars.tolSolver;
  A.gx = P->gx; A.ub = P->ub; A.Gd = P->Gd; A.a = P->obj_a; A.Px = P->Px; A.Pd = P->Pd; A.y = P->y;
  A.x = P->x; A.dir = P->dir; A.qq = P->mq > 0 ? P->qq : nullptr; A.t_dev = t_dev;
  A.dual_y = P->dy; A.dual_v = P->dv; A.kd = P->kd; A.pw = P->obj_pow;
  CVXB_LAUNCH(h, linesearch_kernel, 1, VT, 0, A, h.d_scal, h.d_flag);
  return CVXB_OK;
}

// One optimistic Newton step (assembly, linear solve, line search, re-evaluation) as a CUDA graph: the launch
// sequence depends only on the problem's shape, so it is captured once per problem and mode and replayed
// with t / first-step flag read from device memory.  ~150 kernel launches become one graph launch.
int enqueue_step(cvxb_problem_s* P, const cvxb_params& pars, double t, int mode, int iter0, const double* t_dev) {
  Handle& h = *P->h;
  CVXB_TRY(barrier_hessian(P, t, t_dev));
  if (mode == 0)
    CVXB_TRY(kkt_enqueue(h, P->kw, pars, P->H, P->ldn, P->A, P->ldp, P->y, P->eqdiff, pars.tolEqSolve, false, false,
                         P->dir, P->nu));
  else
    CVXB_TRY(chol_enqueue(h, P->kw, pars, P->H, P->ldn, P->y, -1.0, pars.tolEqSolve, false, false, P->dir));
  CVXB_TRY(enqueue_linesearch(P, pars, t, mode, iter0, t_dev));
  return barrier_eval(P, t, t_dev);
}

// the parameters baked into the captured kernels' arguments
static bool same_step_pars(const cvxb_params& a, const cvxb_params& b) {
  return a.alpha == b.alpha && a.beta == b.beta && a.tolSolver == b.tolSolver && a.tolEqSolve == b.tolEqSolve &&
         a.ruizMaxSweeps == b.ruizMaxSweeps && a.ruizTol == b.ruizTol && a.cholRegDelta == b.cholRegDelta &&
         a.cholMinDiag == b.cholMinDiag;
}

int run_step(cvxb_problem_s* P, const cvxb_params& pars, double t, int mode, int iter0) {
  NvtxRange nvtx("cvxb Newton step");
  Handle& h = *P->h;
  if (!h.use_graphs) return enqueue_step(P, pars, t, mode, iter0, nullptr);
  cudaGraphExec_t& exec = P->step_graph[mode];
  if (!exec || !same_step_pars(P->graph_pars, pars)) {
    if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
    if (P->step_graph[1 - mode]) { cudaGraphExecDestroy(P->step_graph[1 - mode]); P->step_graph[1 - mode] = nullptr; }
    if (!same_step_pars(P->graph_pars, pars))
      for (int k = 0; k < 2; ++k)
        if (P->loop_graph[k]) { cudaGraphExecDestroy(P->loop_graph[k]); P->loop_graph[k] = nullptr; }
    P->graph_pars = pars;
    const long long l0 = h.launches;
    cudaGraph_t graph = nullptr;
    CVXB_CUDA_OK(cudaStreamBeginCapture(h.stream, cudaStreamCaptureModeThreadLocal));
    h.capturing = true;
    h.capture_flops = 0.0;
    int st = enqueue_step(P, pars, t, mode, iter0, h.d_scal + S_T);
    h.capturing = false;
    P->graph_flops[mode] = h.capture_flops;
    cudaError_t e = cudaStreamEndCapture(h.stream, &graph);
    if (st != CVXB_OK || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      h.launches = l0;
      h.use_graphs = 0;                    // capture not possible here: fall back to plain launches for good
      return enqueue_step(P, pars, t, mode, iter0, nullptr);
    }
    P->graph_launches[mode] = h.launches - l0;
    h.launches = l0;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      cudaGetLastError();
      exec = nullptr;
      h.use_graphs = 0;
      return enqueue_step(P, pars, t, mode, iter0, nullptr);
    }
  }
  CVXB_CUDA_OK(cudaGraphLaunch(exec, h.stream));
  h.launches += P->graph_launches[mode];
  return CVXB_OK;
}

// after the step's status read: fold the SYRK's external event pair into the profile
static void graph_profile_tick(Handle& h, double flops) {
  if (!h.use_graphs || !h.prof_on) return;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, h.gev0, h.gev1) == cudaSuccess) {
    h.prof_ms_graph += ms;
    h.prof_launches_graph++;
    h.prof_flops += flops;
  } else {
    cudaGetLastError();
  }
}

// The whole centering stage as one graph launch.  Built once per problem and mode: WHILE node, body = the captured
// Newton step + loop_decide_kernel.  Returns false when the device-driven loop is not available here (then the
// caller drives the stage step by step as before).
static bool stage_loop_build(cvxb_problem_s* P, const cvxb_params& pars, double t, int mode) {
  Handle& h = *P->h;
  if (!h.use_loop || !h.use_graphs || P->loop_failed) return false;
  if (P->loop_graph[mode] && same_step_pars(P->graph_pars, pars)) return true;
  if (!same_step_pars(P->graph_pars, pars)) {
    for (int k = 0; k < 2; ++k) {
      if (P->loop_graph[k]) { cudaGraphExecDestroy(P->loop_graph[k]); P->loop_graph[k] = nullptr; }
      if (P->step_graph[k]) { cudaGraphExecDestroy(P->step_graph[k]); P->step_graph[k] = nullptr; }
    }
    P->graph_pars = pars;
  }
  cudaGraph_t graph = nullptr;
  auto fail = [&]() {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    P->loop_failed = true;
    return false;
  };
  if (cudaGraphCreate(&graph, 0) != cudaSuccess) return fail();
  cudaGraphConditionalHandle handle;
  if (cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault) != cudaSuccess) return fail();
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  if (cudaGraphAddNode(&node, graph, nullptr, 0, &np) != cudaSuccess) return fail();
  cudaGraph_t body = np.conditional.phGraph_out[0];
  const long long l0 = h.launches;
  if (cudaStreamBeginCaptureToGraph(h.stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fail();
  h.capturing = true;
  h.capture_plain = true;
  h.capture_flops = 0.0;
  int st = enqueue_step(P, pars, t, mode, 0, h.d_scal + S_T);
  P->loop_flops[mode] = h.capture_flops;
  if (st == CVXB_OK) {
    loop_decide_kernel<<<1, 1, 0, h.stream>>>(handle, h.d_scal, h.d_flag);
    h.launches++;
    if (cudaGetLastError() != cudaSuccess) st = CVXB_ECUDA;
  }
  h.capturing = false;
  h.capture_plain = false;
  cudaGraph_t out = nullptr;
  cudaError_t e = cudaStreamEndCapture(h.stream, &out);
  P->loop_step_launches[mode] = h.launches - l0;
  h.launches = l0;
  if (st != CVXB_OK || e != cudaSuccess) return fail();
  if (cudaGraphInstantiate(&P->loop_graph[mode], graph, 0) != cudaSuccess) { P->loop_graph[mode] = nullptr; return fail(); }
  cudaGraphDestroy(graph);
  return true;
}

struct InnerResult {
  double nd = 0, normGrad = 0, eqGap = 0;
  int iter = 0;
  bool maxedOut = false;
  long long executed = 0, trials = 0;
  int fallbacks = 0, regularized = 0;
};

struct RunStats {
  long long budget = 0;       // remaining Newton steps when stepLimit > 0
  bool limited = false;
};

static int ls_status_to_error(Handle& h, const char* who) {
  int s = h.h_flag[F_LS_STATUS];
  if (s == 0) return CVXB_OK;
  if (s == 3) {
    set_last_error("%s: barrierFunction: x not strictly feasible at a line-search trial point "
                   "(IllegalArgumentException, BarrierSolver.scala:284)", who);
    return CVXB_ENOTFEASIBLE;
  }
  set_last_error("%s: Line search: %s", who, s == 1 ? "backtracking into the set C failed."
                                                    : "sufficient decrease not reached after 100 iterations.");
  return CVXB_ELINESEARCH;
}

// One launch of the device-driven stage loop from the state in R / rs; updates them from the loop's state block.
// *reason = why the loop stopped (LoopReason).
static int stage_loop_run(cvxb_problem_s* P, const cvxb_params& pars, int mode, RunStats& rs, InnerResult& R, int* reason) {
  Handle& h = *P->h;
  LoopSeed seed;
  seed.iter = R.iter;
  seed.budget = (int)(rs.budget > 2000000000ll ? 2000000000ll : rs.budget);
  seed.limited = rs.limited ? 1 : 0;
  seed.maxIter = pars.maxIter;
  seed.mode = mode;
  seed.nd = R.nd; seed.normGrad = R.normGrad; seed.eqGap = R.eqGap; seed.tol = pars.tolSolver;
  CVXB_LAUNCH(h, loop_seed_kernel, 1, 1, 0, seed, h.d_scal, h.d_flag);
  CVXB_CUDA_OK(cudaGraphLaunch(P->loop_graph[mode], h.stream));
  CVXB_TRY(fetch_status(h));
  *reason = h.h_flag[F_L_REASON];
  const int exec = h.h_flag[F_L_EXEC];
  const bool aborted = *reason == LOOP_NEEDS_HOST || *reason == LOOP_LS_FAILED || *reason == LOOP_INFEASIBLE;
  h.launches += (long long)(exec + (aborted ? 1 : 0)) * P->loop_step_launches[mode];
  h.prof_flops += (double)(exec + (aborted ? 1 : 0)) * P->loop_flops[mode];      // SYRKs stamped inside the loop
  R.iter = h.h_flag[F_L_ITER];
  R.executed += exec;
  R.trials += h.h_flag[F_L_TRIALS];
  if (rs.limited) rs.budget -= exec;
  R.nd = h.h_scal[S_L_ND];
  R.normGrad = h.h_scal[S_L_NORMGRAD];
  if (mode == 0) R.eqGap = h.h_scal[S_L_EQGAP];
  return CVXB_OK;
}

// EqualityConstrainedSolver.solve (EqualityConstrainedSolver.scala:37-107) at barrier parameter t, from P->x.
// The stage runs as ONE graph launch (WHILE node, loop test on the device): the host reads the status block once when
// the loop ends and only takes part in a step when the device refused its linear solve (fallback chain of
// KKTSystem.solve).  Without that facility (CVXB_NO_LOOP, profiling on, graph build refused) it drives the stage step
// by step: one graph launch and one status read per Newton step.
int inner_solve_eq(cvxb_problem_s* P, const cvxb_params& pars, double t, RunStats& rs, InnerResult& R) {
  NvtxRange nvtx("cvxb EqualityConstrainedSolver.solve (one barrier stage)");
  Handle& h = *P->h;
  const double tol = pars.tolSolver;
  R = InnerResult();
  R.nd = tol + 1;
  CVXB_LAUNCH(h, set_stage_kernel, 1, 1, 0, t, h.d_scal, h.d_flag);
  CVXB_TRY(barrier_eval(P, t));
  CVXB_TRY(fetch_status(h));
  if (h.h_flag[F_INFEAS]) { set_last_error("gradientBarrierFunction: x not strictly feasible"); return CVXB_ENOTFEASIBLE; }
  R.normGrad = h.h_scal[S_NORMGRAD];
  R.eqGap = h.h_scal[S_EQNORM];
  bool pending = false;      // a step whose status block is already on the host (the device loop stopped inside it)
  while (R.iter < pars.maxIter && ((R.nd > tol && R.normGrad > tol) || R.eqGap > tol)) {
    if (rs.limited && rs.budget <= 0) break;
    if (!pending && stage_loop_build(P, pars, t, 0)) {
      int reason = LOOP_DONE;
      CVXB_TRY(stage_loop_run(P, pars, 0, rs, R, &reason));
      if (reason == LOOP_LS_FAILED) return ls_status_to_error(h, "EqualityConstrainedSolver");
      if (reason == LOOP_INFEASIBLE) { set_last_error("gradientBarrierFunction: x not strictly feasible"); return CVXB_ENOTFEASIBLE; }
      if (reason != LOOP_NEEDS_HOST) continue;       // the loop test above decides (done / budget / spin: iter == maxIter)
      pending = true;
    }
    if (!pending) {
      CVXB_TRY(run_step(P, pars, t, 0, 0));
      CVXB_TRY(fetch_status(h));
      graph_profile_tick(h, P->graph_flops[0]);
    }
    pending = false;
    if (h.h_flag[F_BAD]) {
      // optimistic attempt refused on the device (x untouched): walk the reference's fallback chain
      cvxb_kkt_info info;
      CVXB_TRY(kkt_solve_fallbacks(h, P->kw, pars, P->H, P->ldn, P->A, P->ldp, P->y, P->eqdiff, pars.tolEqSolve,
                                   P->dir, P->nu, &info));
      if (info.path) R.fallbacks++;
      if (info.regularized) R.regularized++;
      CVXB_TRY(enqueue_linesearch(P, pars, t, 0, 0));
      CVXB_TRY(barrier_eval(P, t));
      CVXB_TRY(fetch_status(h));
    }
    CVXB_TRY(ls_status_to_error(h, "EqualityConstrainedSolver"));
    if (h.h_flag[F_INFEAS]) { set_last_error("gradientBarrierFunction: x not strictly feasible"); return CVXB_ENOTFEASIBLE; }
    R.nd = h.h_scal[S_ND];
    R.executed++;
    if (rs.limited) rs.budget--;
    R.iter++;
    if (h.h_flag[F_STEP_TAKEN]) {
      R.trials += h.h_flag[F_LS_TRIALS];
      R.normGrad = h.h_scal[S_NORMGRAD];
      R.eqGap = h.h_scal[S_EQNORM];
    } else if ((R.nd > tol && R.normGrad > tol) || R.eqGap > tol) {
      // No step was taken, so x, H and d repeat exactly: the reference spins here until maxIter
      // (its loop condition keeps ||b-Ax|| > tol alive).  Same result, without re-running identical steps.
      if (!(R.nd > tol)) { R.iter = pars.maxIter; break; }
    }
  }
  R.maxedOut = R.iter >= pars.maxIter;
  return CVXB_OK;
}

// UnconstrainedSolver.solve (UnconstrainedSolver.scala:34-125); device-driven like inner_solve_eq
int inner_solve_uncon(cvxb_problem_s* P, const cvxb_params& pars, double t, RunStats& rs, InnerResult& R) {
  NvtxRange nvtx("cvxb UnconstrainedSolver.solve (one barrier stage)");
  Handle& h = *P->h;
  const double tol = pars.tolSolver;
  const int n = P->n;
  R = InnerResult();
  R.nd = tol + 1;
  CVXB_LAUNCH(h, set_stage_kernel, 1, 1, 0, t, h.d_scal, h.d_flag);
  CVXB_TRY(barrier_eval(P, t));
  CVXB_TRY(fetch_status(h));
  if (h.h_flag[F_INFEAS]) { set_last_error("gradientBarrierFunction: x not strictly feasible"); return CVXB_ENOTFEASIBLE; }
  R.normGrad = h.h_scal[S_NORMGRAD];
  bool pending = false;
  while (R.iter < pars.maxIter && R.nd > tol && R.normGrad > tol) {
    if (rs.limited && rs.budget <= 0) break;
    const bool first = R.iter == 0;
    if (!pending && stage_loop_build(P, pars, t, 1)) {
      int reason = LOOP_DONE;
      CVXB_TRY(stage_loop_run(P, pars, 1, rs, R, &reason));
      if (reason == LOOP_LS_FAILED) return ls_status_to_error(h, "UnconstrainedSolver");
      if (reason == LOOP_INFEASIBLE) { set_last_error("gradientBarrierFunction: x not strictly feasible"); return CVXB_ENOTFEASIBLE; }
      if (reason != LOOP_NEEDS_HOST) continue;
      pending = true;
    }
    const bool iter0 = pending ? (R.iter == 0) : first;
    if (!pending) {
      CVXB_TRY(run_step(P, pars, t, 1, iter0));
      CVXB_TRY(fetch_status(h));
      graph_profile_tick(h, P->graph_flops[1]);
    }
    pending = false;
    if (h.h_flag[F_BAD]) {
      cvxb_kkt_info info;
      int st = chol_solve_retry(h, P->kw, pars, P->H, P->ldn, P->y, -1.0, pars.tolEqSolve, P->dir, &info);
      if (st == CVXB_ELINSOLVE) {
        // choleskySolve(H + 1e-9 I, -y)    UnconstrainedSolver.scala:58-61
        if (!P->Hreg) CVXB_TRY(palloc(P, &P->Hreg, (size_t)P->ldn * n));
        CVXB_TRY(copy_matrix(h, n, n, P->H, P->ldn, P->Hreg, P->ldn));
        CVXB_TRY(add_diag(h, n, pars.newtonRegDelta, P->Hreg, P->ldn));
        st = chol_solve_device(h, P->kw, pars, P->Hreg, P->ldn, P->y, -1.0, pars.tolEqSolve, P->dir, &info);
        R.fallbacks++;
        if (st == CVXB_ELINSOLVE)    // MatrixUtils.symSolve(H, -y)   UnconstrainedSolver.scala:65
          st = svd_solve_device(h, n, P->H, P->ldn, P->y, -1.0, pars.tolEqSolve, P->dir, nullptr, true);
      }
      if (st != CVXB_OK) return st;
      if (info.regularized) R.regularized++;
      CVXB_TRY(enqueue_linesearch(P, pars, t, 1, iter0));
      CVXB_TRY(barrier_eval(P, t));
      CVXB_TRY(fetch_status(h));
    }
    CVXB_TRY(ls_status_to_error(h, "UnconstrainedSolver"));
    if (h.h_flag[F_INFEAS]) { set_last_error("gradientBarrierFunction: x not strictly feasible"); return CVXB_ENOTFEASIBLE; }
    R.nd = h.h_scal[S_ND];
    R.executed++;
    if (rs.limited) rs.budget--;
    R.iter++;
    if (h.h_flag[F_STEP_TAKEN]) {
      R.trials += h.h_flag[F_LS_TRIALS];
      R.normGrad = h.h_scal[S_NORMGRAD];
    }
  }
  R.maxedOut = R.iter >= pars.maxIter;
  return CVXB_OK;
}

enum Termination { TERM_STANDARD = 0, TERM_PHASE1 = 1 };

// BarrierSolver.solveWithEQs / solveWithoutEQs (BarrierSolver.scala:70-177), from P->x
int barrier_loop(cvxb_problem_s* P, const cvxb_params& pars, int term, RunStats& rs, cvxb_solution* out) {
  NvtxRange nvtx(term == TERM_PHASE1 ? "cvxb phase I barrier loop" : "cvxb BarrierSolver.solve");
  Handle& h = *P->h;
  const double mu = pars.mu;
  double t = pars.t0;
  double dualityGap = 1.7976931348623157e308, equalityGap = 1.7976931348623157e308, objValue = 1.7976931348623157e308;
  const bool withEqs = P->p > 0;
  const double maxIter = 1000.0 / mu;
  int it = 0;
  InnerResult R;
  long long steps = 0, executed = 0, trials = 0;
  int fallbacks = 0, regularized = 0;
  auto terminated = [&]() {
    if (term == TERM_STANDARD) return dualityGap < pars.tolSolver && equalityGap < pars.tolSolver;
    return objValue < 0.0 && equalityGap < pars.phase1EqTol;    // CvxUtils.scala:78-87
  };
  // first check uses the MaxValue state, never true
  while (!terminated() && it < maxIter) {
    int st = withEqs ? inner_solve_eq(P, pars, t, rs, R) : inner_solve_uncon(P, pars, t, rs, R);
    if (st != CVXB_OK) return st;
    if (it < 128) out->stage_newton_steps[it] = R.iter;
    steps += R.iter; executed += R.executed; trials += R.trials;
    fallbacks += R.fallbacks; regularized += R.regularized;
    objValue = h.h_scal[S_F0];             // objF.valueAt(x) at the stage's final iterate
    dualityGap = (double)P->m / t;
    equalityGap = withEqs ? R.eqGap : 0.0;
    t = mu * t;
    it++;
    if (rs.limited && rs.budget <= 0) break;
  }
  out->newtonDecrement = R.nd; out->has_newtonDecrement = 1;
  out->dualityGap = dualityGap; out->has_dualityGap = 1;
  out->equalityGap = withEqs ? equalityGap : 0.0; out->has_equalityGap = withEqs ? 1 : 0;
  out->normGrad = R.normGrad; out->has_normGrad = 1;
  out->has_normDualResidual = 0; out->normDualResidual = 0;
  out->has_lambda = out->has_nu = 0;
  out->iter = R.iter;
  out->maxedOut = R.maxedOut ? 1 : 0;
  out->objective = objValue;
  out->outer_stages = it;
  out->newton_steps = steps;
  out->executed_newton_steps = executed;
  out->linesearch_trials = trials;
  out->kkt_fallbacks = fallbacks;
  out->kkt_regularized = regularized;
  return CVXB_OK;
}

// ConstraintSet.withFeasiblePoint -> phase_I_Analysis (ConstraintSet.scala:326-395, 556-575)
int run_phase1(cvxb_problem_s* P, const cvxb_params& pars, RunStats& rs, cvxb_solution* ph_out) {
  Handle& h = *P->h;
  const int n = P->n, m = P->mlin, p = P->p, mq = P->mq;
  const int m1 = m + 2 * p;             // linear rows of the feasibility problem; its quadratic rows follow
  if (!P->phase1) {
    CVXB_TRY(problem_alloc(h, n + 1, m1, 0, CVXB_OBJ_LINEAR, &P->phase1, mq, 0));
    cvxb_problem_s* Q = P->phase1;
    dim3 grid((m1 + 127) / 128, n + 1 > 1024 ? 1024 : n + 1);
    if (m1 > 0)      // a problem may have quadratic constraints only (joptP1, SimpleOptimizationProblems.scala:347-377)
      CVXB_LAUNCH(h, phase1_build_kernel, grid, 128, 0, n, m, p, P->G, P->ldm, P->gr, P->ub, P->A, P->ldp, P->b,
                  pars.phase1EqTol, Q->G, Q->ldm, Q->gr, Q->ub);
    if (mq > 0) {     // Constraint.phase_I of a quadratic constraint: P1 = [P 0; 0 0], a1 = [a; -1]  (Constraint.scala:64-89)
      CVXB_LAUNCH(h, phase1_quad_kernel, dim3((n + 1 + 127) / 128, n + 1 > 1024 ? 1024 : n + 1, mq), 128, 0, n, mq, P->ldq,
                  P->Pq, mq * P->ldq, P->qa, Q->ldq, Q->Pq, mq * Q->ldq, Q->qa);
      CVXB_CUDA_OK(cudaMemcpyAsync(Q->gr + m1, P->gr + m, mq * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
      CVXB_CUDA_OK(cudaMemcpyAsync(Q->ub + m1, P->ub + m, mq * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
    }
    CVXB_LAUNCH(h, set_unit_kernel, (n + 1 + 255) / 256, 256, 0, n + 1, n, Q->obj_a);   // f(x,s) = s  (:131-144)
    Q->obj_r = 0.0;
  }
  cvxb_problem_s* Q = P->phase1;
  // start (pointWhereDefined, 1 + max(g - ub))
  CVXB_CUDA_OK(cudaMemsetAsync(Q->x, 0, (size_t)Q->ldn * sizeof(double), h.stream));
  CVXB_CUDA_OK(cudaMemcpyAsync(Q->x, P->x_def, n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
  CVXB_TRY(quad_refresh(Q));
  CVXB_TRY(gemv_n(h, Q->m, n + 1, 1.0, Q->G, Q->ldm, Q->x, 0.0, Q->gx));
  CVXB_LAUNCH(h, phase1_start_kernel, 1, VT, 0, Q->m, n, Q->gx, Q->gr, Q->ub, Q->qcorr, Q->x);
  cvxb_solution sol;
  memset(&sol, 0, sizeof(sol));
  int st = barrier_loop(Q, pars, TERM_PHASE1, rs, &sol);
  if (ph_out) {
    double *x = ph_out->x, *l = ph_out->lambda, *nu = ph_out->nu;
    *ph_out = sol;
    ph_out->x = x; ph_out->lambda = l; ph_out->nu = nu;
  }
  if (st != CVXB_OK) return st;
  double s_feas = 0;
  CVXB_CUDA_OK(cudaMemcpyAsync(&s_feas, Q->x + n, sizeof(double), cudaMemcpyDeviceToHost, h.stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(h.stream));
  if (ph_out) ph_out->phase1_s = s_feas;
  if (rs.limited && rs.budget <= 0) return CVXB_OK;     // step-limited benchmark run: stop here
  if (!(s_feas < pars.tolSolver)) {   // FeasibilityReport.isFeasible(tol)
    set_last_error("Problem not feasible within tolerance %g: phase I slack s = %g (InfeasibleProblemException)",
                   pars.tolSolver, s_feas);
    return CVXB_EINFEASIBLE;
  }
  CVXB_CUDA_OK(cudaMemcpyAsync(P->x_feas, Q->x, n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
  P->has_feasible = true;
  return CVXB_OK;
}

int run_phase1_public(cvxb_problem_s* P, const cvxb_params& pars, long long* budget, bool* limited, cvxb_solution* ph) {
  RunStats rs;
  rs.budget = *budget;
  rs.limited = *limited;
  int st = run_phase1(P, pars, rs, ph);
  *budget = rs.budget;
  return st;
}

int upload_vec(Handle& h, double* dst, const double* src, int n) {
  if (!src || n <= 0) return CVXB_OK;
  bool dev = (h.flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  CVXB_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               h.stream));
  return CVXB_OK;
}
int upload_mat(Handle& h, double* dst, int ldd, const double* src, int lds, int rows, int cols) {
  if (!src || rows <= 0 || cols <= 0) return CVXB_OK;
  bool dev = (h.flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  CVXB_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)ldd * sizeof(double), src, (size_t)lds * sizeof(double),
                                 (size_t)rows * sizeof(double), cols, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                 h.stream));
  return CVXB_OK;
}
int download_vec(Handle& h, double* dst, const double* src, int n) {
  if (!dst || n <= 0) return CVXB_OK;
  bool dev = (h.flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  CVXB_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                               h.stream));
  return CVXB_OK;
}

}  // namespace cvxb

#define CHECK_HP(h, prob)                                                                        \
  if (!(h) || !(prob)) { cvxb::set_last_error("null handle or problem"); return CVXB_EINVAL; }  \
  if ((prob)->h != (h)) { cvxb::set_last_error("problem belongs to another handle"); return CVXB_EINVAL; } \
  cvxb::DeviceGuard _guard((h)->device)

extern "C" {

int cvxb_problem_create(cvxb_handle h, const cvxb_problem_desc* d, cvxb_problem* out) {
  if (!h || !d || !out) { cvxb::set_last_error("cvxb_problem_create: null argument"); return CVXB_EINVAL; }
  if (d->n < 1 || d->m < 0 || d->mq < 0 || d->p < 0) {
    cvxb::set_last_error("cvxb_problem_create: need n >= 1, m >= 0, p >= 0 (got %d, %d, %d)", d->n, d->m, d->p);
    return CVXB_EDIM;
  }
  if (d->objective == CVXB_OBJ_PNORM && !(d->obj_pow >= 2.0)) {
    cvxb::set_last_error("p-norm needs p >= 2 but p = %g", d->obj_pow);       // assert of ObjectiveFunctions.scala:72
    return CVXB_EINVAL;
  }
  if (d->objective < CVXB_OBJ_LINEAR || d->objective > CVXB_OBJ_PNORM) { cvxb::set_last_error("unknown objective kind"); return CVXB_EINVAL; }
  if (d->objective == CVXB_OBJ_KLDUAL && (d->obj_k < 1 || !d->obj_P || !d->obj_R || !d->obj_a || d->obj_ldP < d->n)) {
    cvxb::set_last_error("cvxb_problem_create: the dual KL objective needs obj_a = w, obj_P = B (n x obj_k), obj_R, obj_k >= 1");
    return CVXB_EINVAL;
  }
  if ((d->m > 0 && (!d->G || !d->ub)) || (d->p > 0 && (!d->A || !d->b)) || (!d->x_feasible && !d->x_defined) ||
      (d->objective != CVXB_OBJ_KL && d->objective != CVXB_OBJ_PNORM && !d->obj_a) || (d->objective == CVXB_OBJ_QUADRATIC && !d->obj_P)) {
    cvxb::set_last_error("cvxb_problem_create: missing array for this problem family");
    return CVXB_EINVAL;
  }
  if (d->ldg < d->m || (d->p > 0 && d->lda < d->p) || (d->objective == CVXB_OBJ_QUADRATIC && d->obj_ldP < d->n)) {
    cvxb::set_last_error("cvxb_problem_create: leading dimension too small");
    return CVXB_EDIM;
  }
  cvxb::DeviceGuard _guard(h->device);
  if (d->mq < 0 || (d->mq > 0 && (!d->q_P || !d->q_a || !d->q_r || !d->q_ub))) {
    cvxb::set_last_error("cvxb_problem_create: quadratic constraints need q_P, q_a, q_r, q_ub");
    return CVXB_EINVAL;
  }
  cvxb_problem_s* P = nullptr;
  CVXB_TRY(problem_alloc(*h, d->n, d->m, d->p, d->objective, &P, d->mq, d->objective == CVXB_OBJ_KLDUAL ? d->obj_k : 0));
  int st = CVXB_OK;
  auto T = [&](int s) { if (st == CVXB_OK) st = s; };
  T(upload_mat(*h, P->G, P->ldm, d->G, d->ldg, d->m, d->n));
  T(upload_vec(*h, P->gr, d->g_r, d->m));
  T(upload_vec(*h, P->ub, d->ub, d->m));
  T(upload_mat(*h, P->A, P->ldp, d->A, d->lda, d->p, d->n));
  T(upload_vec(*h, P->b, d->b, d->p));
  T(upload_vec(*h, P->obj_a, d->obj_a, d->n));
  if (d->objective == CVXB_OBJ_QUADRATIC) T(upload_mat(*h, P->obj_P, P->ldn, d->obj_P, d->obj_ldP, d->n, d->n));
  if (d->objective == CVXB_OBJ_KLDUAL) {
    T(upload_mat(*h, P->obj_P, P->ldn, d->obj_P, d->obj_ldP, d->n, d->obj_k));
    T(upload_vec(*h, P->objR, d->obj_R, d->obj_k));
  }
  for (int k = 0; k < d->mq; ++k)       // stacked blocks: rows k*ldq.. of an (mq*ldq) x n matrix
    T(upload_mat(*h, P->Pq + (size_t)k * P->ldq, d->mq * P->ldq, d->q_P + (size_t)k * d->n * d->n, d->n, d->n, d->n));
  if (d->mq > 0) {
    T(upload_mat(*h, P->qa, P->ldq, d->q_a, d->n, d->n, d->mq));
    T(upload_vec(*h, P->gr + d->m, d->q_r, d->mq));
    T(upload_vec(*h, P->ub + d->m, d->q_ub, d->mq));
  }
  P->obj_r = d->obj_r;
  P->obj_pow = d->objective == CVXB_OBJ_PNORM ? d->obj_pow : 2.0;
  if (d->x_feasible) { T(upload_vec(*h, P->x_feas, d->x_feasible, d->n)); P->has_feasible = true; }
  T(upload_vec(*h, P->x_def, d->x_defined ? d->x_defined : d->x_feasible, d->n));
  if (st == CVXB_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) { cvxb::set_last_error("upload failed"); st = CVXB_ECUDA; }
  if (st != CVXB_OK) { problem_free(P); return st; }
  *out = P;
  return CVXB_OK;
}

// ---------------------------------------------------------------------------- equality elimination
// BarrierSolver.reduced / affineTransformed (BarrierSolver.scala:209-256) for the closed-form families: with
// x = z0 + F u the linear constraints become (G F) u <= ub - r - G z0 (LinearConstraint.affineTransformed,
// LinearConstraint.scala:46-52), quadratic constraints and the quadratic objective get P -> F'PF,
// a -> F'(a + P z0), r -> r + a'z0 + z0'Pz0/2 (QuadraticConstraint.scala:54-64, ObjectiveFunction.scala:26-40).
// The reference transforms the completed barrier function (F' H_x F per step, 2n^2k + nk^2 extra flops per Newton
// step); building the reduced data once gives the same iterates from one m x k SYRK per step.
__global__ void __launch_bounds__(VT) affine_shift_kernel(int n, const double* __restrict__ a, const double* __restrict__ Pz,
                                                          const double* __restrict__ z0, double* __restrict__ out,
                                                          double* __restrict__ r_out) {
  // out = a + P z0 ;  *r_out = z0 . (a + P z0 / 2)
  __shared__ double buf[33];
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += VT) {
    const double aj = a ? a[j] : 0.0, pz = Pz ? Pz[j] : 0.0;
    out[j] = aj + pz;
    s = fma(z0[j], aj + 0.5 * pz, s);
  }
  s = block_sum(s, buf);
  if (threadIdx.x == 0) *r_out = s;
}

static int problem_reduce(cvxb_problem_s* P, cvxb::SolutionSpaceDev* S, double tol, cvxb_problem_s** out) {
  Handle& h = *P->h;
  const int n = P->n, k = S->k(), ml = P->mlin, mq = P->mq;
  if (S->n != n) { cvxb::set_last_error("reduced: dimension mismatch F.rows=%d not equal to dim(problem)=%d", S->n, n); return CVXB_EDIM; }
  if (P->p > 0) {
    cvxb::set_last_error("reduced: the solver still carries equality constraints (the reference would keep A F u = b - A z0, "
                         "a 0 = 0 system); build the solver without them");
    return CVXB_ENOTIMPL;
  }
  const bool compose = P->objective == CVXB_OBJ_KL || P->objective == CVXB_OBJ_PNORM;
  if (P->objective != CVXB_OBJ_LINEAR && P->objective != CVXB_OBJ_QUADRATIC && !compose) {
    cvxb::set_last_error("reduced: the affine transform of this objective family (kind %d) is not built on the device", P->objective);
    return CVXB_ENOTIMPL;
  }
  cvxb_problem_s* R = nullptr;
  CVXB_TRY(problem_alloc(h, k, ml, 0, compose ? (int)CVXB_OBJ_COMPOSED : P->objective, &R, mq, 0));
  if (compose) {
    // f(u) = f_inner(z0 + F u): keep F and z0 with the reduced problem (ObjectiveFunction.affineTransformed,
    // ObjectiveFunction.scala:26-40); value, gradient F'grad f and Hessian F' hess f F are evaluated per step
    R->cmp_n = n; R->cmp_ld = pad_ld(n); R->cmp_kind = P->objective;
    const size_t ldc = (size_t)R->cmp_ld;
    double* blk = nullptr;
    cudaError_t e = cudaMalloc((void**)&blk, sizeof(double) * (2 * ldc * k + 5 * ldc + (size_t)pad_ld(k)));
    if (e != cudaSuccess) { cvxb::set_last_error("CUDA error %s in reduced", cudaGetErrorString(e)); problem_free(R); return CVXB_ECUDA; }
    R->owned.push_back(blk);
    cudaMemsetAsync(blk, 0, sizeof(double) * (2 * ldc * k + 5 * ldc + (size_t)pad_ld(k)), h.stream);
    R->cmpF = blk; R->cmpFs = blk + ldc * k;
    double* v = blk + 2 * ldc * k;
    R->cmpz0 = v; R->cmpx = v + ldc; R->cmpd = v + 2 * ldc; R->cmpg = v + 3 * ldc; R->cmpw = v + 4 * ldc; R->cmpgf = v + 5 * ldc;
    cudaMemcpy2DAsync(R->cmpF, ldc * sizeof(double), S->F(), (size_t)S->ldq * sizeof(double), (size_t)n * sizeof(double), k,
                      cudaMemcpyDeviceToDevice, h.stream);
    cudaMemcpyAsync(R->cmpz0, S->z0, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream);
  }
  double* T1 = nullptr;        // n x k scratch, then vectors
  double* vec = nullptr;       // 2 * ldn + (mq + 2) scalars
  int st = CVXB_OK;
  auto T = [&](int s) { if (st == CVXB_OK) st = s; };
  auto CU = [&](cudaError_t e) { if (st == CVXB_OK && e != cudaSuccess) { cvxb::set_last_error("CUDA error %s in reduced", cudaGetErrorString(e)); st = CVXB_ECUDA; } };
  const int ldt = pad_ld(n);
  const bool need_T1 = P->objective == CVXB_OBJ_QUADRATIC || mq > 0;
  if (need_T1) CU(cudaMalloc((void**)&T1, sizeof(double) * (size_t)ldt * k));
  CU(cudaMalloc((void**)&vec, sizeof(double) * (2 * (size_t)ldt + mq + 8)));
  double *Pz = vec, *shifted = vec + ldt, *scal = vec + 2 * ldt;      // scal[0] objective, scal[1 + kq] quadratic constraint kq
  const double* F = S->F();
  if (st == CVXB_OK) {
    CU(cudaMemsetAsync(vec, 0, sizeof(double) * (2 * (size_t)ldt + mq + 8), h.stream));
    if (ml > 0) {
      GemmArgs g{ml, k, n, P->G, P->ldm, false, F, S->ldq, true, R->G, R->ldm, 1.0, 0.0, 0};
      T(gemm_dmma(h, g));
      CU(cudaMemcpyAsync(R->gr, P->gr, sizeof(double) * ml, cudaMemcpyDeviceToDevice, h.stream));
      T(gemv_n(h, ml, n, 1.0, P->G, P->ldm, S->z0, 1.0, R->gr));
    }
    CU(cudaMemcpyAsync(R->ub, P->ub, sizeof(double) * (ml + mq), cudaMemcpyDeviceToDevice, h.stream));
    // objective
    if (P->objective == CVXB_OBJ_QUADRATIC) {
      GemmArgs g1{n, k, n, P->obj_P, P->ldn, false, F, S->ldq, true, T1, ldt, 1.0, 0.0, 0};
      T(gemm_dmma(h, g1));
      GemmArgs g2{k, k, n, F, S->ldq, true, T1, ldt, true, R->obj_P, R->ldn, 1.0, 0.0, 2};
      T(gemm_dmma(h, g2));
      T(gemv_n(h, n, n, 1.0, P->obj_P, P->ldn, S->z0, 0.0, Pz));
    }
    if (st == CVXB_OK && !compose) {
      CVXB_LAUNCH(h, affine_shift_kernel, 1, VT, 0, n, P->obj_a, P->objective == CVXB_OBJ_QUADRATIC ? Pz : nullptr, S->z0, shifted, scal);
      T(gemv_t(h, n, k, 1.0, F, S->ldq, shifted, 0.0, R->obj_a));
    }
    // quadratic constraints
    for (int q = 0; q < mq && st == CVXB_OK; ++q) {
      const double* Pk = P->Pq + (size_t)q * P->ldq;
      const int ldpk = mq * P->ldq;
      GemmArgs g1{n, k, n, Pk, ldpk, false, F, S->ldq, true, T1, ldt, 1.0, 0.0, 0};
      T(gemm_dmma(h, g1));
      GemmArgs g2{k, k, n, F, S->ldq, true, T1, ldt, true, R->Pq + (size_t)q * R->ldq, mq * R->ldq, 1.0, 0.0, 2};
      T(gemm_dmma(h, g2));
      T(gemv_n(h, n, n, 1.0, Pk, ldpk, S->z0, 0.0, Pz));
      if (st != CVXB_OK) break;
      CVXB_LAUNCH(h, affine_shift_kernel, 1, VT, 0, n, P->qa + (size_t)q * P->ldq, Pz, S->z0, shifted, scal + 1 + q);
      T(gemv_t(h, n, k, 1.0, F, S->ldq, shifted, 0.0, R->qa + (size_t)q * R->ldq));
    }
    // starting points: u = F'(x - z0)      (SolutionSpace.parameter)
    T(cvxb::solution_space_parameter(h, S, P->x_def, R->x_def));
    if (P->has_feasible) T(cvxb::solution_space_parameter(h, S, P->x_feas, R->x_feas));
  }
  std::vector<double> hs((size_t)mq + 2, 0.0), hq((size_t)mq + 1, 0.0);
  double dist = 0.0;
  if (st == CVXB_OK) {
    CU(cudaMemcpyAsync(hs.data(), scal, sizeof(double) * (mq + 1), cudaMemcpyDeviceToHost, h.stream));
    if (mq > 0) CU(cudaMemcpyAsync(hq.data(), P->gr + ml, sizeof(double) * mq, cudaMemcpyDeviceToHost, h.stream));
    if (P->has_feasible) {      // assert norm(x0 - (z0 + F u0)) < tolEqSolve   (BarrierSolver.scala:214-218)
      T(cvxb::solution_space_map(h, S, R->x_feas, Pz));
    }
    CU(cudaStreamSynchronize(h.stream));
  }
  if (st == CVXB_OK && P->has_feasible) {
    std::vector<double> a((size_t)n), b((size_t)n);
    CU(cudaMemcpy(a.data(), Pz, sizeof(double) * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(b.data(), P->x_feas, sizeof(double) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) dist += (a[i] - b[i]) * (a[i] - b[i]);
    dist = sqrt(dist);
  }
  cudaFree(T1);
  cudaFree(vec);
  if (st == CVXB_OK && P->has_feasible && !(dist < tol)) {
    cvxb::set_last_error("reduced: u0 does not map to x0 under the variable transform (||x0 - (z0 + F u0)|| = %.3g, A x0 != b?)", dist);
    st = CVXB_EINVAL;
  }
  if (st != CVXB_OK) { problem_free(R); return st; }
  R->obj_r = P->obj_r + hs[0];
  R->obj_pow = P->obj_pow;
  R->has_feasible = P->has_feasible;
  if (mq > 0) {
    for (int q = 0; q < mq; ++q) hq[q] += hs[1 + q];
    if (cudaMemcpy(R->gr + ml, hq.data(), sizeof(double) * mq, cudaMemcpyHostToDevice) != cudaSuccess) { problem_free(R); return CVXB_ECUDA; }
  }
  *out = R;
  return CVXB_OK;
}

int cvxb_problem_reduce(cvxb_handle h, cvxb_problem prob, cvxb_solution_space space, const cvxb_params* pars, cvxb_problem* out) {
  CHECK_HP(h, prob);
  if (!space || !out) { cvxb::set_last_error("cvxb_problem_reduce: null argument"); return CVXB_EINVAL; }
  cvxb_params dp;
  if (!pars) { cvxb_default_params(&dp); pars = &dp; }
  return problem_reduce(prob, (cvxb::SolutionSpaceDev*)space, pars->tolEqSolve, out);
}

int cvxb_problem_destroy(cvxb_problem prob) {
  if (!prob) return CVXB_OK;
  cvxb::DeviceGuard _guard(prob->h->device);
  cudaStreamSynchronize(prob->h->stream);
  problem_free(prob);
  return CVXB_OK;
}

static void init_run(const cvxb_params* pars, RunStats& rs) {
  rs.limited = pars->stepLimit > 0;
  rs.budget = pars->stepLimit;
}

int cvxb_phase1(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, double* x_feas, cvxb_solution* ph) {
  CHECK_HP(h, prob);
  cvxb_params dp;
  if (!pars) { cvxb_default_params(&dp); pars = &dp; }
  RunStats rs;
  init_run(pars, rs);
  CVXB_CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  int st = run_phase1(prob, *pars, rs, ph);
  cudaEventRecord(h->ev1, h->stream);
  cudaEventSynchronize(h->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  if (ph) {
    ph->solve_ms = ms;
    if (prob->phase1) download_vec(*h, ph->x, prob->phase1->x, prob->n + 1);
  }
  if (st == CVXB_OK && x_feas) CVXB_TRY(download_vec(*h, x_feas, prob->x_feas, prob->n));
  cudaStreamSynchronize(h->stream);
  return st;
}

int cvxb_barrier_solve(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, cvxb_solution* out) {
  CHECK_HP(h, prob);
  if (!out) { cvxb::set_last_error("cvxb_barrier_solve: null solution"); return CVXB_EINVAL; }
  cvxb_params dp;
  if (!pars) { cvxb_default_params(&dp); pars = &dp; }
  RunStats rs;
  init_run(pars, rs);
  double *ox = out->x, *ol = out->lambda, *onu = out->nu;
  memset(out, 0, sizeof(*out));
  out->x = ox; out->lambda = ol; out->nu = onu;
  CVXB_CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  long long ph_steps = 0, ph_exec = 0;
  int ph_stages = 0;
  double ph_s = 0;
  if (!prob->has_feasible) {   // OptimizationProblem.withoutFeasiblePoint: phase I first (OptimizationProblem.scala:174-196)
    cvxb_solution ph;
    memset(&ph, 0, sizeof(ph));
    int st = run_phase1(prob, *pars, rs, &ph);
    ph_steps = ph.newton_steps; ph_exec = ph.executed_newton_steps; ph_stages = ph.outer_stages; ph_s = ph.phase1_s;
    if (st != CVXB_OK || (rs.limited && rs.budget <= 0)) {
      out->phase1_newton_steps = ph_steps; out->phase1_executed_steps = ph_exec; out->phase1_stages = ph_stages;
      out->phase1_s = ph_s;
      cudaEventRecord(h->ev1, h->stream); cudaEventSynchronize(h->ev1);
      float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); out->solve_ms = ms;
      if (st == CVXB_OK && prob->phase1) {   // step-limited run that ended inside phase I: hand back the current iterate
        download_vec(*h, out->x, prob->phase1->x, prob->n);
        cudaStreamSynchronize(h->stream);
      }
      return st;
    }
  }
  CVXB_CUDA_OK(cudaMemcpyAsync(prob->x, prob->x_feas, prob->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  int st = barrier_loop(prob, *pars, TERM_STANDARD, rs, out);
  out->phase1_newton_steps = ph_steps; out->phase1_executed_steps = ph_exec; out->phase1_stages = ph_stages;
  out->phase1_s = ph_s;
  cudaEventRecord(h->ev1, h->stream);
  cudaEventSynchronize(h->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  out->solve_ms = ms;
  if (st != CVXB_OK) return st;
  CVXB_TRY(download_vec(*h, out->x, prob->x, prob->n));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_kldual_primal_optimum(cvxb_handle h, cvxb_problem prob, double* x_primal) {
  CHECK_HP(h, prob);
  if (prob->objective != CVXB_OBJ_KLDUAL || !x_primal) { cvxb::set_last_error("not a dual KL problem"); return CVXB_EINVAL; }
  cvxb_problem_s* P = prob;
  CVXB_TRY(gemv_t(*h, P->n, P->kd, 1.0, P->obj_P, P->ldn, P->x, 0.0, P->du));
  CVXB_LAUNCH(*h, dual_y_kernel, (P->kd + 255) / 256, 256, 0, P->kd, P->objR, P->du, P->dy);
  CVXB_TRY(download_vec(*h, x_primal, P->dy, P->kd));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_barrier_newton_direction(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, const double* x, double t,
                                  double* H_out, double* g_out, double* dx, double* nu, cvxb_kkt_info* info) {
  CHECK_HP(h, prob);
  if (!x) { cvxb::set_last_error("null x"); return CVXB_EINVAL; }
  cvxb_params dp;
  if (!pars) { cvxb_default_params(&dp); pars = &dp; }
  cvxb_problem_s* P = prob;
  CVXB_TRY(upload_vec(*h, P->x, x, P->n));
  CVXB_TRY(barrier_eval(P, t));
  CVXB_TRY(barrier_hessian(P, t));
  CVXB_TRY(fetch_status(*h));
  if (h->h_flag[F_INFEAS]) { cvxb::set_last_error("hessianBarrierFunction: x not strictly feasible"); return CVXB_ENOTFEASIBLE; }
  int st;
  if (P->p > 0)
    st = kkt_solve_device(*h, P->kw, *pars, P->H, P->ldn, P->A, P->ldp, P->y, P->eqdiff, pars->tolEqSolve, P->dir, P->nu, info);
  else
    st = chol_solve_device(*h, P->kw, *pars, P->H, P->ldn, P->y, -1.0, pars->tolEqSolve, P->dir, info);
  if (H_out) {
    bool dev = (h->flags & CVXB_FLAG_DEVICE_PTRS) != 0;
    CVXB_CUDA_OK(cudaMemcpy2DAsync(H_out, (size_t)P->n * sizeof(double), P->H, (size_t)P->ldn * sizeof(double),
                                   (size_t)P->n * sizeof(double), P->n, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                   h->stream));
  }
  CVXB_TRY(download_vec(*h, g_out, P->y, P->n));
  if (st == CVXB_OK) {
    CVXB_TRY(download_vec(*h, dx, P->dir, P->n));
    if (P->p > 0) CVXB_TRY(download_vec(*h, nu, P->nu, P->p));
  }
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return st;
}

// g(x) for every constraint (linear rows, then quadratic), and ConstraintSet.isSatisfiedStrictlyBy(x).
int cvxb_constraint_values(cvxb_handle h, cvxb_problem prob, const double* x, double* g, int* strictly_satisfied) {
  CHECK_HP(h, prob);
  if (!x) { cvxb::set_last_error("null x"); return CVXB_EINVAL; }
  if (h->flags & CVXB_FLAG_DEVICE_PTRS) { cvxb::set_last_error("cvxb_constraint_values takes host pointers"); return CVXB_EINVAL; }
  cvxb_problem_s* P = prob;
  CVXB_TRY(upload_vec(*h, P->x, x, P->n));
  CVXB_TRY(barrier_eval(P, 1.0));
  std::vector<double> gv((size_t)P->m), ub((size_t)P->m);
  CVXB_CUDA_OK(cudaMemcpyAsync(gv.data(), P->gx, sizeof(double) * P->m, cudaMemcpyDeviceToHost, h->stream));
  CVXB_CUDA_OK(cudaMemcpyAsync(ub.data(), P->ub, sizeof(double) * P->m, cudaMemcpyDeviceToHost, h->stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  int ok = 1;
  for (int i = 0; i < P->m; ++i) {
    if (g) g[i] = gv[i];
    if (!(gv[i] * (1.0 + 3e-16) < ub[i])) ok = 0;      // Constraint.scala:23
  }
  if (strictly_satisfied) *strictly_satisfied = ok;
  return CVXB_OK;
}

}  // extern "C"

// Batched small-problem barrier solver: B independent problems with
// n <= 64 variables, m <= 128 linear inequalities and p in {0,1} equalities, ONE CTA PER PROBLEM.
// The whole barrier solve of a problem -- every Newton step's Hessian assembly (FP64 DMMA out of shared
// memory), Ruiz equilibration, Cholesky, triangular solves, Schur complement (a scalar for p = 1),
// residual checks, line search and the outer t *= mu loop -- runs on-chip: G (66 KB) and the n x n work
// matrix (33 KB) live in shared memory for the life of the problem; HBM sees one read of the problem
// and one write of x.  CTAs are persistent and pull problem indices from an atomic counter, so the
// different iteration counts of different problems balance themselves.
//
// Semantics are those of the large path (solver.cu / kkt.cu), i.e. of BarrierSolver.scala:70-188,
// EqualityConstrainedSolver.scala:37-107, UnconstrainedSolver.scala:34-125 (incl. D4, D7),
// KKTSystem.scala:43-246 (path 0, regularised retry, path 1 = H + A'A) and MatrixUtils.scala:240-516.
// The decomposition last resort kktSymSolve (KKTSystem.scala:283-310) runs inside the CTA as well (one-sided Jacobi SVD
// of the (n+1)^2 KKT matrix in the shared-memory block of G, which is reloaded afterwards).
// Problems flagged `phase1` start from a point where they are merely defined: the CTA first runs the reference's phase-I
// feasibility analysis (ConstraintSet.scala:326-395, 556-575) on the (n+1) x (m+2p) problem it builds in place.
#include "kkt.cuh"
#include "vecops.cuh"

using namespace cvxb;

struct cvxb_batch_s {
  cvxb_handle_s* h = nullptr;
  int B = 0, n = 0, m = 0, p = 0;
  int* objective = nullptr;
  int* pcount = nullptr;
  double *obj_a = nullptr, *obj_r = nullptr, *obj_P = nullptr, *G = nullptr, *ub = nullptr, *A = nullptr, *b = nullptr,
         *x0 = nullptr;
  // outputs (device)
  double *x = nullptr, *objval = nullptr, *gap = nullptr, *eqgap = nullptr;
  int *status = nullptr, *steps = nullptr, *stages = nullptr;
  int* stage_steps = nullptr;    // B x CVXB_BATCH_STAGES: Newton steps of each of the first outer stages
  int* order = nullptr;          // pickup order of the problems (longest expected first), or NULL = index order
  int* phase1 = nullptr;         // B or NULL
  int *ph_steps = nullptr, *ph_stages = nullptr;
  double* ph_s = nullptr;
  long long* cycles = nullptr;   // B: SM clock cycles each problem spent in its CTA
  double* records = nullptr;     // B x (n + CVXB_BATCH_RECORD_EXTRA): [x, objective, gap, status, steps, stages] per problem
  double* scratch = nullptr;     // per-CTA copy of H (n x n)
  unsigned* counter = nullptr;
  int grid = 0;
  std::vector<void*> owned;
};

namespace cvxb {
namespace {

constexpr int BN = 64, BM_ = 128;          // capacity
constexpr int LDG = BM_ + 4;               // G col-major in shared memory; LDG mod 16 == 4 -> conflict-free DMMA fragments
constexpr int LDH = BN + 1;                // H / L col-major, odd stride
constexpr int BT = 256;                    // threads per CTA
constexpr int BSUB = 16;
constexpr double IN_SET = 1.0 + 3e-16;

struct BatchArgs {
  int B, n, m, p;
  const int* objective;
  const int* pcount;
  const double *obj_a, *obj_r, *obj_P, *G, *ub, *A, *b, *x0;
  double *x, *objval, *gap, *eqgap;
  int *status, *steps, *stages;
  int* stage_steps;
  const int* order;
  long long* cycles;
  double* records;
  double* scratch;
  unsigned* counter;
  const int* phase1;       // B or NULL: 1 = x0 is only a point where the problem is defined, run phase I first
  int* ph_steps;           // phase-I Newton steps / outer stages / final slack s (NULL when no problem asks for phase I)
  int* ph_stages;
  double* ph_s;
  cvxb_params P;
};

struct Smem {
  double G[BN * LDG];
  double L[BN * LDH];
  double x[BN], y[BN], dir[BN], dr[BN], qs[BN], ya[BN], yq[BN], aeq[BN], oa[BN], Px[BN], Pd[BN], rdiag[BN], zrhs[BN];
  double work[4 * BN];      // t1 | t2 | xtry | spare during a solve; Ruiz column partial sums before it
  double gx[BM_], ub[BM_], inv[BM_], Gd[BM_];
  double red[40];
  int ired[40];
  double sc[16];
  int fl[8];
  int pcur;                 // equalities of the current problem (0 or 1)
  int ncur, mcur;           // dimensions of the problem the CTA is working on (the phase-I pass: n + 1, m + 2p)
  int stage_steps[CVXB_BATCH_STAGES];
};
#define S_T1 (S.work)
#define S_T2 (S.work + BN)
#define S_XTRY (S.work + 2 * BN)
#define S_COLSQ (S.work)

__device__ __forceinline__ void dmma884b(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

#ifndef CVXB_BATCH_TIMING
#define CVXB_BATCH_TIMING 0
#endif
__device__ long long g_batch_clk[12];
#define BCLK(slot)                                                              \
  do {                                                                          \
    if (CVXB_BATCH_TIMING && blockIdx.x == 0 && threadIdx.x == 0) {             \
      long long _c = clock64();                                                 \
      g_batch_clk[slot] += _c - g_batch_clk[11];                                \
      g_batch_clk[11] = _c;                                                     \
    }                                                                           \
  } while (0)

__device__ long long g_potrf_clk[8];      // CVXB_BATCH_TIMING: diag / panel / update / barriers inside b_potrf
#define PCLK(slot)                                                              \
  do {                                                                          \
    if (CVXB_BATCH_TIMING && blockIdx.x == 0 && threadIdx.x == 0) {             \
      long long _c = clock64();                                                 \
      g_potrf_clk[slot] += _c - g_potrf_clk[7];                                 \
      g_potrf_clk[7] = _c;                                                      \
    }                                                                           \
  } while (0)

#define HL(i, j) S.L[(i) + (j) * LDH]
#define GG(i, j) S.G[(i) + (j) * LDG]

// Three block sums behind ONE set of barriers (a block_sum costs three __syncthreads; a Newton step had ~18 of them).  The
// additions of each sum happen in the same order as in vecops.cuh: block_sum (xor tree inside the warp, then the xor tree
// over the 8 warp totals), so the results are bit-identical to three separate calls.  `buf` >= 35 doubles.
__device__ __forceinline__ void b_sum3(double& a, double& b, double& c, double* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  __syncthreads();
  if (lane == 0) { buf[warp] = a; buf[8 + warp] = b; buf[16 + warp] = c; }
  __syncthreads();
  if (warp == 0) {
    double t = lane < 24 ? buf[lane] : 0.0;
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    if (lane < 24 && (lane & 7) == 0) buf[32 + (lane >> 3)] = t;
  }
  __syncthreads();
  a = buf[32];
  b = buf[33];
  c = buf[34];
}

// ---- evaluation at S.x for parameter t: gx, 1/slack, barrier value, gradient, eq residual -----------------
// returns false when x is not strictly feasible (slack <= 0)
__device__ bool b_eval(Smem& S, const BatchArgs& A, int kind, double obj_r, const double* Pg, double beq, double t,
                       double* fval, double* f0out, double* normGrad, double* eqdiff) {
  const int tid = threadIdx.x, n = S.ncur, m = S.mcur;
  if (tid < m) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = fma(GG(tid, j), S.x[j], s);
    S.gx[tid] = s;
  }
  if (kind == CVXB_OBJ_QUADRATIC && tid >= 128 && tid < 128 + n) {
    const int i = tid - 128;
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = fma(Pg[i + (size_t)j * n], S.x[j], s);
    S.Px[i] = s;
  }
  __syncthreads();
  double ls = 0.0;
  int bad = 0;
  if (tid < m) {
    double d = S.ub[tid] - S.gx[tid];
    if (!(d > 0.0)) bad = 1;
    S.inv[tid] = 1.0 / d;
    ls = log(d);
  }
  double f0 = 0.0;
  if (tid < n) {
    double xj = S.x[tid];
    if (kind == CVXB_OBJ_LINEAR) f0 = S.oa[tid] * xj;
    else if (kind == CVXB_OBJ_QUADRATIC) f0 = S.oa[tid] * xj + 0.5 * xj * S.Px[tid];
    else f0 = xj * log(xj * (double)n);
  }
  double nbad = (double)bad;
  b_sum3(ls, nbad, f0, S.red);          // (writes of S.inv above are ordered before the gradient's reads by its barriers)
  bad = nbad != 0.0;
  f0 += obj_r;
  // gradient: y_j = t grad f0_j + sum_i G(i,j)/d_i ; 4 threads per column
  {
    const int j = tid >> 2, part = tid & 3;
    double s = 0.0;
    if (j < n)
      for (int i = part; i < m; i += 4) s = fma(GG(i, j), S.inv[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (j < n && part == 0) {
      double xj = S.x[j], gf;
      if (kind == CVXB_OBJ_LINEAR) gf = S.oa[j];
      else if (kind == CVXB_OBJ_QUADRATIC) gf = S.oa[j] + S.Px[j];
      else gf = 1.0 + log(xj) + log((double)n);
      S.y[j] = t * gf + s;
    }
  }
  __syncthreads();
  double g2 = 0.0, ax = 0.0;
  if (tid < n) {
    g2 = S.y[tid] * S.y[tid];
    if (S.pcur) ax = S.aeq[tid] * S.x[tid];
  }
  double dummy = 0.0;
  b_sum3(g2, ax, dummy, S.red);
  *fval = t * f0 - ls;
  *f0out = f0;
  *normGrad = sqrt(g2);
  *eqdiff = S.pcur ? (beq - ax) : 0.0;
  return bad == 0;
}

// ---- H = t hess f0 + G' diag(inv^2) G, full symmetric in S.L, via DMMA out of shared memory --------------
// The 8 x 8 output tiles of the lower triangle are dealt out in PAIRS of neighbours in a tile row: both share the
// A-operand fragment (rows of G' scaled by inv^2), and their four independent accumulator chains double the DMMAs in
// flight per warp -- the phase is bound by the latency of LDS -> DMMA chains, not by the tensor pipe (20 work items in
// 3 rounds over the 8 warps instead of 36 tiles in 5).  inv^2 is staged once in S.Gd (dead until the line search).
__device__ void b_hessian(Smem& S, const BatchArgs& A, int kind, const double* Pg, double t) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = S.ncur, m = S.mcur;
  const int g = lane >> 2, tq = lane & 3;
  const int nt = (n + 7) >> 3;
  const int m4 = (m + 3) & ~3;
  if (tid < BM_) S.Gd[tid] = (tid < m) ? S.inv[tid] * S.inv[tid] : 0.0;
  __syncthreads();
  int item = 0;
  for (int ti = 0; ti < nt; ++ti) {
    for (int tj = 0; tj <= ti; tj += 2, ++item) {
      if ((item & (BT / 32 - 1)) != warp) continue;
      const bool two = tj + 1 <= ti;
      const int i0 = ti * 8, j0 = tj * 8, j1 = two ? j0 + 8 : j0;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0, h0 = 0.0, h1 = 0.0;
      for (int kk = 0; kk < m4; kk += 8) {
        {
          const int r = kk + tq;
          const double a = GG(r, i0 + g) * S.Gd[r];
          dmma884b(c0, c1, a, GG(r, j0 + g));
          if (two) dmma884b(f0, f1, a, GG(r, j1 + g));
        }
        if (kk + 4 < m4) {
          const int r = kk + 4 + tq;
          const double a = GG(r, i0 + g) * S.Gd[r];
          dmma884b(e0, e1, a, GG(r, j0 + g));
          if (two) dmma884b(h0, h1, a, GG(r, j1 + g));
        }
      }
      c0 += e0; c1 += e1; f0 += h0; f1 += h1;
      const int i = i0 + g;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half && !two) break;
        const int jb = half ? j1 : j0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = jb + 2 * tq + e;
          if (i < n && j < n && j <= i) {
            double v = half ? (e ? f1 : f0) : (e ? c1 : c0);
            if (kind == CVXB_OBJ_QUADRATIC) v += t * Pg[i + (size_t)j * n];
            else if (kind == CVXB_OBJ_KL && i == j) v += t / S.x[i];
            HL(i, j) = v;
            HL(j, i) = v;
          }
        }
      }
    }
  }
  __syncthreads();
}

// ---- Ruiz equilibration of the full symmetric H in S.L -> S.dr  (MatrixUtils.scala:240-268) ---------------
__device__ void b_ruiz(Smem& S, const BatchArgs& A) {
  const int tid = threadIdx.x, n = S.ncur, lane = tid & 31, warp = tid >> 5;
  if (tid < n) S.dr[tid] = 1.0;
  __syncthreads();
  for (int sweep = 0; sweep < A.P.ruizMaxSweeps; ++sweep) {
    const int j = tid & 63, part = tid >> 6;
    double s = 0.0;
    if (j < n) {
      const double dj = S.dr[j];
      for (int i = part; i < n; i += 4) {
        double q = (S.dr[i] * dj) * HL(i, j);
        s = fma(q, q, s);
      }
    }
    S_COLSQ[part * BN + j] = s;
    __syncthreads();
    // two barriers per sweep: the 64 column owners live in warps 0 and 1, so rho = max|1-u| needs one
    // shuffle reduction per warp and a two-entry exchange.  (A one-barrier variant -- every warp owning 8 columns and
    // computing u_j itself -- measured 38 % slower: the sqrt / divide sequences then issue in all 8 warps of both
    // resident CTAs instead of in 2.)
    if (warp < 2) {
      double dev = 0.0;
      if (tid < n) {
        double tot = (S_COLSQ[tid] + S_COLSQ[BN + tid]) + (S_COLSQ[2 * BN + tid] + S_COLSQ[3 * BN + tid]);
        double u = sqrt(sqrt(tot));
        if (u > 0) S.dr[tid] = S.dr[tid] * (1.0 / u);
        dev = fabs(1.0 - u);
        if (dev != dev) dev = 1e308;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
      if (lane == 0) S.red[warp] = dev;
    }
    __syncthreads();
    const double rho = fmax(S.red[0], S.red[1]);
    if (!(rho > A.P.ruizTol)) break;      // uniform: every thread sees the same rho
  }
  __syncthreads();
}

// ---- in-place Cholesky of the lower triangle of S.L (n <= 64); returns 0 or the failing column ------------
__device__ int b_potrf(Smem& S, int n, double* mind_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { S.fl[0] = 0; S.sc[0] = 1e300; }
  __syncthreads();
  const int nblk = (n + BSUB - 1) / BSUB;
  PCLK(0);
  for (int kb = 0; kb < nblk; ++kb) {
    const int o = kb * BSUB;
    const int bs = (n - o) < BSUB ? (n - o) : BSUB;
    PCLK(4);
    if (warp == 0) {
      // branch-free in-warp sweep (factor.cu: warp_diag_factor): selects only, updates unpredicated
      double a[BSUB], rr[BSUB];
#pragma unroll
      for (int c = 0; c < BSUB; ++c)
        a[c] = (lane < bs && c < bs && c <= lane) ? HL(o + lane, o + c) : ((c == lane) ? 1.0 : 0.0);
      double mind = 1e300;
      int fail = 0;
#pragma unroll
      for (int j = 0; j < BSUB; ++j) {
        double d = __shfl_sync(0xffffffffu, a[j], j);
        const bool bad = !(d > 0.0) || d > 1e300;
        fail = (bad && !fail && j < bs) ? (o + j + 1) : fail;
        d = bad ? 1.0 : d;
        double r;
        if (d < 1e-30 || d > 1e30) r = rsqrt(d);       // warp-uniform
        else {                                   // float seed + one cubic correction (factor.cu: pivot_rsqrt)
          const double y0 = (double)rsqrtf((float)d);
          const double e = fma(-(d * y0), y0, 1.0);
          r = fma(y0, fma(e, 0.375, 0.5) * e, y0);
        }
        const double l = d * r;
        mind = (j < bs && l < mind) ? l : mind;
        rr[j] = r;
        a[j] = (lane == j) ? l : a[j] * r;
#pragma unroll
        for (int c = j + 1; c < BSUB; ++c) {
          const double tt = __shfl_sync(0xffffffffu, a[j], c);
          a[c] = fma(-a[j], tt, a[c]);
        }
      }
      if (lane == 0) {
        if (fail && !S.fl[0]) S.fl[0] = fail;
        if (mind < S.sc[0]) S.sc[0] = mind;
      }
#pragma unroll
      for (int c = 0; c < BSUB; ++c) {
        if (lane < bs && c <= lane) HL(o + lane, o + c) = a[c];
        if (lane == c) S.rdiag[o + c] = rr[c];
      }
    }
    PCLK(1);
    __syncthreads();
    const int r0 = o + BSUB;
    const int nrows = n - r0;
    if (nrows <= 0) continue;
    if (tid < nrows) {
      // right-looking substitution (factor.cu: leaf panel): once x[c] is known it is folded into all later columns, so the
      // dependent chain is 16 x (multiply + one FMA) instead of the 136 in-order FMAs of the dot-product form
      const int r = r0 + tid;
      double xr[BSUB];
#pragma unroll
      for (int c = 0; c < BSUB; ++c) xr[c] = HL(r, o + c);
#pragma unroll
      for (int c = 0; c < BSUB; ++c) {
        xr[c] *= S.rdiag[o + c];
#pragma unroll
        for (int c2 = c + 1; c2 < BSUB; ++c2) xr[c2] = fma(-xr[c], HL(o + c2, o + c), xr[c2]);
      }
#pragma unroll
      for (int c = 0; c < BSUB; ++c) HL(r, o + c) = xr[c];
    }
    __syncthreads();
    PCLK(2);
    {
      // trailing update A(r,c) -= sum_k L(r,o+k) L(c,o+k), r >= c >= r0, as 8 x 8 DMMA tiles of the lower triangle
      // (rows past n only feed discarded outputs and are read as zero)
      const int g = lane >> 2, tq = lane & 3;
      const int T = (nrows + 7) >> 3;
      const int ntile = T * (T + 1) / 2;
      for (int tl = warp; tl < ntile; tl += BT / 32) {
        int ti = (int)((sqrtf(8.0f * (float)tl + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
        while (ti * (ti + 1) / 2 > tl) --ti;
        const int tj = tl - ti * (ti + 1) / 2;
        const int rr = r0 + 8 * ti + g, cr = r0 + 8 * tj + g;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int kk = 0; kk < BSUB; kk += 4)
          dmma884b(c0, c1, rr < n ? HL(rr, o + kk + tq) : 0.0, cr < n ? HL(cr, o + kk + tq) : 0.0);
        const int cc = r0 + 8 * tj + 2 * tq;
        if (rr < n) {
          if (cc <= rr) HL(rr, cc) -= c0;
          if (cc + 1 <= rr) HL(rr, cc + 1) -= c1;
        }
      }
    }
    __syncthreads();
    PCLK(3);
  }
  *mind_out = S.sc[0];
  return S.fl[0];
}

// ---- triangular solves with the factor in S.L, by warp 0 (and warp 1 for a second right-hand side) -------
// forward: v := L^-1 v ; backward: v := L^-T v.  Rows lane and lane + 32.  (A blocked variant -- inverses of the 16-column
// diagonal blocks kept in their upper triangles, 4 block steps of 16 x 16 products instead of 64 row steps -- measured
// 17 % slower overall: twice the instructions, and the row-step chain is shorter than it looks.)
__device__ __forceinline__ void b_trsv_warp(Smem& S, int n, double* v, bool trans) {
  const int lane = threadIdx.x & 31;
  double b0 = lane < n ? v[lane] : 0.0, b1 = lane + 32 < n ? v[lane + 32] : 0.0;
  // selects instead of branches around the shuffles; entries already solved are simply never read again
  if (!trans) {
    for (int j = 0; j < n; ++j) {
      const double src = (j < 32) ? b0 : b1;
      const double yj = __shfl_sync(0xffffffffu, src, j & 31) * S.rdiag[j];
      const double l0 = (lane > j && lane < n) ? HL(lane, j) : 0.0;
      const double l1 = (lane + 32 > j && lane + 32 < n) ? HL(lane + 32, j) : 0.0;
      b0 = (lane == j) ? yj : fma(-l0, yj, b0);
      b1 = (lane + 32 == j) ? yj : fma(-l1, yj, b1);
    }
  } else {
    for (int j = n - 1; j >= 0; --j) {
      const double src = (j < 32) ? b0 : b1;
      const double xj = __shfl_sync(0xffffffffu, src, j & 31) * S.rdiag[j];
      const double l0 = (lane < j) ? HL(j, lane) : 0.0;
      const double l1 = (lane + 32 < j) ? HL(j, lane + 32) : 0.0;
      b0 = (lane == j) ? xj : fma(-l0, xj, b0);
      b1 = (lane + 32 == j) ? xj : fma(-l1, xj, b1);
    }
  }
  if (lane < n) v[lane] = b0;
  if (lane + 32 < n) v[lane + 32] = b1;
}

// t2 = L (L' v)   (for the residual with L L' in place of Q)
__device__ void b_llt_apply(Smem& S, int n, const double* v, double* out) {
  const int tid = threadIdx.x;
  if (tid < n) {
    double s = 0.0;
    for (int i = tid; i < n; ++i) s = fma(HL(i, tid), v[i], s);      // (L'v)_tid
    S_T1[tid] = s;
  }
  __syncthreads();
  if (tid < n) {
    double s = 0.0;
    for (int k = 0; k <= tid; ++k) s = fma(HL(tid, k), S_T1[k], s);
    out[tid] = s;
  }
  __syncthreads();
}

// restore H (full) from the per-CTA scratch, optionally adding delta*I and a rank-one term aa'.
// The scratch copy has the fixed column stride BN (= 64): thread t walks rows t & 63 of columns t >> 6, +4, ... -- no
// integer division by the runtime n in these whole-matrix passes (it was a fifth of the kernel's instructions).
__device__ void b_load_H(Smem& S, int n, const double* Hs, double diag_add, bool rank1) {
  const int i = threadIdx.x & (BN - 1);
  if (i < n)
    for (int j = threadIdx.x >> 6; j < n; j += BT / BN) {
      double v = Hs[i + j * BN];
      if (i == j) v += diag_add;
      if (rank1) v += S.aeq[i] * S.aeq[j];
      HL(i, j) = v;
    }
  __syncthreads();
}
__device__ void b_store_H(Smem& S, int n, double* Hs) {
  const int i = threadIdx.x & (BN - 1);
  if (i < n)
    for (int j = threadIdx.x >> 6; j < n; j += BT / BN) Hs[i + j * BN] = HL(i, j);
  __syncthreads();
}

// (d d') o H in place, + delta on the diagonal
__device__ void b_scale_H(Smem& S, int n, double delta) {
  const int i = threadIdx.x & (BN - 1);
  if (i < n) {
    const double di = S.dr[i];
    for (int j = threadIdx.x >> 6; j < n; j += BT / BN) {
      double v = (di * S.dr[j]) * HL(i, j);
      if (i == j) v += delta;
      HL(i, j) = v;
    }
  }
  __syncthreads();
}

// One solvePD / choleskySolve attempt chain on the matrix currently in the scratch copy Hs (+ modifiers).
//   p == 0: dir = choleskySolve(Hmod, -y)          (rhs = -y)
//   p == 1: KKT  Hmod dir + a' w = -q, a.dir = beq_rhs
// returns true when the attempt chain (plain, then regularised) produced an accepted solution.
__device__ bool b_linear_solve(Smem& S, const BatchArgs& A, const double* Hs, double diag_add, bool rank1,
                               const double* q, double brhs, double tol, int* regularized, bool h_in_smem = false) {
  const int tid = threadIdx.x, n = S.ncur;
  bool have_dr = false;
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (!(h_in_smem && attempt == 0)) b_load_H(S, n, Hs, diag_add, rank1);     // (the plain H is still in S.L on the first try)
    BCLK(3);
    if (!have_dr) { b_ruiz(S, A); have_dr = true; }
    BCLK(4);
    b_scale_H(S, n, attempt ? A.P.cholRegDelta : 0.0);
    double mind;
    BCLK(5);
    int fail = b_potrf(S, n, &mind);
    BCLK(6);
    if (attempt == 0 && (fail || !(mind > A.P.cholMinDiag))) { *regularized = 1; continue; }
    if (fail) return false;
    if (S.pcur == 0) {
      // w = L^-1 (d o (-q)) ; u = L^-T w ; dir = d o u ; residual || Hmod dir + q || / relsize(q)
      if (tid < n) S.qs[tid] = S.dr[tid] * (-q[tid]);
      __syncthreads();
      if (tid < 32) { b_trsv_warp(S, n, S.qs, false); __syncwarp(); b_trsv_warp(S, n, S.qs, true); }
      __syncthreads();
      if (tid < n) S.dir[tid] = S.dr[tid] * S.qs[tid];
      __syncthreads();
      double r2 = 0.0, q2 = 0.0;
      if (tid < n) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) {
          double hij = Hs[tid + j * BN];
          if (tid == j) hij += diag_add;
          s = fma(hij, S.dir[j], s);
        }
        double r = s + q[tid];
        r2 = r * r;
        q2 = q[tid] * q[tid];
      }
      double dummy = 0.0;
      b_sum3(r2, q2, dummy, S.red);
      double e1 = relative_size(sqrt(r2), sqrt(q2), tol);
      return e1 <= tol;
    }
    // ---- p == 1 block elimination (KKTSystem.scala:99-167 on the equilibrated system)
    if (tid < n) { S.ya[tid] = S.dr[tid] * S.aeq[tid]; S.qs[tid] = S.dr[tid] * q[tid]; S.yq[tid] = S.qs[tid]; }
    __syncthreads();
    if (tid < 32) b_trsv_warp(S, n, S.ya, false);
    else if (tid < 64) b_trsv_warp(S, n, S.yq, false);
    __syncthreads();
    double sa = 0.0, saq = 0.0, nq = 0.0;
    if (tid < n) { sa = S.ya[tid] * S.ya[tid]; saq = S.ya[tid] * S.yq[tid]; nq = S.qs[tid] * S.qs[tid]; }
    b_sum3(sa, saq, nq, S.red);
    if (!(sa > 0.0)) return false;               // cholesky(S) of the 1 x 1 Schur complement fails
    const double K = sqrt(sa);
    const double z = -(brhs + saq);
    const double w = (z / K) / K;
    if (tid < n) S.yq[tid] = S.yq[tid] + S.ya[tid] * w;
    __syncthreads();
    if (tid < 32) b_trsv_warp(S, n, S.yq, true);
    __syncthreads();
    if (tid < n) { S_T2[tid] = -S.yq[tid]; }        // xs
    __syncthreads();
    b_llt_apply(S, n, S_T2, S_XTRY);               // L L' xs
    double r2 = 0.0, axs = 0.0;
    if (tid < n) {
      double r = S_XTRY[tid] + S.dr[tid] * S.aeq[tid] * w + S.qs[tid];
      r2 = r * r;
      S.dir[tid] = S.dr[tid] * S_T2[tid];
      axs = S.aeq[tid] * S.dir[tid];
    }
    double dummy = 0.0;
    b_sum3(r2, axs, dummy, S.red);
    double e1 = relative_size(sqrt(r2), sqrt(nq), tol);
    double e2 = relative_size(fabs(axs - brhs), fabs(brhs), tol);
    if (tid == 0) S.sc[1] = w;
    __syncthreads();
    return (e1 <= tol) && (e2 <= tol);
  }
  return false;
}

// ---- KKTSystem.kktSymSolve (KKTSystem.scala:63, 283-310): last resort of the fallback chain for p == 1 -------------------
// M = [H a; a' 0] ((n+1)^2), rhs = (-q, b): symSolve = MatrixUtils.diagonalizationSolve (MatrixUtils.scala:603-751) through
// a one-sided (Hestenes) Jacobi SVD inside the CTA -- the in-CTA twin of eig.cu: svd_solve_device, same pairing order,
// same threshold, same two acceptance tests (range test, residual test; defect D3: no recovery once they fail).
// W = M V and V live in the shared-memory block of G and L (G is reloaded by the caller afterwards; L is scratch here).
// With with_eq = false the same routine is MatrixUtils.symSolve(H, -q), the last resort of UnconstrainedSolver.solve
// (UnconstrainedSolver.scala:58-65) after choleskySolve(H) and choleskySolve(H + 1e-9 I) have failed.
__device__ __noinline__ bool b_kkt_sym_solve(Smem& S, int n, bool with_eq, const double* Hs, const double* q, double brhs,
                                             double tol) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = n + (with_eq ? 1 : 0), ld = BN + 1;
  double* W = S.G;
  double* V = S.G + ld * ld;
  double* vec = S.L + 8;                 // V ends two entries into L
  double *z = vec, *b0 = vec + 72, *xs = vec + 144, *aw = vec + 216, *rhs = vec + 288;
  __syncthreads();
  for (int idx = tid; idx < N * N; idx += BT) {
    const int i = idx % N, j = idx / N;
    double v;
    if (i < n && j < n) v = Hs[i + j * BN];
    else if (i == n && j == n) v = 0.0;
    else v = S.aeq[i < n ? i : j];
    W[i + j * ld] = v;
    V[i + j * ld] = (i == j) ? 1.0 : 0.0;
  }
  if (tid < N) rhs[tid] = tid < n ? -q[tid] : brhs;
  __syncthreads();
  const int np = N + (N & 1), mrr = np - 1;
  for (int sweep = 0; sweep < 60 && N > 1; ++sweep) {
    if (tid == 0) S.fl[1] = 0;
    __syncthreads();
    for (int r = 0; r < np - 1; ++r) {
      for (int k = warp; k < np / 2; k += BT / 32) {
        const int ci = (r + k) % mrr;
        const int cj = (k == 0) ? mrr : (r + mrr - k) % mrr;
        if (ci >= N || cj >= N) continue;            // dummy column of an odd-sized problem
        double* wi = W + ci * ld;
        double* wj = W + cj * ld;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int t = lane; t < N; t += 32) {
          const double x = wi[t], y = wj[t];
          a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
        }
        a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
        if (!(fabs(g) > 1e-15 * sqrt(a * b))) continue;     // already orthogonal (also when a column is zero); warp-uniform
        const double zeta = (b - a) / (2.0 * g);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), sn = c * tt;
        double* vi = V + ci * ld;
        double* vj = V + cj * ld;
        for (int t = lane; t < N; t += 32) {
          double x = wi[t], y = wj[t];
          wi[t] = c * x - sn * y;
          wj[t] = sn * x + c * y;
          x = vi[t]; y = vj[t];
          vi[t] = c * x - sn * y;
          vj[t] = sn * x + c * y;
        }
        if (lane == 0) atomicAdd(&S.fl[1], 1);
      }
      __syncthreads();
    }
    if (S.fl[1] == 0) break;
    __syncthreads();
  }
  __syncthreads();
  // The matrix is symmetric: eigen-decomposition semantics with U = V (orthonormal by construction) and the Rayleigh
  // quotient lambda_j = v_j . w_j (eig.cu: sym_coeff_kernel):  c_j = v_j . rhs, a_j = c_j where lambda_j != 0 (exact-zero
  // test as in the reference), z_j = c_j / lambda_j.  aw doubles as the coefficient vector a until it is overwritten.
  for (int j = warp; j < N; j += BT / 32) {
    const double* w = W + j * ld;
    const double* v = V + j * ld;
    double lam = 0.0, vb = 0.0;
    for (int t = lane; t < N; t += 32) { lam = fma(v[t], w[t], lam); vb = fma(v[t], rhs[t], vb); }
    lam = warp_sum(lam); vb = warp_sum(vb);
    if (lane == 0) { z[j] = vb; aw[j] = lam; }
  }
  __syncthreads();
  if (warp == 0) {
    // eigenvalues at the rounding level of the largest one are solved as zeros (eig.cu: sym_finish_kernel)
    double mx = 0.0;
    for (int j = lane; j < N; j += 32) mx = fmax(mx, fabs(aw[j]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const double thr = 64.0 * 2.220446049250313e-16 * (double)N * mx;
    for (int j = lane; j < N; j += 32) {
      const double lam = aw[j], c = z[j];
      z[j] = (fabs(lam) > thr) ? c / lam : 0.0;
      aw[j] = (lam != 0.0) ? c : 0.0;
    }
  }
  __syncthreads();
  double sb = 0.0, sx = 0.0;
  if (tid < N)
    for (int j = 0; j < N; ++j) { sb = fma(V[tid + j * ld], aw[j], sb); sx = fma(V[tid + j * ld], z[j], sx); }
  __syncthreads();
  if (tid < N) {
    b0[tid] = sb;        // V (V' rhs) over the nonzero eigenvalues: projection of rhs onto the range
    xs[tid] = sx;        // V Lambda^-1 V' rhs
  }
  __syncthreads();
  if (tid < N) {
    double sa = 0.0;
    if (tid < n) {
      for (int j = 0; j < n; ++j) sa = fma(Hs[tid + j * BN], xs[j], sa);
      if (with_eq) sa = fma(S.aeq[tid], xs[n], sa);
    } else {
      for (int j = 0; j < n; ++j) sa = fma(S.aeq[j], xs[j], sa);
    }
    aw[tid] = sa;
  }
  __syncthreads();
  double nb = 0.0, d0 = 0.0, d1 = 0.0;
  if (tid < N) {
    const double bi = rhs[tid], r0 = bi - b0[tid], r1 = aw[tid] - bi;
    nb = bi * bi; d0 = r0 * r0; d1 = r1 * r1;
  }
  nb = block_sum(nb, S.red);
  d0 = block_sum(d0, S.red);
  d1 = block_sum(d1, S.red);
  const double relDist = relative_size(sqrt(d0), sqrt(nb), tol), relErr = relative_size(sqrt(d1), sqrt(nb), tol);
  if (tid < n) S.dir[tid] = xs[tid];
  if (tid == 0 && with_eq) S.sc[1] = xs[n];
  __syncthreads();
  return (relDist <= tol) && (relErr <= tol);
}

// Phase-I form of G (ConstraintSet.phase_I_Analysis): one more column of -1 and, for an equality, the two rows +-a
// (n, m, p: the dimensions of the problem itself)
__device__ void b_phase1_G(Smem& S, int n, int m, int p) {
  const int tid = threadIdx.x;
  if (tid < m + 2 * p) GG(tid, n) = -1.0;
  if (p && tid < n) { GG(m, tid) = S.aeq[tid]; GG(m + 1, tid) = -S.aeq[tid]; }
  __syncthreads();
}

// G of the current problem back into shared memory (after b_kkt_sym_solve used its space); phase1: in its phase-I form
__device__ __noinline__ void b_reload_G(Smem& S, const double* Gg, int n, int m, int p, bool phase1) {
  for (int idx = threadIdx.x; idx < BN * LDG; idx += BT) S.G[idx] = 0.0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < m * n; idx += BT) GG(idx % m, idx / m) = Gg[idx];
  __syncthreads();
  if (phase1) b_phase1_G(S, n, m, p);
}

__device__ bool b_in_set(Smem& S, int m, double s) {
  int out = 0;
  if (threadIdx.x < m) {
    const double lin = s * S.Gd[threadIdx.x];
    const double g = S.gx[threadIdx.x] + lin;
    const double margin = 3.6e-15 * (fabs(S.gx[threadIdx.x]) + fabs(lin) + fabs(S.ub[threadIdx.x]));   // see solver.cu: ls_in_set
    if (!(g * IN_SET + margin < S.ub[threadIdx.x])) out = 1;
  }
  return block_or(out, S.ired) == 0;
}

__device__ double b_value(Smem& S, const BatchArgs& A, int kind, double t, double s, double f0, double c1, double c2,
                          int* throws) {
  const int tid = threadIdx.x;
  double ls = 0.0;
  int bad = 0;
  if (tid < S.mcur) {
    double d = S.ub[tid] - (S.gx[tid] + s * S.Gd[tid]);
    if (!(d > 0.0)) bad = 1;
    ls = log(d);
  }
  double v = 0.0;
  if (kind == CVXB_OBJ_KL && tid < S.ncur) {
    double xj = S.x[tid] + s * S.dir[tid];
    v = xj * log(xj * (double)S.ncur);
  }
  double nbad = (double)bad;
  b_sum3(ls, nbad, v, S.red);
  *throws = nbad != 0.0;
  const double f0s = (kind == CVXB_OBJ_KL) ? v : f0 + s * c1 + 0.5 * s * s * c2;
  return t * f0s - ls;
}

struct LoopOut {
  int status, stage, total_steps;
  double objv, gap, eqgap;
};

// BarrierSolver.solveWithEQs / solveWithoutEQs (BarrierSolver.scala:70-177) on the problem currently in shared memory
// (S.ncur, S.mcur are ITS dimensions: the phase-I pass runs with n + 1 variables and m + 2p rows), from S.x.
//   term 0: standard termination (duality gap and equality gap below tolSolver);
//   term 1: phase I (CvxUtils.scala:78-87): objective value below zero -- a strictly feasible point has been found.
__device__ __noinline__ void b_barrier_loop(Smem& S, const BatchArgs& A, int kind, double obj_r, const double* Pg, double beq,
                                            int p, int term, double* Hs, const double* Gg, int p_orig, bool record_stages,
                                            LoopOut& out) {
  const int tid = threadIdx.x, n = S.ncur, m = S.mcur;
  const double tol = A.P.tolSolver, tolEq = A.P.tolEqSolve;
  int status = CVXB_OK, stage = 0, total_steps = 0;
  double t = A.P.t0, gap = 1.7976931348623157e308, eqgap = 1.7976931348623157e308;
  double objv = term ? 1.7976931348623157e308 : 0.0;
  const double maxStage = 1000.0 / A.P.mu;
  while (!(term ? (objv < 0.0 && eqgap < A.P.phase1EqTol) : (gap < tol && eqgap < tol)) && stage < maxStage &&
         status == CVXB_OK) {
    // ---------------- inner Newton solve at parameter t
    int iter = 0;
    double nd = tol + 1, fval, f0, normGrad, eqd, trust = 0.0;
    if (!b_eval(S, A, kind, obj_r, Pg, beq, t, &fval, &f0, &normGrad, &eqd)) { status = CVXB_ENOTFEASIBLE; break; }
    double eqnorm = fabs(eqd);
    while (iter < A.P.maxIter && (p ? ((nd > tol && normGrad > tol) || eqnorm > tol) : (nd > tol && normGrad > tol))) {
      BCLK(0);
      b_hessian(S, A, kind, Pg, t);
      BCLK(1);
      b_store_H(S, n, Hs);
      BCLK(2);
      int reg = 0;
      bool ok = b_linear_solve(S, A, Hs, 0.0, false, S.y, eqd, tolEq, &reg, true);
      if (!ok) {
        if (p) {
          // path 1: K = H + a a', z = q - a' b   (KKTSystem.scala:57-59)
          if (tid < n) S.zrhs[tid] = S.y[tid] - S.aeq[tid] * eqd;
          __syncthreads();
          ok = b_linear_solve(S, A, Hs, 0.0, true, S.zrhs, eqd, tolEq, &reg);
          if (!ok) {
            // path 2: decomposition of the full KKT matrix  (KKTSystem.scala:63, 283-310)
            ok = b_kkt_sym_solve(S, n, true, Hs, S.y, eqd, tolEq);
            b_reload_G(S, Gg, A.n, A.m, p_orig, false);
            if (!ok) { status = CVXB_EUNSOLVABLE; break; }
          }
        } else {
          ok = b_linear_solve(S, A, Hs, A.P.newtonRegDelta, false, S.y, 0.0, tolEq, &reg);   // H + 1e-9 I
          if (!ok) {
            // MatrixUtils.symSolve(H, -y)   (UnconstrainedSolver.scala:65)
            ok = b_kkt_sym_solve(S, n, false, Hs, S.y, 0.0, tolEq);
            b_reload_G(S, Gg, A.n, A.m, p_orig, term == 1);       // (the phase-I pass works on the widened G)
            if (!ok) { status = CVXB_EUNSOLVABLE; break; }
          }
        }
        if (!ok) { status = CVXB_ELINSOLVE; break; }
      }
      BCLK(7);
      // q = d . grad
      double q = 0.0, c1 = 0.0, c2 = 0.0;
      if (kind == CVXB_OBJ_QUADRATIC && tid >= 128 && tid < 128 + n) {
        const int i = tid - 128;
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(Pg[i + (size_t)j * n], S.dir[j], s);
        S.Pd[i] = s;
      }
      if (tid < m) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(GG(tid, j), S.dir[j], s);
        S.Gd[tid] = s;
      }
      __syncthreads();
      if (tid < n) {
        double dj = S.dir[tid];
        q = dj * S.y[tid];
        if (kind == CVXB_OBJ_LINEAR) c1 = S.oa[tid] * dj;
        else if (kind == CVXB_OBJ_QUADRATIC) { c1 = (S.oa[tid] + S.Px[tid]) * dj; c2 = dj * S.Pd[tid]; }
      }
      b_sum3(q, c1, c2, S.red);
      nd = -q / 2;
      bool moved = false;
      if (nd > tol) {
        int it = 0, thr = 0;
        double step;
        if (p) {
          double s = 1.0;
          while (!b_in_set(S, m, s) && it < 100) { s *= A.P.beta; ++it; }
          if (it == 100) { status = CVXB_ELINESEARCH; break; }
          while (it < 100) {
            double v = b_value(S, A, kind, t, s, f0, c1, c2, &thr);
            if (thr) break;
            if (!(v > fval + A.P.alpha * s * q)) break;
            s *= A.P.beta; ++it;
          }
          if (thr) { status = CVXB_ENOTFEASIBLE; break; }
          if (it == 100) { status = CVXB_ELINESEARCH; break; }
          step = s;
        } else {
          const double hnorm = sqrt(-q);
          if (iter == 0) trust = hnorm;
          const double scl = (iter == 0 || hnorm <= trust) ? 1.0 : trust / hnorm;
          double tt = 1.0;
          while (!b_in_set(S, m, scl * tt) && it < 200) { tt *= A.P.beta; ++it; }
          if (it == 100) { status = CVXB_ELINESEARCH; break; }
          if (b_in_set(S, m, scl)) {
            (void)b_value(S, A, kind, t, scl * tt, f0, c1, c2, &thr);
            if (thr) { status = CVXB_ENOTFEASIBLE; break; }
          }
          while (it < 200) {
            double v = b_value(S, A, kind, t, scl * tt, f0, c1, c2, &thr);
            if (thr) break;
            if (!(v > fval + A.P.alpha * tt * q)) break;
            tt *= A.P.beta; ++it;
          }
          if (thr) { status = CVXB_ENOTFEASIBLE; break; }
          if (it == 100) { status = CVXB_ELINESEARCH; break; }
          step = scl * tt;
        }
        BCLK(8);
        if (tid < n) S.x[tid] = S.x[tid] + S.dir[tid] * step;
        __syncthreads();
        if (!b_eval(S, A, kind, obj_r, Pg, beq, t, &fval, &f0, &normGrad, &eqd)) { status = CVXB_ENOTFEASIBLE; break; }
        eqnorm = fabs(eqd);
        moved = true;
        BCLK(9);
      }
      ++iter;
      if (!moved && p && ((nd > tol && normGrad > tol) || eqnorm > tol)) { iter = A.P.maxIter; break; }   // identical repeats
    }
    if (status != CVXB_OK) break;
    total_steps += iter;
    if (record_stages && tid == 0 && stage < CVXB_BATCH_STAGES) S.stage_steps[stage] = iter;
    objv = f0;
    gap = (double)m / t;
    eqgap = p ? eqnorm : 0.0;
    t *= A.P.mu;
    ++stage;
  }
  out.status = status; out.stage = stage; out.total_steps = total_steps;
  out.objv = objv; out.gap = gap; out.eqgap = eqgap;
}

__global__ void __launch_bounds__(BT, 2) batched_barrier_kernel(const __grid_constant__ BatchArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  __shared__ int s_pid;
  const int tid = threadIdx.x, n = A.n, m = A.m;
  double* Hs = A.scratch + (size_t)blockIdx.x * BN * BN;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_pid = (int)atomicAdd(A.counter, 1u);
    __syncthreads();
    if (s_pid >= A.B) break;
    // longest-expected-first pickup (cvxb_batch_create): the tail of the batch is then made of the short problems
    const int pid = A.order ? A.order[s_pid] : s_pid;
    const long long clk0 = clock64();
    // ---- load the problem
    const int kind = A.objective[pid];
    const int p = A.pcount ? A.pcount[pid] : A.p;
    if (tid == 0) { S.pcur = p; S.ncur = n; S.mcur = m; }
    const double obj_r = A.obj_r ? A.obj_r[pid] : 0.0;
    const double* Pg = A.obj_P ? A.obj_P + (size_t)pid * n * n : nullptr;
    const double* Gg = A.G + (size_t)pid * m * n;
    for (int idx = tid; idx < BN * LDG; idx += BT) S.G[idx] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < m * n; idx += BT) GG(idx % m, idx / m) = Gg[idx];
    if (tid < BM_) S.ub[tid] = tid < m ? A.ub[(size_t)pid * m + tid] : 1.0;
    if (tid < n) {
      S.x[tid] = A.x0[(size_t)pid * n + tid];
      S.oa[tid] = (kind != CVXB_OBJ_KL && A.obj_a) ? A.obj_a[(size_t)pid * n + tid] : 0.0;
      S.aeq[tid] = p ? A.A[(size_t)pid * n + tid] : 0.0;
      S.Px[tid] = 0.0; S.Pd[tid] = 0.0;
    }
    const double beq = p ? A.b[pid] : 0.0;
    __syncthreads();
    if (tid < CVXB_BATCH_STAGES) S.stage_steps[tid] = 0;

    LoopOut ph;
    ph.status = CVXB_OK; ph.stage = 0; ph.total_steps = 0; ph.objv = 0.0; ph.gap = 0.0; ph.eqgap = 0.0;
    double ph_s = 0.0;
    const bool need_phase1 = A.phase1 && A.phase1[pid];
    if (need_phase1) {
      // ---- phase I (ConstraintSet.phase_I_Analysis, ConstraintSet.scala:326-395, 556-575; Constraint.phase_I
      // Constraint.scala:64-89; EqualityConstraint.scala:84-100): minimise s over (x, s) subject to
      // g_i(x) - s <= ub_i and +-(a.x - b) - s <= phase1EqTol, built in place: one more column (-1) and two more rows
      // (+-a) of G, no equalities, linear objective s; start (x0, 1 + max_i (g_i(x0) - ub_i)).
      const int n1 = n + 1, m1 = m + 2 * p;
      b_phase1_G(S, n, m, p);
      if (tid == 0) {
        if (p) { S.ub[m] = beq + A.P.phase1EqTol; S.ub[m + 1] = -beq + A.P.phase1EqTol; }
        S.x[n] = 0.0;
        S.pcur = 0;
        S.ncur = n1;
        S.mcur = m1;
      }
      if (tid < n1) S.oa[tid] = (tid == n) ? 1.0 : 0.0;
      __syncthreads();
      double viol = -1e308;
      if (tid < m1) {
        double sx = 0.0;
        for (int j = 0; j < n; ++j) sx = fma(GG(tid, j), S.x[j], sx);
        viol = sx - S.ub[tid];
      }
      viol = -block_min(-viol, S.red);
      if (tid == 0) S.x[n] = 1.0 + viol;
      __syncthreads();
      b_barrier_loop(S, A, CVXB_OBJ_LINEAR, 0.0, nullptr, 0.0, 0, 1, Hs, Gg, p, false, ph);
      __syncthreads();
      ph_s = S.x[n];
      if (ph.status == CVXB_OK && !(ph_s < A.P.tolSolver)) ph.status = CVXB_EINFEASIBLE;   // FeasibilityReport.isFeasible(tol)
      __syncthreads();
      // back to the problem itself, from the feasible point found
      if (tid < m1) GG(tid, n) = 0.0;
      if (p && tid < n1) { GG(m, tid) = 0.0; GG(m + 1, tid) = 0.0; }
      if (tid == 0) {
        if (p) { S.ub[m] = 1.0; S.ub[m + 1] = 1.0; }
        S.pcur = p;
        S.ncur = n;
        S.mcur = m;
        S.x[n] = 0.0;
      }
      if (tid < n1) S.oa[tid] = (tid < n && kind != CVXB_OBJ_KL && A.obj_a) ? A.obj_a[(size_t)pid * n + tid] : 0.0;
      if (tid < BN) { S.Px[tid] = 0.0; S.Pd[tid] = 0.0; }
      __syncthreads();
    }
    LoopOut r;
    if (ph.status == CVXB_OK) {
      b_barrier_loop(S, A, kind, obj_r, Pg, beq, p, 0, Hs, Gg, p, true, r);
    } else {
      r = ph;
      r.stage = 0; r.total_steps = 0; r.objv = 0.0;
      r.gap = 1.7976931348623157e308; r.eqgap = 1.7976931348623157e308;
    }
    __syncthreads();
    const int status = r.status, total_steps = r.total_steps, stage = r.stage;
    const double objv = r.objv, gap = r.gap, eqgap = r.eqgap;
    if (tid < n) A.x[(size_t)pid * n + tid] = S.x[tid];
    {   // the packed record the multi-GPU gather ships (one all-gather, no host staging)
      double* rec = A.records + (size_t)pid * (n + CVXB_BATCH_RECORD_EXTRA);
      if (tid < n) rec[tid] = S.x[tid];
      if (tid == 0) {
        rec[n] = objv; rec[n + 1] = gap; rec[n + 2] = (double)status; rec[n + 3] = (double)total_steps;
        rec[n + 4] = (double)stage;
      }
    }
    __syncthreads();
    if (tid < CVXB_BATCH_STAGES) A.stage_steps[(size_t)pid * CVXB_BATCH_STAGES + tid] = S.stage_steps[tid];
    if (tid == 0) {
      A.cycles[pid] = clock64() - clk0;
      A.status[pid] = status;
      A.steps[pid] = total_steps;
      A.stages[pid] = stage;
      A.objval[pid] = objv;
      A.gap[pid] = gap;
      A.eqgap[pid] = p ? eqgap : 0.0;
      if (A.ph_steps) { A.ph_steps[pid] = ph.total_steps; A.ph_stages[pid] = ph.stage; A.ph_s[pid] = ph_s; }
    }
  }
}

template <typename T>
int balloc(cvxb_batch_s* Bt, T** ptr, size_t count) {
  void* q = nullptr;
  CVXB_CUDA_OK(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
  Bt->owned.push_back(q);
  *ptr = (T*)q;
  return CVXB_OK;
}

template <typename T>
int bupload(cvxb_batch_s* Bt, T** dst, const T* src, size_t count) {
  if (!src) { *dst = nullptr; return CVXB_OK; }
  CVXB_TRY(balloc(Bt, dst, count));
  bool dev = (Bt->h->flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  CVXB_CUDA_OK(cudaMemcpyAsync(*dst, src, count * sizeof(T), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               Bt->h->stream));
  return CVXB_OK;
}

}  // namespace
}  // namespace cvxb

extern "C" {

int cvxb_batch_create(cvxb_handle h, const cvxb_batch_desc* d, cvxb_batch* out) {
  if (!h || !d || !out) { cvxb::set_last_error("cvxb_batch_create: null argument"); return CVXB_EINVAL; }
  if (d->B < 1 || d->n < 1 || d->n > BN || d->m < 1 || d->m > BM_ || d->p < 0 || d->p > 1) {
    cvxb::set_last_error("cvxb_batch_create: need B >= 1, 1 <= n <= 64, 1 <= m <= 128, p in {0,1} (got %d, %d, %d, %d)", d->B,
                         d->n, d->m, d->p);
    return CVXB_EDIM;
  }
  if (!d->objective || !d->G || !d->ub || !d->x0 || (d->p && (!d->A || !d->b))) {
    cvxb::set_last_error("cvxb_batch_create: missing array");
    return CVXB_EINVAL;
  }
  if (!(h->flags & CVXB_FLAG_DEVICE_PTRS)) {
    // host descriptors: refuse what the kernel has no branch for instead of solving something else silently
    for (int i = 0; i < d->B; ++i) {
      const int k = d->objective[i];
      if (k != CVXB_OBJ_LINEAR && k != CVXB_OBJ_QUADRATIC && k != CVXB_OBJ_KL) {
        cvxb::set_last_error("cvxb_batch_create: problem %d has objective kind %d; the batched solver handles LINEAR, "
                             "QUADRATIC and KL", i, k);
        return CVXB_ENOTIMPL;
      }
      if ((k == CVXB_OBJ_LINEAR || k == CVXB_OBJ_QUADRATIC) && (!d->obj_a || !d->obj_r)) {
        cvxb::set_last_error("cvxb_batch_create: problem %d is linear / quadratic but obj_a or obj_r is NULL", i);
        return CVXB_EINVAL;
      }
      if (k == CVXB_OBJ_QUADRATIC && !d->obj_P) {
        cvxb::set_last_error("cvxb_batch_create: problem %d is quadratic but obj_P is NULL", i);
        return CVXB_EINVAL;
      }
      if (d->pcount && (d->pcount[i] < 0 || d->pcount[i] > d->p)) {
        cvxb::set_last_error("cvxb_batch_create: pcount[%d] = %d outside 0..p = %d", i, d->pcount[i], d->p);
        return CVXB_EDIM;
      }
      if (d->phase1 && d->phase1[i]) {
        const int pi = d->pcount ? d->pcount[i] : d->p;
        if (d->n + 1 > BN || d->m + 2 * pi > BM_) {
          cvxb::set_last_error("cvxb_batch_create: problem %d asks for phase I, whose feasibility problem has n + 1 = %d variables "
                               "and m + 2p = %d rows; the kernel holds 64 x 128", i, d->n + 1, d->m + 2 * pi);
          return CVXB_EDIM;
        }
      }
    }
  } else if (d->phase1 && (d->n + 1 > BN || d->m + 2 * d->p > BM_)) {
    cvxb::set_last_error("cvxb_batch_create: phase I needs n <= 63 and m + 2p <= 128");
    return CVXB_EDIM;
  }
  cvxb::DeviceGuard _guard(h->device);
  cvxb_batch_s* Bt = new cvxb_batch_s();
  Bt->h = h; Bt->B = d->B; Bt->n = d->n; Bt->m = d->m; Bt->p = d->p;
  const size_t B = d->B, n = d->n, m = d->m;
  int st = CVXB_OK;
  auto T = [&](int s) { if (st == CVXB_OK) st = s; };
  T(bupload(Bt, &Bt->objective, d->objective, B));
  T(bupload(Bt, &Bt->pcount, d->pcount, B));
  T(bupload(Bt, &Bt->obj_a, d->obj_a, B * n));
  T(bupload(Bt, &Bt->obj_r, d->obj_r, B));
  T(bupload(Bt, &Bt->obj_P, d->obj_P, B * n * n));
  T(bupload(Bt, &Bt->G, d->G, B * m * n));
  T(bupload(Bt, &Bt->ub, d->ub, B * m));
  if (d->p) { T(bupload(Bt, &Bt->A, d->A, B * n)); T(bupload(Bt, &Bt->b, d->b, B)); }
  T(bupload(Bt, &Bt->x0, d->x0, B * n));
  T(balloc(Bt, &Bt->x, B * n)); T(balloc(Bt, &Bt->objval, B)); T(balloc(Bt, &Bt->gap, B)); T(balloc(Bt, &Bt->eqgap, B));
  T(balloc(Bt, &Bt->status, B)); T(balloc(Bt, &Bt->steps, B)); T(balloc(Bt, &Bt->stages, B));
  T(balloc(Bt, &Bt->records, B * (n + CVXB_BATCH_RECORD_EXTRA)));
  T(balloc(Bt, &Bt->stage_steps, B * CVXB_BATCH_STAGES));
  T(balloc(Bt, &Bt->cycles, B));
  if (d->phase1) {
    T(bupload(Bt, &Bt->phase1, d->phase1, B));
    T(balloc(Bt, &Bt->ph_steps, B)); T(balloc(Bt, &Bt->ph_stages, B)); T(balloc(Bt, &Bt->ph_s, B));
  }
  if (!(h->flags & CVXB_FLAG_DEVICE_PTRS) && !getenv("CVXB_BATCH_INDEX_ORDER")) {
    // Pickup order: problems are handed to the CTAs through one atomic ticket counter; dealing out the long ones
    // first (quadratic objectives: ~100 Newton steps each on the BASELINE mix, against ~70 for the KL problems, and a
    // P x product per evaluation) leaves only short problems for the last, partly filled wave.
    std::vector<int> ord((size_t)d->B);
    size_t k = 0;
    for (int pass = 0; pass < 3; ++pass) {
      const int want = pass == 0 ? CVXB_OBJ_QUADRATIC : (pass == 1 ? CVXB_OBJ_LINEAR : CVXB_OBJ_KL);
      for (int i = 0; i < d->B; ++i)
        if (d->objective[i] == want) ord[k++] = i;
    }
    T(bupload(Bt, &Bt->order, ord.data(), B));
  }
  Bt->grid = h->sm_count * 2;
  if (const char* e = getenv("CVXB_BATCH_CTAS_PER_SM")) Bt->grid = h->sm_count * (atoi(e) > 0 ? atoi(e) : 2);   // residency experiments
  if (Bt->grid > d->B) Bt->grid = d->B;
  T(balloc(Bt, &Bt->scratch, (size_t)Bt->grid * BN * BN));
  T(balloc(Bt, &Bt->counter, 1));
  if (st == CVXB_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) st = CVXB_ECUDA;
  if (st != CVXB_OK) { for (void* q : Bt->owned) cudaFree(q); delete Bt; return st; }
  *out = Bt;
  return CVXB_OK;
}

int cvxb_debug_batch_clocks(long long* out, int reset) {
  if (out) CVXB_CUDA_OK(cudaMemcpyFromSymbol(out, g_batch_clk, 12 * sizeof(long long)));
  if (out && getenv("CVXB_POTRF_CLOCKS")) CVXB_CUDA_OK(cudaMemcpyFromSymbol(out, g_potrf_clk, 8 * sizeof(long long)));
  if (reset) {
    long long z[12] = {0};
    CVXB_CUDA_OK(cudaMemcpyToSymbol(g_batch_clk, z, sizeof(z)));
    CVXB_CUDA_OK(cudaMemcpyToSymbol(g_potrf_clk, z, 8 * sizeof(long long)));
  }
  return CVXB_OK;
}

int cvxb_batch_device_records(cvxb_batch Bt, double** dev_records, int* row_doubles) {
  if (!Bt || !dev_records) { cvxb::set_last_error("cvxb_batch_device_records: null argument"); return CVXB_EINVAL; }
  *dev_records = Bt->records;
  if (row_doubles) *row_doubles = Bt->n + CVXB_BATCH_RECORD_EXTRA;
  return CVXB_OK;
}

int cvxb_batch_destroy(cvxb_batch Bt) {
  if (!Bt) return CVXB_OK;
  cvxb::DeviceGuard _guard(Bt->h->device);
  cudaStreamSynchronize(Bt->h->stream);
  for (void* q : Bt->owned) cudaFree(q);
  delete Bt;
  return CVXB_OK;
}

int cvxb_batch_barrier_solve(cvxb_handle h, cvxb_batch Bt, const cvxb_params* pars, cvxb_batch_result* out) {
  if (!h || !Bt || !out) { cvxb::set_last_error("cvxb_batch_barrier_solve: null argument"); return CVXB_EINVAL; }
  if (Bt->h != h) { cvxb::set_last_error("batch belongs to another handle"); return CVXB_EINVAL; }
  cvxb::DeviceGuard _guard(h->device);
  cvxb_params dp;
  if (!pars) { cvxb_default_params(&dp); pars = &dp; }
  CVXB_CUDA_OK(cudaFuncSetAttribute(batched_barrier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  BatchArgs A;
  A.B = Bt->B; A.n = Bt->n; A.m = Bt->m; A.p = Bt->p;
  A.objective = Bt->objective; A.pcount = Bt->pcount; A.obj_a = Bt->obj_a; A.obj_r = Bt->obj_r; A.obj_P = Bt->obj_P; A.G = Bt->G; A.ub = Bt->ub;
  A.A = Bt->A; A.b = Bt->b; A.x0 = Bt->x0;
  A.x = Bt->x; A.objval = Bt->objval; A.gap = Bt->gap; A.eqgap = Bt->eqgap;
  A.status = Bt->status; A.steps = Bt->steps; A.stages = Bt->stages; A.records = Bt->records; A.stage_steps = Bt->stage_steps;
  A.order = Bt->order; A.cycles = Bt->cycles;
  A.phase1 = Bt->phase1; A.ph_steps = Bt->ph_steps; A.ph_stages = Bt->ph_stages; A.ph_s = Bt->ph_s;
  A.scratch = Bt->scratch; A.counter = Bt->counter; A.P = *pars;
  cvxb::NvtxRange nvtx("cvxb batched barrier solve");
  CVXB_CUDA_OK(cudaMemsetAsync(Bt->counter, 0, sizeof(unsigned), h->stream));
  CVXB_CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  batched_barrier_kernel<<<Bt->grid, BT, sizeof(Smem), h->stream>>>(A);
  h->launches++;
  CVXB_CUDA_OK(cudaGetLastError());
  CVXB_CUDA_OK(cudaEventRecord(h->ev1, h->stream));
  bool dev = (h->flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  cudaMemcpyKind k = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const size_t B = Bt->B, n = Bt->n;
  if (out->x) CVXB_CUDA_OK(cudaMemcpyAsync(out->x, Bt->x, B * n * sizeof(double), k, h->stream));
  if (out->status) CVXB_CUDA_OK(cudaMemcpyAsync(out->status, Bt->status, B * sizeof(int), k, h->stream));
  if (out->newton_steps) CVXB_CUDA_OK(cudaMemcpyAsync(out->newton_steps, Bt->steps, B * sizeof(int), k, h->stream));
  if (out->outer_stages) CVXB_CUDA_OK(cudaMemcpyAsync(out->outer_stages, Bt->stages, B * sizeof(int), k, h->stream));
  if (out->cycles) CVXB_CUDA_OK(cudaMemcpyAsync(out->cycles, Bt->cycles, B * sizeof(long long), k, h->stream));
  if (out->stage_newton_steps)
    CVXB_CUDA_OK(cudaMemcpyAsync(out->stage_newton_steps, Bt->stage_steps, B * CVXB_BATCH_STAGES * sizeof(int), k, h->stream));
  if (out->objective) CVXB_CUDA_OK(cudaMemcpyAsync(out->objective, Bt->objval, B * sizeof(double), k, h->stream));
  if (out->duality_gap) CVXB_CUDA_OK(cudaMemcpyAsync(out->duality_gap, Bt->gap, B * sizeof(double), k, h->stream));
  if (out->equality_gap) CVXB_CUDA_OK(cudaMemcpyAsync(out->equality_gap, Bt->eqgap, B * sizeof(double), k, h->stream));
  if (Bt->ph_steps) {
    if (out->phase1_newton_steps) CVXB_CUDA_OK(cudaMemcpyAsync(out->phase1_newton_steps, Bt->ph_steps, B * sizeof(int), k, h->stream));
    if (out->phase1_stages) CVXB_CUDA_OK(cudaMemcpyAsync(out->phase1_stages, Bt->ph_stages, B * sizeof(int), k, h->stream));
    if (out->phase1_s) CVXB_CUDA_OK(cudaMemcpyAsync(out->phase1_s, Bt->ph_s, B * sizeof(double), k, h->stream));
  } else if (!dev) {
    if (out->phase1_newton_steps) memset(out->phase1_newton_steps, 0, B * sizeof(int));
    if (out->phase1_stages) memset(out->phase1_stages, 0, B * sizeof(int));
    if (out->phase1_s) memset(out->phase1_s, 0, B * sizeof(double));
  }
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CVXB_CUDA_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  out->solve_ms = ms;
  return CVXB_OK;
}

}  // extern "C"

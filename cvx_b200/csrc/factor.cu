// Equilibration, Cholesky and triangular solves of the KKT path.
//   ruiz_equilibrate   MatrixUtils.ruizEquilibrate        MatrixUtils.scala:240-268
//   potrf_lower        Breeze cholesky -> LAPACK dpotrf   MatrixUtils.scala:452-461, KKTSystem.scala:140
//   trsm_lower         LAPACK dtrtrs / forwardSolve / backSolve   MatrixUtils.scala:362-430
// The factorisation is the recursive (cache-oblivious) right-looking Cholesky: every flop outside the
// NB x NB diagonal leaves is a large-K DMMA GEMM/SYRK (gemm_dmma.cu).  A leaf CTA factors its block in
// shared memory and also writes the block's triangular inverse, so that every leaf triangular solve is
// a DMMA GEMM with the inverse too (same device as MAGMA's trsm with inverted diagonal blocks).
#include "common.cuh"
#include <cooperative_groups.h>

namespace cvxb {

int potrf_lookahead(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, int col0,
                    double* B = nullptr, int ldb = 0, int r = 0);
int rl_max_n();

namespace {

int trsm_rec(Handle& h, int n, int r, const double* L, int ldl, const double* invD, double* B, int ldb, bool trans);

// ------------------------------------------------------------------------------------------- Ruiz
constexpr int RUIZ_MAX_FUSED = 64;    // rho slots behind the ticket words (api.cu allocates 16 unsigned + 64 u64)
__global__ void ruiz_init_kernel(int n, double* d, int* flag, double* scal, unsigned* ticket) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = 1.0;
  if (i == 0) {
    flag[F_RUIZ_DONE] = 0;
    flag[F_RUIZ_SWEEPS] = 0;
    scal[S_RUIZ_RHO] = 1.0;
    ticket[0] = 0;
  }
  if (i < RUIZ_MAX_FUSED) ((unsigned long long*)(ticket + 16))[i] = 0ull;     // per-sweep rho slots of the fused kernel
}

// The whole equilibration in ONE cooperative launch: CTA per column (128 threads, 16-byte loads, four column
// chunks and their d entries in flight per thread), u_j / the new d_j computed by the column's own CTA, rho by
// atomicMax on the bit pattern (non-negative doubles order like unsigned integers; a NaN sorts above +inf and
// ends the loop exactly like the per-sweep kernel's `!(rho > tol)`), d double-buffered between `d` and `dalt`
// so that a sweep costs one grid barrier.  Converged early -> the kernel returns: no no-op launches.
__global__ void __launch_bounds__(128) ruiz_fused_kernel(int n, const double* __restrict__ Hm, int ldh, double* d, double* dalt,
                                                         unsigned long long* rho_bits, int* flag, double* scal,
                                                         int max_sweeps, double tol) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* din = d;
  double* dout = dalt;
  int sweeps = 0;
  double rho = 1.0;
  const int npair = n >> 1;
  for (int s = 0; s < max_sweeps; ++s) {
    double rmax = 0.0;
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
      const double2* col = reinterpret_cast<const double2*>(Hm + (size_t)j * ldh);
      const double2* dv = reinterpret_cast<const double2*>(din);
      const double dj = __ldcg(din + j);       // written by other CTAs in the previous sweep: read through L2
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      int i = tid;
      for (; i + 384 < npair; i += 512) {
        const double2 h0 = col[i], h1 = col[i + 128], h2 = col[i + 256], h3 = col[i + 384];
        const double2 e0 = __ldcg(dv + i), e1 = __ldcg(dv + i + 128), e2 = __ldcg(dv + i + 256), e3 = __ldcg(dv + i + 384);
        double q;
        q = (e0.x * dj) * h0.x; s0 = fma(q, q, s0); q = (e0.y * dj) * h0.y; s0 = fma(q, q, s0);
        q = (e1.x * dj) * h1.x; s1 = fma(q, q, s1); q = (e1.y * dj) * h1.y; s1 = fma(q, q, s1);
        q = (e2.x * dj) * h2.x; s2 = fma(q, q, s2); q = (e2.y * dj) * h2.y; s2 = fma(q, q, s2);
        q = (e3.x * dj) * h3.x; s3 = fma(q, q, s3); q = (e3.y * dj) * h3.y; s3 = fma(q, q, s3);
      }
      for (; i < npair; i += 128) {
        const double2 h0 = col[i];
        const double2 e0 = __ldcg(dv + i);
        double q;
        q = (e0.x * dj) * h0.x; s0 = fma(q, q, s0); q = (e0.y * dj) * h0.y; s0 = fma(q, q, s0);
      }
      if ((n & 1) && tid == 0) {
        const double q = (__ldcg(din + n - 1) * dj) * Hm[(size_t)j * ldh + n - 1];
        s1 = fma(q, q, s1);
      }
      double sum = (s0 + s1) + (s2 + s3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) red[warp] = sum;
      __syncthreads();
      if (tid == 0) {
        const double tot = (red[0] + red[1]) + (red[2] + red[3]);
        const double u = sqrt(sqrt(tot));
        dout[j] = u > 0 ? dj * (1.0 / u) : dj;
        const double a = fabs(1.0 - u);
        if (a > rmax || a != a) rmax = a;
      }
      __syncthreads();
    }
    if (tid == 0 && blockIdx.x < n) atomicMax(rho_bits + s, (unsigned long long)__double_as_longlong(rmax));
    grid.sync();
    rho = __longlong_as_double((long long)__ldcg(rho_bits + s));
    double* tmp = din; din = dout; dout = tmp;
    sweeps = s + 1;
    if (!(rho > tol)) break;
  }
  if (din != d)       // odd number of sweeps: the result sits in the alternate buffer (every read of d is behind a barrier)
    for (int j = blockIdx.x * blockDim.x + tid; j < n; j += gridDim.x * blockDim.x) d[j] = __ldcg(din + j);
  if (blockIdx.x == 0 && tid == 0) {
    scal[S_RUIZ_RHO] = rho;
    flag[F_RUIZ_SWEEPS] = sweeps;
    flag[F_RUIZ_DONE] = 1;
  }
}

// The same equilibration for matrices far beyond L2 (n >= 4096: every sweep is one HBM pass), reading only the LOWER
// triangle: H is symmetric, so a 128 x 128 tile (bi >= bj) yields the partial column sums of its 128 columns AND, through its
// row sums, the partial sums of the columns bi*128.. that its mirror image would have contributed.  Phase A: tiles are
// dealt out over the persistent CTAs (warp w owns 16 columns of the tile, lane l the rows l, l+32, l+64, l+96: coalesced
// 256-byte segments, 16 loads in flight per thread); the tile's column sums go to colpart[bi][j], its row sums (summed
// over the 8 warps in a fixed order) to rowpart[bj][i] -- every partial has exactly one writer and the final sum runs in a
// fixed order, so the result is deterministic (no floating-point atomics).  Phase B: thread per column adds the partials
// of its block column and block row, computes u_j, the new d_j and the warp's max |1 - u|.  Two grid barriers per sweep
// instead of one, half the bytes: 69 -> ~40 us per sweep at n = 8192.
constexpr int RS_T = 128;
__global__ void __launch_bounds__(256, 4) ruiz_sym_kernel(int n, const double* __restrict__ Hm, int ldh, double* d, double* dalt,
                                                       double* part, int ldp, unsigned long long* rho_bits, int* flag,
                                                       double* scal, int max_sweeps, double tol) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[8][RS_T];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = (n + RS_T - 1) / RS_T;
  const int T = nt * (nt + 1) / 2;
  double* colpart = part;
  double* rowpart = part + (size_t)nt * ldp;
  double* din = d;
  double* dout = dalt;
  int sweeps = 0;
  double rho = 1.0;
  for (int s = 0; s < max_sweeps; ++s) {
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
      int bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while ((long long)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
      while ((long long)bi * (bi + 1) / 2 > t) --bi;
      const int bj = t - bi * (bi + 1) / 2;
      const int i0 = bi * RS_T, j0 = bj * RS_T;
      const bool diag = bi == bj;
      int ri[4];
      double di[4], racc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ri[k] = i0 + lane + 32 * k;
        di[k] = ri[k] < n ? __ldcg(din + ri[k]) : 0.0;
        racc[k] = 0.0;
      }
      // the d entries of this warp's 16 columns: one coalesced load, handed out by shuffle
      const int jw = j0 + 16 * warp;
      const double djl = (lane < 16 && jw + lane < n) ? __ldcg(din + jw + lane) : 0.0;
#pragma unroll
      for (int grp = 0; grp < 4; ++grp) {
        // the 16 loads of 4 columns x 4 rows are issued before the first value is used
        double hv[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = jw + 4 * grp + c;
          const double* col = Hm + (size_t)j * ldh;
#pragma unroll
          for (int k = 0; k < 4; ++k) hv[c][k] = (j < n && ri[k] < n && (!diag || ri[k] >= j)) ? col[ri[k]] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = jw + 4 * grp + c;
          const double dj = __shfl_sync(0xffffffffu, djl, 4 * grp + c);
          double v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double q = (di[k] * dj) * hv[c][k];
            v[k] = q * q;
          }
          double cs = (v[0] + v[1]) + (v[2] + v[3]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
          if (lane == 0 && j < n) colpart[(size_t)bi * ldp + j] = cs;
#pragma unroll
          for (int k = 0; k < 4; ++k) racc[k] += (diag && ri[k] == j) ? 0.0 : v[k];      // a diagonal element counts once
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) red[warp][lane + 32 * k] = racc[k];
      __syncthreads();
      if (tid < RS_T) {
        double r = ((red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid])) +
                   ((red[4][tid] + red[5][tid]) + (red[6][tid] + red[7][tid]));
        if (i0 + tid < n) rowpart[(size_t)bj * ldp + i0 + tid] = r;
      }
      __syncthreads();
    }
    grid.sync();
    double rmax = 0.0;
    for (int j = blockIdx.x * 256 + tid; j < n; j += gridDim.x * 256) {
      const int b = j / RS_T;
      double tot = 0.0;
      for (int r = b; r < nt; ++r) tot += __ldcg(colpart + (size_t)r * ldp + j);
      for (int c = 0; c <= b; ++c) tot += __ldcg(rowpart + (size_t)c * ldp + j);
      const double u = sqrt(sqrt(tot));
      const double dj = __ldcg(din + j);
      dout[j] = u > 0 ? dj * (1.0 / u) : dj;
      const double a = fabs(1.0 - u);
      if (a > rmax || a != a) rmax = a;
    }
    unsigned long long bits = (unsigned long long)__double_as_longlong(rmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
      bits = other > bits ? other : bits;
    }
    if (lane == 0 && bits != 0ull) atomicMax(rho_bits + s, bits);
    grid.sync();
    rho = __longlong_as_double((long long)__ldcg(rho_bits + s));
    double* tmp = din; din = dout; dout = tmp;
    sweeps = s + 1;
    if (!(rho > tol)) break;
  }
  if (din != d)
    for (int j = blockIdx.x * blockDim.x + tid; j < n; j += gridDim.x * blockDim.x) d[j] = __ldcg(din + j);
  if (blockIdx.x == 0 && tid == 0) {
    scal[S_RUIZ_RHO] = rho;
    flag[F_RUIZ_SWEEPS] = sweeps;
    flag[F_RUIZ_DONE] = 1;
  }
}

// One sweep: colsq[j] = sum_i ((d_i d_j) H_ij)^2 (H symmetric: column norm == row norm), then the
// last CTA to finish updates d and rho (MatrixUtils.scala:252-262).  A finished equilibration makes
// the remaining enqueued sweeps no-ops, so the host never has to look at rho.
__global__ void __launch_bounds__(256) ruiz_sweep_kernel(int n, const double* __restrict__ Hm, int ldh, double* d,
                                                         double* colsq, int* flag, double* scal, unsigned* ticket,
                                                         int max_sweeps, double tol) {
  if (flag[F_RUIZ_DONE]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  if (j < n) {
    const double* col = Hm + (size_t)j * ldh;
    const double dj = d[j];
    // four independent loads per lane in flight: the sweep is an L2-latency-bound stream of one column per warp
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = lane;
    for (; i + 96 < n; i += 128) {
      const double h0 = col[i], h1 = col[i + 32], h2 = col[i + 64], h3 = col[i + 96];
      const double q0 = (d[i] * dj) * h0;
      const double q1 = (d[i + 32] * dj) * h1;
      const double q2 = (d[i + 64] * dj) * h2;
      const double q3 = (d[i + 96] * dj) * h3;
      s0 = fma(q0, q0, s0);
      s1 = fma(q1, q1, s1);
      s2 = fma(q2, q2, s2);
      s3 = fma(q3, q3, s3);
    }
    for (; i < n; i += 32) {
      double q0 = (d[i] * dj) * col[i];
      s0 = fma(q0, q0, s0);
    }
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) colsq[j] = s;
  }
  __shared__ bool last;
  __shared__ double red[256];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double rho = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double u = sqrt(sqrt(((volatile double*)colsq)[k]));
    if (u > 0) d[k] = d[k] * (1.0 / u);
    double a = fabs(1.0 - u);
    if (a > rho || a != a) rho = a;
  }
  red[threadIdx.x] = rho;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double a = red[threadIdx.x + o];
      if (a > red[threadIdx.x] || a != a) red[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double r = red[0];
    scal[S_RUIZ_RHO] = r;
    int sw = flag[F_RUIZ_SWEEPS] + 1;
    flag[F_RUIZ_SWEEPS] = sw;
    if (!(r > tol) || sw >= max_sweeps) flag[F_RUIZ_DONE] = 1;
    ticket[0] = 0;
  }
}

__global__ void scaled_lower_kernel(int n, const double* __restrict__ Hm, int ldh, const double* __restrict__ d,
                                    double delta, double* __restrict__ L, int ldl, int full) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double di = d ? d[i] : 1.0;
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    double v = 0.0;
    if (full || i >= j) {
      v = d ? (di * d[j]) * Hm[(size_t)j * ldh + i] : Hm[(size_t)j * ldh + i];
      if (i == j) v += delta;
    }
    L[(size_t)j * ldl + i] = v;
  }
}

// ------------------------------------------------------------------------------------------- leaf
constexpr int LDW = NB + 4;   // stride = 4 mod 16 doubles: the DMMA fragment loads of the leaf (8 rows x 4 k) are bank-conflict
                              // free; the few row-strided sweeps (final store of the inverse) pay an 8-way conflict once
constexpr int LEAF_SMEM = (NB + 1) * LDW * (int)sizeof(double);   // L (lower) + inverse (transposed, upper, shifted)
constexpr int LEAF_THREADS = 512;
constexpr int LEAF_WARPS = LEAF_THREADS / 32;

// W holds L(i,j), i >= j, at W[i + j*LDW]; the inverse X(i,j), i >= j, at W[j + (i+1)*LDW].
#define LW(i, j) W[(i) + (j) * LDW]
#define XW(i, j) W[(j) + ((i) + 1) * LDW]

constexpr int SUB = 16;      // diagonal sub-block factored inside one warp

__device__ __forceinline__ void dmma_leaf(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
constexpr int NACC = 8;      // outputs per thread in the widest inverse-assembly level (128*64/2 / 512)

// 1/sqrt(d) for a pivot: single-precision MUFU.RSQ seed (2^-22) + ONE third-order correction
//   y = y0 (1 + e/2 + 3e^2/8),  e = 1 - d y0^2      (error ~ e^3 < 1e-19)
// Five dependent FP64 operations instead of the ~13 of rsqrt() + refinement: the 128 pivots of a leaf
// are a serial chain of FP64-latency-bound operations, so this is the critical path of the factorisation.
__device__ __forceinline__ double pivot_rsqrt(double d) {
  if (d < 1e-30 || d > 1e30) return rsqrt(d);          // outside float range: library path (rare, warp-uniform)
  const double y0 = (double)rsqrtf((float)d);
  const double t = d * y0;
  const double e = fma(-t, y0, 1.0);
  double p = fma(e, 0.375, 0.5);
  p = p * e;
  return fma(y0, p, y0);
}

// In-warp Cholesky of one SUB x SUB diagonal sub-block at offset o (bs valid rows): lane i < SUB holds
// row i in registers, pivots travel by shuffle, no block barrier inside.  The pivot uses rsqrt plus one
// Newton correction instead of IEEE sqrt + divide (each a ~400-cycle dependent chain, and there are 128
// pivots on the critical path of a leaf); reciprocal pivots go to rdiag for every later solve.
__device__ __forceinline__ void warp_diag_factor(double* W, double* rdiag, int o, int bs, int col0, int* s_fail,
                                                 double* s_mind) {
  const int lane = threadIdx.x & 31;
  double a[SUB], rr[SUB];
#pragma unroll
  for (int c = 0; c < SUB; ++c)
    a[c] = (lane < bs && c < bs && c <= lane) ? LW(o + lane, o + c) : ((c == lane) ? 1.0 : 0.0);
  double mind = 1e300;
  int fail = 0;
  // Straight-line, branch-free sweep (selects only): a divergent branch next to a full-mask shuffle costs a
  // reconvergence barrier per shuffle.  Entries above the diagonal (lane < c) are never read, so the
  // updates run unpredicated.
#pragma unroll
  for (int j = 0; j < SUB; ++j) {
    double d = __shfl_sync(0xffffffffu, a[j], j);
    const bool bad = !(d > 0.0) || d > 1e300;        // dpotrf: ajj <= 0 or NaN (inf would poison rsqrt)
    fail = (bad && !fail && j < bs) ? (col0 + o + j + 1) : fail;
    d = bad ? 1.0 : d;
    const double r = pivot_rsqrt(d);                 // 1/sqrt(d), full double accuracy, 5 dependent FP64 ops
    const double l = d * r;                          // sqrt(d) to ~1 ulp
    mind = (j < bs && l < mind) ? l : mind;
    rr[j] = r;
    a[j] = (lane == j) ? l : a[j] * r;
#pragma unroll
    for (int c = j + 1; c < SUB; ++c) {
      const double t = __shfl_sync(0xffffffffu, a[j], c);
      a[c] = fma(-a[j], t, a[c]);
    }
  }
  if (lane == 0) {
    if (fail && !*s_fail) *s_fail = fail;
    if (mind < *s_mind) *s_mind = mind;
  }
#pragma unroll
  for (int c = 0; c < SUB; ++c) {
    if (lane < bs && c <= lane) LW(o + lane, o + c) = a[c];
    if (lane == c) rdiag[o + c] = rr[c];
  }
}

// inverse of the diagonal sub-block at o: lane j holds column j of X = L^-1 (forward substitution,
// L(i,k) broadcast from shared memory, reciprocal pivots from rdiag)
__device__ __forceinline__ void warp_diag_inverse(double* W, const double* rdiag, int o, int bs) {
  const int lane = threadIdx.x & 31;
  // right-looking: x[i] is final after one multiply, then folded into every later row (short dependent chain)
  double x[SUB];
#pragma unroll
  for (int i = 0; i < SUB; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < SUB; ++i) {
    if (i < bs) {
      x[i] = (i >= lane) ? x[i] * rdiag[o + i] : 0.0;
#pragma unroll
      for (int i2 = i + 1; i2 < SUB; ++i2)
        if (i2 < bs) x[i2] = fma(-LW(o + i2, o + i), x[i], x[i2]);
    } else {
      x[i] = (i >= lane) ? x[i] : 0.0;
    }
  }
#pragma unroll
  for (int i = 0; i < SUB; ++i)
    if (i < bs && lane < bs && i >= lane) XW(o + i, o + lane) = x[i];
}

// One doubling level of the triangular inverse for ONE pair of S2-blocks at offset b, done by a single warp (full
// leaf, no bounds): T = L21 X11 into the X21 slot, then X21 = -X22 T.  Used inside the factorisation windows, where
// the other warps are busy; only __syncwarp is needed between the products.
template <int S2>
__device__ __forceinline__ void warp_level_job(double* W, int b) {
  constexpr int TPS = S2 / 8, NT = TPS * TPS;
  const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  double c0[NT], c1[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int r = b + S2 + 8 * (t / TPS), c = b + 8 * (t % TPS);
    c0[t] = c1[t] = 0.0;
#pragma unroll
    for (int kk = 0; kk < S2; kk += 4) {
      if (kk < 8 * (t % TPS)) continue;              // X11 lower triangular (compile-time after unrolling)
      const int k = b + kk + tq;
      const double bv = (k >= c + g) ? XW(k, c + g) : 0.0;
      dmma_leaf(c0[t], c1[t], LW(r + g, k), bv);
    }
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int r = b + S2 + 8 * (t / TPS) + g, c = b + 8 * (t % TPS) + 2 * tq;
    XW(r, c) = c0[t];
    XW(r, c + 1) = c1[t];
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int r = b + S2 + 8 * (t / TPS) + g, c = b + 8 * (t % TPS) + g;
    c0[t] = c1[t] = 0.0;
#pragma unroll
    for (int kk = 0; kk < S2; kk += 4) {
      if (kk >= 8 * (t / TPS) + 8) continue;         // X22 lower triangular
      const int k = b + S2 + kk + tq;
      const double av = (k <= r) ? XW(r, k) : 0.0;
      dmma_leaf(c0[t], c1[t], av, XW(k, c));
    }
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int r = b + S2 + 8 * (t / TPS) + g, c = b + 8 * (t % TPS) + 2 * tq;
    XW(r, c) = -c0[t];
    XW(r, c + 1) = -c1[t];
  }
  __syncwarp();
}

__device__ long long g_leaf_clk[8];
#define LEAF_CLK(slot)                                            \
  do {                                                            \
    if (CVXB_LEAF_TIMING && threadIdx.x == 0) {                   \
      long long _c = clock64();                                   \
      g_leaf_clk[slot] += _c - _t0;                               \
      _t0 = _c;                                                   \
    }                                                             \
  } while (0)
#ifndef CVXB_LEAF_TIMING
#define CVXB_LEAF_TIMING 0
#endif

// blockIdx.x selects the diagonal block when only inverting (FACTOR == false).
// 128 x 128 leaf as 8 x 8 sub-blocks of 16.  Per sub-block column: in-warp diagonal factorisation, panel
// solve by per-row forward substitution, rank-16 trailing update -- three block barriers.  The
// triangular inverse is assembled afterwards: the 8 diagonal inverses concurrently (one warp each), then
// three doubling levels X21 = -X22 (L21 X11).
// Tail of the pipelined leaf: the products of one doubling level for the pair of s2-blocks at `base`, their 8 x 8
// output tiles dealt to the warps wr = 0..wn-1 of a group (tile tl = wr + u*wn).
//   level_T:  T = L21 * X11, written straight into the X21 slot (it reads only L21 and X11: no hazard)
//   level_X:  X21 = -X22 * T into registers; the caller stores after a barrier (T lives where X21 goes)
template <int MAXT>
__device__ __forceinline__ void level_T(double* W, int s2, int base, int wr, int wn, int g, int tq) {
  const int tps = s2 >> 3, ntile = tps * tps;
  double c0[MAXT], c1[MAXT];
  int rT[MAXT], cT[MAXT];
  bool on[MAXT];
#pragma unroll
  for (int u = 0; u < MAXT; ++u) {
    c0[u] = c1[u] = 0.0;
    int tl = wr + u * wn;
    on[u] = tl < ntile;
    if (!on[u]) tl = 0;
    rT[u] = base + s2 + 8 * (tl / tps);
    cT[u] = base + 8 * (tl % tps);
  }
#pragma unroll 2
  for (int kk = 0; kk < s2; kk += 4) {
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      if (on[u] && kk >= cT[u] - base) {              // X11 lower triangular: k-blocks above the tile's columns are zero
        const int k = base + kk + tq, c = cT[u] + g;
        const double bv = (k >= c) ? XW(k, c) : 0.0;
        dmma_leaf(c0[u], c1[u], LW(rT[u] + g, k), bv);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < MAXT; ++u) {
    const int r = rT[u] + g, c = cT[u] + 2 * tq;
    if (on[u]) { XW(r, c) = c0[u]; XW(r, c + 1) = c1[u]; }
  }
}

template <int MAXT>
struct LevelX {
  double c0[MAXT], c1[MAXT];
  int rT[MAXT], cT[MAXT];
  bool on[MAXT];
  __device__ __forceinline__ void compute(double* W, int s2, int base, int wr, int wn, int g, int tq) {
    const int tps = s2 >> 3, ntile = tps * tps;
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      c0[u] = c1[u] = 0.0;
      int tl = wr + u * wn;
      on[u] = wr >= 0 && tl < ntile;
      if (!on[u]) tl = 0;
      rT[u] = base + s2 + 8 * (tl / tps);
      cT[u] = base + 8 * (tl % tps);
    }
#pragma unroll 2
    for (int kk = 0; kk < s2; kk += 4) {
#pragma unroll
      for (int u = 0; u < MAXT; ++u) {
        if (on[u] && kk < rT[u] - base - s2 + 8) {      // X22 lower triangular: k-blocks beyond the tile's rows are zero
          const int k = base + s2 + kk + tq, r = rT[u] + g, c = cT[u] + g;
          const double av = (k <= r) ? XW(r, k) : 0.0;
          dmma_leaf(c0[u], c1[u], av, XW(k, c));
        }
      }
    }
  }
  __device__ __forceinline__ void store(double* W, int g, int tq) {
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      const int r = rT[u] + g, c = cT[u] + 2 * tq;
      if (on[u]) { XW(r, c) = -c0[u]; XW(r, c + 1) = -c1[u]; }
    }
  }
};

template <bool FACTOR>
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_kernel(int n_total, int nb_first, double* A, int lda, double* invD, int* flag, double* scal, int flag_slot,
            int mindiag_slot, int col0) {
  extern __shared__ __align__(16) double W[];
  __shared__ double rdiag[NB];
  __shared__ double s_mind;
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;        // 32 x LEAF_WARPS
  int nb, off;
  if (FACTOR) { nb = nb_first; off = 0; }
  else {
    off = blockIdx.x * NB;
    nb = n_total - off;
    if (nb > NB) nb = NB;
  }
  double* Ab = A + (size_t)off * lda + off;
  double* Xg = invD + (size_t)(FACTOR ? 0 : blockIdx.x) * NB * NB;
  asm volatile("griddepcontrol.wait;" ::: "memory");      // programmatic dependent launch: no-op for a plain launch
  long long _t0 = CVXB_LEAF_TIMING ? clock64() : 0;
  if (FACTOR && nb == NB && !(lda & 1) && !((uintptr_t)Ab & 15)) {
    // full block: 16-byte cp.async straight into shared memory, all 16 chunks of a thread in flight at once (the
    // chunk straddling the diagonal drags one element of the X region along; that region is zero-filled below)
#pragma unroll
    for (int c = tid; c < NB * (NB / 2); c += LEAF_THREADS) {
      const int j = c >> 6, q = c & 63;
      if (2 * q + 1 >= j) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&LW(2 * q, j));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Ab + (size_t)j * lda + 2 * q));
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
  } else {
    for (int j = ty; j < nb; j += LEAF_WARPS)
      for (int i = tx; i < nb; i += 32)
        if (i >= j) LW(i, j) = Ab[(size_t)j * lda + i];
  }
  if (tid == 0) { s_fail = 0; s_mind = 1e300; }
  __syncthreads();
  LEAF_CLK(0);

  const int nblk = (nb + SUB - 1) / SUB;
  const int g = tx >> 2, tq = tx & 3;
  // ---- zero-fill the X region first (the doubling levels read not-yet-written entries as zeros); when factoring,
  //      warps 1..15 do it while warp 0 factors the first diagonal block
  if (!FACTOR || ty > 0)
    for (int i = FACTOR ? ty - 1 : ty; i < nb; i += FACTOR ? LEAF_WARPS - 1 : LEAF_WARPS)   // lanes along j: unit stride
      for (int j = tx; j <= i; j += 32) XW(i, j) = 0.0;
  // In a full leaf the inverse is assembled INSIDE the factorisation windows (see below); first pair of each level
  // that is still to do when the factorisation ends:
  const bool pipelined = FACTOR && nb == NB;
  int first16 = 0, first32 = 0;
  bool diag_inv_done = false, tail_done = false;
  if (FACTOR) {
    // Software pipeline over the 8 sub-block columns.  The serial pivot chain (warp 0) is the critical path, so
    // everything else is moved beside it:   window kb = { warp 0: update + factor diagonal block kb+1 }  ||
    // { warps 1..15: the rest of trailing update kb; warp 15: inverse of diagonal block kb; warps 14 / 13: the
    // doubling levels of the inverse whose inputs are final }, then the panel solve kb+1.  Two block barriers per
    // sub-block column instead of three, and only the last levels of the inverse remain after the loop.
    if (ty == 0) warp_diag_factor(W, rdiag, 0, nb < SUB ? nb : SUB, col0, &s_fail, &s_mind);
    __syncthreads();
    LEAF_CLK(1);
    for (int kb = 0; kb < nblk; ++kb) {
      const int o = kb * SUB;
      const int r0 = o + SUB;
      const int nrows = nb - r0;
      if (nrows <= 0) break;
      // panel: row r of A(r0.., o..o+15) := row * L_d^-T by forward substitution, one thread per row
      if (tid < nrows) {
        const int r = r0 + tid;
        // right-looking substitution: once x[c] is known it is folded into all later columns at once, so the
        // dependent chain is 16 x (multiply + one FMA) instead of the 136 in-order FMAs of the dot-product form
        double x[SUB];
#pragma unroll
        for (int c = 0; c < SUB; ++c) x[c] = LW(r, o + c);
#pragma unroll
        for (int c = 0; c < SUB; ++c) {
          x[c] *= rdiag[o + c];
#pragma unroll
          for (int c2 = c + 1; c2 < SUB; ++c2) x[c2] = fma(-x[c], LW(o + c2, o + c), x[c2]);
        }
#pragma unroll
        for (int c = 0; c < SUB; ++c) LW(r, o + c) = x[c];
      }
      __syncthreads();
      LEAF_CLK(2);
      // window kb.  Trailing update: A(r,c) -= sum_k L(r,o+k) L(c,o+k), r >= c >= r0, as 8 x 8 DMMA tiles
      // (A(m,k) = L(row, o+k), B(k,n) = L(col, o+k); out-of-range rows only feed discarded outputs).  Tiles 0..2
      // are the next diagonal block: warp 0 takes them and goes straight on to factor it.
      {
        const int T = (nrows + 7) >> 3;
        const int ntile = T * (T + 1) / 2;
        const int nfirst = ntile < 3 ? ntile : 3;
        auto tile = [&](int tl) {
          int ti = (int)((sqrtf(8.0f * (float)tl + 1.0f) - 1.0f) * 0.5f);
          while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
          while (ti * (ti + 1) / 2 > tl) --ti;
          const int tj = tl - ti * (ti + 1) / 2;
          const int rr = r0 + 8 * ti + g, cr = r0 + 8 * tj + g;
          double c0 = 0.0, c1 = 0.0;
#pragma unroll
          for (int kk = 0; kk < SUB; kk += 4) dmma_leaf(c0, c1, LW(rr, o + kk + tq), LW(cr, o + kk + tq));
          const int cc = r0 + 8 * tj + 2 * tq;
          if (rr < nb) {
            if (cc <= rr) LW(rr, cc) -= c0;
            if (cc + 1 <= rr) LW(rr, cc + 1) -= c1;
          }
        };
        // tiles 0..2 = the next diagonal block: warps 1..3 take one each and release warp 0 through a named barrier
        if (ty <= 3) {
          if (ty >= 1 && ty - 1 < nfirst) tile(ty - 1);
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        // warps 4, 8, 12 share warp 0's scheduler and FP64 pipe: keep them idle so that the pivot chain's dependent
        // FP64 operations never queue behind DMMAs
        if ((ty & 3) != 0) {
          const int wk = ty - 1 - (ty >> 2);         // 0..11
          for (int tl = nfirst + wk; tl < ntile; tl += LEAF_WARPS - 4) tile(tl);
        }
        if (ty == 0) {
          warp_diag_factor(W, rdiag, r0, nrows < SUB ? nrows : SUB, col0, &s_fail, &s_mind);
        } else if (pipelined) {
          if (ty == LEAF_WARPS - 1) warp_diag_inverse(W, rdiag, o, SUB);                      // X_kb,kb
          if (ty == LEAF_WARPS - 2 && kb >= 2 && (kb & 1) == 0) warp_level_job<16>(W, (kb - 2) * SUB);   // blocks kb-2, kb-1
          if (ty == LEAF_WARPS - 3 && kb == 5) warp_level_job<32>(W, 0);                       // blocks 0..3
        }
      }
      __syncthreads();
      LEAF_CLK(3);
    }
    if (pipelined) { first16 = 3; first32 = 1; }
    if (tid == 0) {
      if (s_fail && flag[flag_slot] == 0) flag[flag_slot] = s_fail;
      if (s_mind < scal[mindiag_slot]) scal[mindiag_slot] = s_mind;
    }
    if (pipelined) {
      // Tail of the inverse.  What is left: the last diagonal block's inverse, pair 3 of level 16 (blocks 6,7), pair 1
      // of level 32 (blocks 4..7) and level 64.  Every T = L21 X11 of these depends only on data that is final by now
      // (X(6,6); X(4..5,4..5); X(0..3,0..3)), so all three are computed in ONE phase beside the last diagonal inverse;
      // afterwards only the three X21 = -X22 T products remain in sequence: 4 dependent phases instead of 7.
      if (ty == 0) warp_diag_inverse(W, rdiag, (nblk - 1) * SUB, SUB);
      else if (ty == 1) level_T<4>(W, 16, 96, 0, 1, g, tq);                 // 4 tiles
      else if (ty <= 5) level_T<4>(W, 32, 64, ty - 2, 4, g, tq);            // 16 tiles on 4 warps
      else level_T<7>(W, 64, 0, ty - 6, LEAF_WARPS - 6, g, tq);             // 64 tiles on 10 warps
      __syncthreads();
      {
        LevelX<1> x16;
        x16.compute(W, 16, 96, ty < 4 ? ty : -1, 4, g, tq);
        __syncthreads();
        x16.store(W, g, tq);
        __syncthreads();
      }
      {
        LevelX<1> x32;
        x32.compute(W, 32, 64, ty, LEAF_WARPS, g, tq);
        __syncthreads();
        x32.store(W, g, tq);
        __syncthreads();
      }
      {
        LevelX<4> x64;
        x64.compute(W, 64, 0, ty, LEAF_WARPS, g, tq);
        __syncthreads();
        x64.store(W, g, tq);
      }
      diag_inv_done = true;
      tail_done = true;
    }
  } else {
    if (tid < nb) {
      double pv = LW(tid, tid);
      if (pv == 0.0) { flag[F_ZERO_DIAG] = 1; pv = 1.0; }
      rdiag[tid] = 1.0 / pv;
    }
  }
  __syncthreads();

  // ---- (rest of the) triangular inverse: the diagonal 16-blocks concurrently, then the doubling levels
  //      X21 = -X22 (L21 X11) for s = 16, 32, 64
  if (!diag_inv_done) {
    if (ty < nblk) {
      const int o = ty * SUB;
      warp_diag_inverse(W, rdiag, o, (nb - o) < SUB ? (nb - o) : SUB);
    }
    __syncthreads();
  }
  LEAF_CLK(4);
  for (int sh = 4; !tail_done && (1 << sh) < nb; ++sh) {           // s = 16, 32, 64
    const int s2 = 1 << sh;
    const int pair0 = (sh == 4) ? first16 : (sh == 5 ? first32 : 0);     // pairs before this one are already done
    const int npairs = (nb + 2 * s2 - 1) / (2 * s2) - pair0;
    // Both products of the level on the DMMA pipe, 8 x 8 output tiles spread over the 16 warps:
    //   T   = L21 * X11      A(m,k) = L(r,k),            B(k,n) = X(k,c) (zero for k < c)
    //   X21 = -X22 * T       A(m,k) = X(r,k) (k <= r),   B(k,n) = T(k,c)
    const int tps = s2 >> 3;                         // tiles per side of one block
    const int ntile = npairs * tps * tps;
    constexpr int MAXT = 4;                          // 64 tiles at the widest level / 16 warps
    double c0[MAXT], c1[MAXT];
    int rT[MAXT], cT[MAXT], bT[MAXT];
    bool onT[MAXT];
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      c0[u] = c1[u] = 0.0;
      int tl = ty + u * LEAF_WARPS;
      onT[u] = tl < ntile;
      if (!onT[u]) tl = 0;
      const int pr = tl / (tps * tps), rem = tl - pr * tps * tps;
      bT[u] = (pr + pair0) * 2 * s2;
      rT[u] = bT[u] + s2 + 8 * (rem / tps);
      cT[u] = bT[u] + 8 * (rem % tps);
    }
#pragma unroll 2
    for (int kk = 0; kk < s2; kk += 4) {
#pragma unroll
      for (int u = 0; u < MAXT; ++u) {
        if (onT[u] && kk >= cT[u] - bT[u]) {            // warp-uniform; X11 is lower triangular: k-blocks above the tile's columns are zero
          const int k = bT[u] + kk + tq, c = cT[u] + g;
          const double bv = (k >= c) ? XW(k, c) : 0.0;
          dmma_leaf(c0[u], c1[u], LW(rT[u] + g, k), bv);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      const int r = rT[u] + g, c = cT[u] + 2 * tq;
      if (onT[u] && r < nb) { XW(r, c) = c0[u]; XW(r, c + 1) = c1[u]; }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < MAXT; ++u) c0[u] = c1[u] = 0.0;
#pragma unroll 2
    for (int kk = 0; kk < s2; kk += 4) {
#pragma unroll
      for (int u = 0; u < MAXT; ++u) {
        if (onT[u] && kk < rT[u] - bT[u] - s2 + 8) {      // X22 is lower triangular: k-blocks beyond the tile's rows are zero
          const int k = bT[u] + s2 + kk + tq, r = rT[u] + g, c = cT[u] + g;
          const double av = (k <= r && r < nb) ? XW(r, k) : 0.0;
          const double bv = (k < nb) ? XW(k, c) : 0.0;
          dmma_leaf(c0[u], c1[u], av, bv);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      const int r = rT[u] + g, c = cT[u] + 2 * tq;
      if (onT[u] && r < nb) { XW(r, c) = -c0[u]; XW(r, c + 1) = -c1[u]; }
    }
    __syncthreads();
  }
  LEAF_CLK(5);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the panel GEMM's CTAs may take their seats
  for (int j = ty; j < NB; j += LEAF_WARPS)
    for (int i = tx; i < NB; i += 32) {
      double x = (i < nb && j < nb && i >= j) ? XW(i, j) : 0.0;
      Xg[(size_t)j * NB + i] = x;
      if (FACTOR && i < nb && j < nb && i >= j) Ab[(size_t)j * lda + i] = LW(i, j);
    }
  LEAF_CLK(6);
}

__global__ void potrf_reset_kernel(int* flag, double* scal, int flag_slot, int mindiag_slot) {
  flag[flag_slot] = 0;
  scal[mindiag_slot] = 1e300;
}

inline int split_point(int n) {
  int a = (n / 2) / NB * NB;
  if (a < NB) a = NB;
  return a;
}

// X L11' = A21  (rows M): A21 := A21 L11^-T, L11 n1 x n1 lower with inverse diagonal blocks
int trsm_right_lt(Handle& h, int M, int n1, const double* L, int ldl, const double* invD, double* A21, int lda) {
  if (M <= 0 || n1 <= 0) return CVXB_OK;
  if (n1 <= NB) {
    // X = A21 * invL11' through a scratch panel (the product cannot alias its operand once the grid has
    // more than one tile column), so that small tiles can spread M x 128 outputs over all SMs
    const int lds = pad_ld(M);
    if ((size_t)lds * n1 + 65536 <= h.part_cap) {
      double* Xs = h.d_part;
      GemmArgs g{M, n1, n1, A21, lda, false, invD, NB, false, Xs, lds, 1.0, 0.0, 0};
      CVXB_TRY(gemm_dmma(h, g));
      return copy_matrix(h, M, n1, Xs, lds, A21, lda);
    }
    // in place: the tile grid has a single column (N = n1 <= 128), so a CTA reads only the rows it writes
    GemmArgs g{M, n1, n1, A21, lda, false, invD, NB, false, A21, lda, 1.0, 0.0, 0, 128};
    return gemm_dmma(h, g);
  }
  int a = split_point(n1), b = n1 - a;
  CVXB_TRY(trsm_right_lt(h, M, a, L, ldl, invD, A21, lda));
  // A2 -= X1 * L_ba'   (k over a, n over b): B(k,n) = L_ba(n,k) = L[(k)*ldl + a + n]  -> N contiguous
  GemmArgs g{M, b, a, A21, lda, false, L + a, ldl, false, A21 + (size_t)a * lda, lda, -1.0, 1.0, 0};
  CVXB_TRY(gemm_dmma(h, g));
  return trsm_right_lt(h, M, b, L + (size_t)a * ldl + a, ldl, invD + (size_t)(a / NB) * NB * NB, A21 + (size_t)a * lda,
                       lda);
}

int potrf_rec(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, int col0, bool plain,
              double* B, int ldb, int r);

// While alive, everything enqueued through the handle goes to one of the bulk lanes: lane 0 = stream3, where GEMMs of at
// least half a wave of tiles run as persistent grids that leave `dag_reserve` SMs free; lanes 1..3 = side streams with
// plain launches.  Every lane has its own scratch for the out-of-place leaf solves.
struct BulkScope {
  Handle& h;
  cudaStream_t keep_stream;
  double* keep_part;
  size_t keep_cap;
  int keep_reserve;
  BulkScope(Handle& h_, int lane) : h(h_), keep_stream(h_.stream), keep_part(h_.d_part), keep_cap(h_.part_cap),
                                    keep_reserve(h_.sk_reserve) {
    h.stream = lane == 0 ? h.stream3 : h.side[lane - 1];
    h.d_part = h.d_part3 + (size_t)lane * PART3_DOUBLES;
    h.part_cap = PART3_DOUBLES;
    h.sk_reserve = lane == 0 ? h.dag_reserve : -1;      // -1: inside the DAG schedule (no nesting), plain launches
  }
  ~BulkScope() {
    h.stream = keep_stream;
    h.d_part = keep_part;
    h.part_cap = keep_cap;
    h.sk_reserve = keep_reserve;
  }
};

// Tile-DAG schedule for matrices beyond the L2-resident regime (n >= dag_min_n; C4: n = 8192 with the 2049 columns of
// [DA', Dq] riding along, C5: n = 16385).  Right-looking over diagonal blocks of `dag_block` columns with a look-ahead of
// one block, on two lanes:
//   chain (the handle's stream, + stream2 inside):  P(k) = look-ahead Cholesky of the diagonal block A_kk -- a serial chain
//       of 128-column leaves (~60 us each) that keeps only a handful of SMs busy;
//   bulk (stream3, persistent GEMM grids on SMs - dag_reserve CTAs), per level k once P(k) is done:
//       A[k+1.., k] := A[k+1.., k] L_kk^-T;  A_{k+1,k+1} -= L_{k+1,k} L_{k+1,k}'  -> P(k+1) may start;
//       rest of the trailing update (one lower-triangular SYRK over all remaining blocks, minus that first block);
//       Y_k := L_kk^-1 B_k;  B[k+1..] -= L[k+1.., k] Y_k.
// P(k+1) runs in the shadow of level k's bulk work (4.4 / 2.3 / 0.8 ms of GEMMs against a chain of ~1 ms at C4), so of the
// four block chains only the first and the tail of the last are exposed; the recursive schedule ran all of them, and the
// forward substitution's rank-128 pieces, with the machine mostly idle.  Every dependency is an event between the two
// lanes, so the schedule is capturable into the per-step CUDA graph like the look-ahead schedule.
// Diagonal blocks of the tile-DAG schedule: `nbk` columns in the middle, where a block's chain hides behind the bulk work
// of the level before; half a block first (nothing can overlap the first chain) and geometrically shrinking blocks at the
// end (the last chain and the last triangular solve of the right-hand sides are exposed too, and the bulk work left to
// hide a chain behind shrinks with the trailing matrix).  All starts are multiples of 128.
static void dag_blocks(int n, int nbk, std::vector<int>& start) {
  start.clear();
  int pos = 0;
  static const bool flat = getenv("CVXB_DAG_FLAT") != nullptr;
  while (pos < n) {
    const int rem = n - pos;
    int b;
    if (flat) b = nbk;
    else if (pos == 0 && n >= 3 * nbk) {
      static const int first_div = getenv("CVXB_DAG_FIRST_DIV") ? atoi(getenv("CVXB_DAG_FIRST_DIV")) : 2;
      b = (nbk / (first_div > 0 ? first_div : 2)) / NB * NB;
      if (b < NB) b = NB;
    }
    else if (rem >= 2 * nbk) b = nbk;
    else if (rem <= 5 * NB) b = rem;
    else b = ((rem + 1) / 2 + NB - 1) / NB * NB;
    if (b > rem) b = rem;
    start.push_back(pos);
    pos += b;
  }
  start.push_back(n);
}

int potrf_dag(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, int col0, double* B,
              int ldb, int r) {
  // Without a wide right-hand-side block there is less bulk work to hide a chain behind: narrower blocks (shorter chains,
  // shorter exposed first and last block) win below n = 12288 (n = 8192: 8.7 ms with 1024 columns against 9.2 ms with
  // 2048), and from there on the chain lane has so much slack that 4 reserved SMs are enough (n = 16385: 52.4 against
  // 53.0 ms).  cvxb_debug_set_schedule switches these defaults off.
  int nbk = h.dag_block;
  const int keep_reserve = h.dag_reserve;
  if (h.dag_auto && r <= 1024) {
    if (n < 12288) nbk = nbk / 2 >= 4 * NB ? (nbk / 2) / NB * NB : nbk;
    else h.dag_reserve = h.dag_reserve > 4 ? 4 : h.dag_reserve;
  }
  struct RestoreReserve {
    Handle& h; int v;
    ~RestoreReserve() { h.dag_reserve = v; }
  } restore_reserve{h, keep_reserve};
  std::vector<int> bstart;
  dag_blocks(n, nbk, bstart);
  const int T = (int)bstart.size() - 1;
  if (2 * T + 10 > (int)h.dag_events.size()) {
    set_last_error("potrf_dag: %d diagonal blocks exceed the event pool", T);
    return CVXB_EINVAL;
  }
  cudaStream_t sa = h.stream, sc = h.stream3;
  // CVXB_DAG_TRACE=1: time stamps (CUDA events) at every piece boundary of both lanes, printed after the factorisation
  static const bool trace_env = getenv("CVXB_DAG_TRACE") != nullptr;
  const bool trace = trace_env && !h.capturing;
  struct Mark { const char* what; int k; int lane; cudaEvent_t ev; };
  std::vector<Mark> marks;
  auto mark = [&](const char* what, int k, int lane) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, lane ? sc : sa);
    marks.push_back(Mark{what, k, lane, e});
  };
  mark("start", 0, 0);
  for (int k = 0; k < T; ++k) {
    const int k0 = bstart[k], kb = bstart[k + 1] - k0;
    const int rem = n - k0 - kb;
    double* Akk = A + (size_t)k0 * lda + k0;
    double* invDk = invD + (size_t)(k0 / NB) * NB * NB;
    cudaEvent_t evP = h.dag_events[2 * k], evR = h.dag_events[2 * k + 1];
    if (k > 0) CVXB_CUDA_OK(cudaStreamWaitEvent(sa, h.dag_events[2 * (k - 1) + 1], 0));       // A_kk fully updated
    mark("P begin", k, 0);
    CVXB_TRY(potrf_rec(h, kb, Akk, lda, invDk, flag_slot, mindiag_slot, col0 + k0, false, nullptr, 0, 0));
    CVXB_CUDA_OK(cudaEventRecord(evP, sa));
    mark("P end", k, 0);
    double* A21 = Akk + kb;                                     // rows below the diagonal block, this block column
    const int kn = k + 1 < T ? bstart[k + 2] - bstart[k + 1] : 0;      // next diagonal block
    // The triangular solves are recursions of ~47 dependent launches per 2048 columns, most of them small (K = 128..512)
    // and bound by launch + pipeline-fill latency (~12 us each), not by their flops.  Independent pieces -- row groups of
    // A21 L_kk^-T, column groups of L_kk^-1 B_k -- therefore run beside each other on the side lanes: their latencies
    // overlap instead of adding up.  Lane 0 keeps the rows of the next diagonal block (the critical path to P(k+1)).
    int nside = 0;                                              // side lanes used at this level
    cudaEvent_t evS[3];
    // a side lane needs P(k) and everything level k-1 wrote below / beside it (its trailing update and right-hand-side
    // update end lane 0's work of that level)
    auto side_begin = [&](int lane) {
      cudaError_t e = cudaStreamWaitEvent(h.side[lane - 1], evP, 0);
      if (e == cudaSuccess && k > 0) e = cudaStreamWaitEvent(h.side[lane - 1], h.dag_events[2 * T + 7 + ((k - 1) & 1)], 0);
      return e;
    };
    auto side_end = [&](int lane) {
      evS[nside] = h.dag_events[2 * T + 1 + 3 * (k & 1) + nside];
      cudaError_t e = cudaEventRecord(evS[nside], h.side[lane - 1]);
      ++nside;
      return e;
    };
    double* Anext = A + (size_t)(k0 + kb) * lda + (k0 + kb);
    // lane 0 first (host enqueue order = the order the critical path needs): rows of the next diagonal block, its update
    CVXB_CUDA_OK(cudaStreamWaitEvent(sc, evP, 0));
    if (rem > 0) {
      BulkScope bulk(h, 0);
      mark("bulk level begin", k, 1);
      CVXB_TRY(trsm_right_lt(h, kn, kb, Akk, lda, invDk, A21, lda));
      mark("trsm right (next block rows) done", k, 1);
      GemmArgs gn{kn, kn, kb, A21, lda, false, A21, lda, false, Anext, lda, -1.0, 1.0, 1};
      CVXB_TRY(gemm_dmma(h, gn));
      CVXB_CUDA_OK(cudaEventRecord(evR, sc));
      mark("next block updated", k, 1);
    }
    const int rows_rest = rem - kn;                             // rows of A21 below the next diagonal block
    int lane_next = 1;
    if (rows_rest > 0) {
      // up to two side lanes for the remaining rows of the TRSM (three without a right-hand-side block)
      int groups = rows_rest >= 4096 ? 2 : 1;
      if (!B && rows_rest >= 6144) groups = 3;
      const int per = ((rows_rest + groups - 1) / groups + NB - 1) / NB * NB;
      for (int gI = 0; gI < groups; ++gI) {
        const int r0 = kn + gI * per, rr = rem - r0 < per ? rem - r0 : per;
        if (rr <= 0) break;
        CVXB_CUDA_OK(side_begin(lane_next));
        {
          BulkScope lane(h, lane_next);
          CVXB_TRY(trsm_right_lt(h, rr, kb, Akk, lda, invDk, A21 + r0, lda));
        }
        CVXB_CUDA_OK(side_end(lane_next));
        ++lane_next;
      }
    }
    cudaEvent_t evY = nullptr;
    if (B && lane_next <= 3) {
      // Y_k = L_kk^-1 B_k on a side lane
      CVXB_CUDA_OK(side_begin(lane_next));
      {
        BulkScope lane(h, lane_next);
        CVXB_TRY(trsm_rec(h, kb, r, Akk, lda, invDk, B + k0, ldb, false));
      }
      evY = h.dag_events[2 * T + 1 + 3 * (k & 1) + nside];
      CVXB_CUDA_OK(cudaEventRecord(evY, h.side[lane_next - 1]));
      ++lane_next;
    }
    {
      BulkScope bulk(h, 0);
      for (int i = 0; i < nside; ++i) CVXB_CUDA_OK(cudaStreamWaitEvent(sc, evS[i], 0));
      if (rem > kn) {                                           // kn is a whole number of 128-tiles here
        mark("trsm right (all rows) done", k, 1);
        GemmArgs gt{rem, rem, kb, A21, lda, false, A21, lda, false, Anext, lda, -1.0, 1.0, 1};
        gt.tri_skip = kn / NB;
        // timed on the bulk lane (events on stream3): the update shares the machine with the chain of P(k+1)
        const bool timed = kb >= 1024 && rem >= 2048;
        if (timed) CVXB_TRY(prof_begin(h, PROF_CHOL_TRAIL));
        CVXB_TRY(gemm_dmma(h, gt));
        if (timed) CVXB_TRY(prof_end(h, PROF_CHOL_TRAIL, (double)kb * ((double)rem * (rem + 1.0) - (double)kn * (kn + 1.0))));
        mark("trailing syrk done", k, 1);
      }
      if (B) {
        if (evY) CVXB_CUDA_OK(cudaStreamWaitEvent(sc, evY, 0));
        else CVXB_TRY(trsm_rec(h, kb, r, Akk, lda, invDk, B + k0, ldb, false));                // Y_k
        mark("Y_k done", k, 1);
        if (rem > 0) {
          GemmArgs gb{rem, r, kb, A21, lda, false, B + k0, ldb, true, B + k0 + kb, ldb, -1.0, 1.0, 0};
          CVXB_TRY(gemm_dmma(h, gb));
          mark("rhs update done", k, 1);
        }
      }
    }
    CVXB_CUDA_OK(cudaEventRecord(h.dag_events[2 * T + 7 + (k & 1)], sc));      // level k complete on lane 0
  }
  cudaEvent_t evJ = h.dag_events[2 * T];
  CVXB_CUDA_OK(cudaEventRecord(evJ, sc));
  CVXB_CUDA_OK(cudaStreamWaitEvent(sa, evJ, 0));
  mark("joined", T, 0);
  if (trace) {
    cudaStreamSynchronize(sa);
    fprintf(stderr, "potrf_dag n=%d r=%d block=%d reserve=%d\n", n, r, nbk, h.dag_reserve);
    for (const Mark& m : marks) {
      float ms = 0;
      cudaEventElapsedTime(&ms, marks[0].ev, m.ev);
      fprintf(stderr, "  %8.3f ms  %s  k=%d  %s\n", ms, m.lane ? "bulk " : "chain", m.k, m.what);
    }
    for (const Mark& m : marks) cudaEventDestroy(m.ev);
  }
  return CVXB_OK;
}

// Recursive halving; sub-problems that fit the L2-resident regime switch to the look-ahead schedule.
// With a right-hand-side block B (n x r) the forward substitution B := L^-1 B rides along: every piece of it is
// issued as soon as the columns of L it needs exist (in the look-ahead regime on the second stream, in the shadow
// of the leaf chain), instead of as a separate sweep over L after the factorisation.
int potrf_rec(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, int col0, bool plain = false,
              double* B = nullptr, int ldb = 0, int r = 0) {
  if (!plain && h.stream3 && h.stream2 && h.dag_block >= NB && n >= h.dag_min_n && n > h.dag_block &&
      n / h.dag_block < 48 && h.sk_reserve == 0 && !h.in_dag) {
    h.in_dag = true;
    const int st = potrf_dag(h, n, A, lda, invD, flag_slot, mindiag_slot, col0, B, ldb, r);
    h.in_dag = false;
    return st;
  }
  // a wide right-hand-side block adds (rem x r x 128) of rank-128 GEMM work per step to the second stream: beyond
  // ~2560 columns that no longer fits in the shadow of the leaf chain (C4, r = 2049: 57.1 ms per step with the
  // look-ahead schedule up to 4608 against 54.3 ms with the switch at 2560)
  static const int la_wide = getenv("CVXB_LA_WIDE") ? atoi(getenv("CVXB_LA_WIDE")) : 2560;
  const int la_max = (r > 1024 && rl_max_n() > la_wide) ? la_wide : rl_max_n();
  if (!plain && n > NB && n <= la_max && h.stream2) {
    CVXB_TRY(prof_begin(h, PROF_LOOKAHEAD));
    CVXB_TRY(potrf_lookahead(h, n, A, lda, invD, flag_slot, mindiag_slot, col0, B, ldb, r));
    return prof_end(h, PROF_LOOKAHEAD, (double)n * n * n / 3.0 + (double)n * n * r);
  }
  if (n <= NB) {
    CVXB_LAUNCH(h, leaf_kernel<true>, 1, LEAF_THREADS, LEAF_SMEM, n, n, A, lda, invD, h.d_flag, h.d_scal, flag_slot,
                mindiag_slot, col0);
    if (B) return trsm_rec(h, n, r, A, lda, invD, B, ldb, false);
    return CVXB_OK;
  }
  int a = split_point(n), b = n - a;
  CVXB_TRY(potrf_rec(h, a, A, lda, invD, flag_slot, mindiag_slot, col0, plain, B, ldb, r));
  double* A21 = A + a;
  double* A22 = A + (size_t)a * lda + a;
  CVXB_TRY(prof_begin(h, PROF_TRSM_RIGHT));
  CVXB_TRY(trsm_right_lt(h, b, a, A, lda, invD, A21, lda));
  CVXB_TRY(prof_end(h, PROF_TRSM_RIGHT, (double)b * a * a));
  // A22 -= A21 A21'  lower: A(m,k) = A21[k*lda + m] (M contiguous), B(k,n) = A21(n,k) (N contiguous)
  GemmArgs g{b, b, a, A21, lda, false, A21, lda, false, A22, lda, -1.0, 1.0, 1};
  g.streamk = true;      // main stream only (the look-ahead schedule has joined)
  CVXB_TRY(prof_begin(h, PROF_CHOL_TRAIL));
  CVXB_TRY(gemm_dmma(h, g));
  CVXB_TRY(prof_end(h, PROF_CHOL_TRAIL, (double)a * b * ((double)b + 1.0)));      // lower triangle, mul + add
  if (B) {   // B2 -= L21 Y1
    GemmArgs gu{b, r, a, A21, lda, false, B, ldb, true, B + a, ldb, -1.0, 1.0, 0};
    CVXB_TRY(gemm_dmma(h, gu));
  }
  return potrf_rec(h, b, A22, lda, invD + (size_t)(a / NB) * NB * NB, flag_slot, mindiag_slot, col0 + a, plain,
                   B ? B + a : nullptr, ldb, r);
}
int potrf_rec_plain(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, int col0,
                    double* B = nullptr, int ldb = 0, int r = 0) {
  return potrf_rec(h, n, A, lda, invD, flag_slot, mindiag_slot, col0, true, B, ldb, r);
}

int trsm_rec(Handle& h, int n, int r, const double* L, int ldl, const double* invD, double* B, int ldb, bool trans) {
  if (n <= NB) {
    const int lds = pad_ld(n);
    if ((size_t)lds * r + 65536 <= h.part_cap) {      // out of place through the scratch panel (see above)
      double* Ys = h.d_part;
      GemmArgs g{n, r, n, invD, NB, trans, B, ldb, true, Ys, lds, 1.0, 0.0, 0};
      CVXB_TRY(gemm_dmma(h, g));
      return copy_matrix(h, n, r, Ys, lds, B, ldb);
    }
    // in place: single tile row (M = n <= 128): a CTA reads only the columns it writes
    GemmArgs g{n, r, n, invD, NB, trans, B, ldb, true, B, ldb, 1.0, 0.0, 0, 128};
    return gemm_dmma(h, g);
  }
  int a = split_point(n), b = n - a;
  const double* L22 = L + (size_t)a * ldl + a;
  const double* invD2 = invD + (size_t)(a / NB) * NB * NB;
  if (!trans) {
    CVXB_TRY(trsm_rec(h, a, r, L, ldl, invD, B, ldb, false));
    // B2 -= L21 Y1 : A(m,k) = L21[k*ldl + m] (M contiguous); B(k,n) = B[n*ldb + k] (K contiguous)
    GemmArgs g{b, r, a, L + a, ldl, false, B, ldb, true, B + a, ldb, -1.0, 1.0, 0};
    CVXB_TRY(gemm_dmma(h, g));
    return trsm_rec(h, b, r, L22, ldl, invD2, B + a, ldb, false);
  }
  CVXB_TRY(trsm_rec(h, b, r, L22, ldl, invD2, B + a, ldb, true));
  // B1 -= L21' X2 : A(m,k) = L21(k,m) = L[m*ldl + a + k] (K contiguous); B(k,n) = B[n*ldb + a + k]
  GemmArgs g{a, r, b, L + a, ldl, true, B + a, ldb, true, B, ldb, -1.0, 1.0, 0};
  CVXB_TRY(gemm_dmma(h, g));
  return trsm_rec(h, a, r, L, ldl, invD, B, ldb, true);
}

// ------------------------------------------------------------------------------- single-RHS solves
// One launch per 128-column block: every CTA recomputes y_k = invD_k b_k (16k FMA, the
// block inverse comes from L2), CTA 0 publishes it, and each CTA folds L[rows, block] y_k into its own
// 256 rows of b.  HBM-bound: L is read exactly once, coalesced down the columns.
constexpr int TRSV_ROWS = 64;

__global__ void __launch_bounds__(256) trsv_fwd_step_kernel(int n, int k0, int kb, const double* __restrict__ L, int ldl,
                                                            const double* __restrict__ invDk, double* __restrict__ b,
                                                            double* __restrict__ out) {
  __shared__ double ys[NB];
  __shared__ double bs[NB];
  __shared__ double red[4][NB];
  const int tid = threadIdx.x;
  if (tid < NB) bs[tid] = tid < kb ? b[k0 + tid] : 0.0;
  __syncthreads();
  {   // y = invD_k b_k : row i = tid & 127, two column halves (invD is lower: zeros above the diagonal)
    const int i = tid & (NB - 1), half = tid >> 7;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double* col = invDk + i + (size_t)(half * 64) * NB;
    const double* bb = bs + half * 64;
#pragma unroll 4
    for (int j = 0; j < 64; j += 4) {
      a0 = fma(col[(size_t)j * NB], bb[j], a0);
      a1 = fma(col[(size_t)(j + 1) * NB], bb[j + 1], a1);
      a2 = fma(col[(size_t)(j + 2) * NB], bb[j + 2], a2);
      a3 = fma(col[(size_t)(j + 3) * NB], bb[j + 3], a3);
    }
    red[half][i] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (tid < NB) {
    const double y = red[0][tid] + red[1][tid];
    ys[tid] = y;
    if (blockIdx.x == 0 && tid < kb) out[k0 + tid] = y;
  }
  __syncthreads();
  // b[r] -= L[r, block] . y : 64 rows per CTA, 4 column groups of 32
  const int rl = tid & 63, grp = tid >> 6;
  const int r = k0 + kb + blockIdx.x * TRSV_ROWS + rl;
  double part = 0.0;
  if (r < n) {
    const double* Lr = L + (size_t)(k0 + grp * 32) * ldl + r;
    const double* yy = ys + grp * 32;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (grp * 32 + j + 3 < kb) {
        a0 = fma(Lr[(size_t)j * ldl], yy[j], a0);
        a1 = fma(Lr[(size_t)(j + 1) * ldl], yy[j + 1], a1);
        a2 = fma(Lr[(size_t)(j + 2) * ldl], yy[j + 2], a2);
        a3 = fma(Lr[(size_t)(j + 3) * ldl], yy[j + 3], a3);
      } else {
        for (int q = 0; q < 4; ++q)
          if (grp * 32 + j + q < kb) a0 = fma(Lr[(size_t)(j + q) * ldl], yy[j + q], a0);
      }
    }
    part = (a0 + a1) + (a2 + a3);
  }
  red[grp][rl] = part;
  __syncthreads();
  if (tid < 64 && r < n) b[r] -= (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
}

// backward (L' x = y): block k from the last to the first; x_k = invD_k' y_k, then
// y[c] -= L[block rows, c] . x_k for every column c < k0 (one warp per column, 64 columns per CTA)
__global__ void __launch_bounds__(256) trsv_bwd_step_kernel(int k0, int kb, const double* __restrict__ L, int ldl,
                                                            const double* __restrict__ invDk, double* __restrict__ y,
                                                            double* __restrict__ out) {
  __shared__ double xs[NB];
  __shared__ double ysb[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < NB) ysb[tid] = tid < kb ? y[k0 + tid] : 0.0;
  __syncthreads();
  for (int i = warp; i < kb; i += 8) {        // x_i = sum_{j >= i} invD(j,i) y_j
    double a = 0.0;
    for (int j = i + lane; j < kb; j += 32) a = fma(invDk[j + (size_t)i * NB], ysb[j], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      xs[i] = a;
      if (blockIdx.x == 0) out[k0 + i] = a;
    }
  }
  __syncthreads();
  const int cbase = blockIdx.x * 64;
  for (int cc = warp; cc < 64; cc += 8) {
    const int c = cbase + cc;
    if (c >= k0) break;
    const double* col = L + (size_t)c * ldl + k0;
    double a = 0.0;
    for (int i = lane; i < kb; i += 32) a = fma(col[i], xs[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) y[c] -= a;
  }
}

__global__ void copy_vec_kernel(int n, const double* __restrict__ a, double* __restrict__ b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = a[i];
}

// ---- single-launch wavefront solves ----------------------------------------------------------------
// One CTA per 128-row block, all co-resident (cooperative launch; n <= 128 * #SMs).  CTA c folds
// L[c,k] y_k into its right-hand side as soon as block k publishes y_k (flag in global memory), then
// solves its own diagonal block with the inverse and publishes y_c.  No launch per block, no grid-wide
// barrier: the critical path is one flag hand-off + two 128 x 128 GEMVs per block.
__device__ __forceinline__ bool wait_flag(const int* flag, int* abort_flag) {
  // thread 0 only; bounded spin so that a scheduling surprise can never hang the GPU
  const volatile int* f = flag;
  const volatile int* ab = abort_flag;
  for (int spin = 0; spin < (1 << 22); ++spin) {      // ~0.3 s worst case, then everybody bails out
    if (*f) return true;
    if (*ab) return false;
    __nanosleep(64);
  }
  *abort_flag = 1;
  return false;
}

constexpr int WLD = NB + 1;                                        // padded stride of the inverse block in shared memory
constexpr int WAVE_SMEM = NB * WLD * (int)sizeof(double);

// Both kernels keep the CTA's inverse diagonal block in shared memory (loaded up front, it depends on
// nothing) and fetch the next off-diagonal block of L into REGISTERS before waiting for its flag, so the
// per-block critical path holds no global-memory round trip except the hand-off itself.
__global__ void __launch_bounds__(256, 1) trsv_fwd_wave_kernel(int n, const double* __restrict__ L, int ldl,
                                                               const double* __restrict__ invD, const double* __restrict__ b,
                                                               double* __restrict__ out, int* ready, int* abort_flag) {
  extern __shared__ double Xs[];          // invD_c, element (i,j) at Xs[i + j*WLD]
  __shared__ double ys[NB];
  __shared__ double red[2][NB];
  __shared__ int ok;
  const int c = blockIdx.x, tid = threadIdx.x;
  const int r0 = c * NB;
  const int nbc = (n - r0) < NB ? (n - r0) : NB;
  const int row = tid & (NB - 1), half = tid >> 7;
  {
    const double* Xg = invD + (size_t)c * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += 256) Xs[(idx & (NB - 1)) + (idx >> 7) * WLD] = Xg[idx];
  }
  const double bval = (tid < nbc) ? b[r0 + tid] : 0.0;
  double acc = 0.0;
  const bool live = row < nbc;
  for (int k = 0; k < c; ++k) {
    double lreg[64];
    const double* Lr = L + (size_t)(k * NB + half * 64) * ldl + r0 + (live ? row : 0);
#pragma unroll
    for (int j = 0; j < 64; ++j) lreg[j] = Lr[(size_t)j * ldl];      // 64 loads in flight, before the wait
    if (tid == 0) ok = wait_flag(ready + k, abort_flag) ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    if (tid < NB) ys[tid] = __ldcg(out + k * NB + tid);
    __syncthreads();
    const double* yy = ys + half * 64;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      a0 = fma(lreg[j], yy[j], a0);
      a1 = fma(lreg[j + 1], yy[j + 1], a1);
      a2 = fma(lreg[j + 2], yy[j + 2], a2);
      a3 = fma(lreg[j + 3], yy[j + 3], a3);
    }
    if (live) acc += (a0 + a1) + (a2 + a3);
    __syncthreads();
  }
  red[half][row] = acc;
  __syncthreads();
  if (tid < NB) ys[tid] = (tid < nbc) ? bval - (red[0][tid] + red[1][tid]) : 0.0;
  __syncthreads();
  {   // y_c = invD_c * rhs  (lower triangular, zeros above the diagonal)
    const double* xr = Xs + row + (half * 64) * WLD;
    const double* bb = ys + half * 64;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
    for (int j = 0; j < 64; j += 4) {
      a0 = fma(xr[j * WLD], bb[j], a0);
      a1 = fma(xr[(j + 1) * WLD], bb[j + 1], a1);
      a2 = fma(xr[(j + 2) * WLD], bb[j + 2], a2);
      a3 = fma(xr[(j + 3) * WLD], bb[j + 3], a3);
    }
    red[half][row] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (tid < NB) __stcg(out + r0 + tid, (tid < nbc) ? red[0][tid] + red[1][tid] : 0.0);
  __threadfence();
  __syncthreads();
  if (tid == 0) atomicExch(ready + c, 1);
}

__global__ void __launch_bounds__(256, 1) trsv_bwd_wave_kernel(int n, int nblk, const double* __restrict__ L, int ldl,
                                                               const double* __restrict__ invD, const double* __restrict__ y,
                                                               double* __restrict__ out, int* ready, int* abort_flag) {
  extern __shared__ double Xs[];
  __shared__ double xs[NB];
  __shared__ double red[2][NB];
  __shared__ int ok;
  const int c = blockIdx.x, tid = threadIdx.x;
  const int c0 = c * NB;
  const int nbc = (n - c0) < NB ? (n - c0) : NB;
  const int col = tid & (NB - 1), half = tid >> 7;
  {
    const double* Xg = invD + (size_t)c * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += 256) Xs[(idx & (NB - 1)) + (idx >> 7) * WLD] = Xg[idx];
  }
  const double yval = (tid < nbc) ? y[c0 + tid] : 0.0;
  const bool live = col < nbc;
  double acc = 0.0;
  for (int k = nblk - 1; k > c; --k) {
    const int k0 = k * NB;
    const int nbk = (n - k0) < NB ? (n - k0) : NB;
    // column (c0+col) of L, rows k0 + half*64 .. +63: contiguous -> 32 16-byte loads in flight before the wait
    double2 lreg[32];
    const double* Lc = L + (size_t)(c0 + (live ? col : 0)) * ldl + k0 + half * 64;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int r = half * 64 + 2 * i;
      lreg[i] = (r + 1 < nbk) ? *reinterpret_cast<const double2*>(Lc + 2 * i)
                              : make_double2(r < nbk ? Lc[2 * i] : 0.0, 0.0);
    }
    if (tid == 0) ok = wait_flag(ready + k, abort_flag) ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    if (tid < NB) xs[tid] = (tid < nbk) ? __ldcg(out + k0 + tid) : 0.0;
    __syncthreads();
    const double* xx = xs + half * 64;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      a0 = fma(lreg[i].x, xx[2 * i], a0);
      a1 = fma(lreg[i].y, xx[2 * i + 1], a1);
    }
    if (live) acc += a0 + a1;
    __syncthreads();
  }
  red[half][col] = acc;
  __syncthreads();
  if (tid < NB) xs[tid] = (tid < nbc) ? yval - (red[0][tid] + red[1][tid]) : 0.0;     // rhs of the diagonal solve
  __syncthreads();
  {   // x_c(i) = sum_j invD_c(j, i) rhs(j)  (zeros for j < i): thread i walks down column i of the inverse
    const double* xc = Xs + col * WLD + half * 64;
    const double* bb = xs + half * 64;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
    for (int j = 0; j < 64; j += 4) {
      a0 = fma(xc[j], bb[j], a0);
      a1 = fma(xc[j + 1], bb[j + 1], a1);
      a2 = fma(xc[j + 2], bb[j + 2], a2);
      a3 = fma(xc[j + 3], bb[j + 3], a3);
    }
    red[half][col] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (tid < NB) __stcg(out + c0 + tid, (tid < nbc) ? red[0][tid] + red[1][tid] : 0.0);
  __threadfence();
  __syncthreads();
  if (tid == 0) atomicExch(ready + c, 1);
}

int trsv_lower(Handle& h, int n, const double* L, int ldl, const double* invD, double* b, bool trans) {
  double* out = h.d_part + (PART_DOUBLES - 32768);     // tail of the scratch block (n <= 32768)
  if (n > 32768) { set_last_error("trsv_lower: n = %d exceeds the single-RHS scratch", n); return CVXB_EINVAL; }
  const int nblk = (n + NB - 1) / NB;
  if (nblk <= h.sm_count && h.wave_ready) {
    // wavefront kernel: every CTA must be resident at once -> cooperative launch (fails cleanly otherwise)
    CVXB_CUDA_OK(cudaMemsetAsync(h.wave_ready, 0, (size_t)(nblk + 1) * sizeof(int), h.stream));
    int* ready = h.wave_ready;
    int* abort_flag = h.d_flag + F_WAVE_ABORT;
    int nn = n, nb = nblk, ld = ldl;
    cudaError_t e;
    if (!trans) {
      void* args[] = {&nn, (void*)&L, &ld, (void*)&invD, (void*)&b, &out, &ready, &abort_flag};
      e = cudaLaunchCooperativeKernel((void*)trsv_fwd_wave_kernel, dim3(nblk), dim3(256), args, WAVE_SMEM, h.stream);
    } else {
      void* args[] = {&nn, &nb, (void*)&L, &ld, (void*)&invD, (void*)&b, &out, &ready, &abort_flag};
      e = cudaLaunchCooperativeKernel((void*)trsv_bwd_wave_kernel, dim3(nblk), dim3(256), args, WAVE_SMEM, h.stream);
    }
    if (e == cudaSuccess) {
      h.launches++;
      CVXB_LAUNCH(h, copy_vec_kernel, (n + 255) / 256, 256, 0, n, out, b);
      return CVXB_OK;
    }
    cudaGetLastError();       // not launchable cooperatively here: per-block kernels below
  }
  if (!trans) {
    for (int k = 0; k < nblk; ++k) {
      int k0 = k * NB, kb = n - k0 < NB ? n - k0 : NB;
      int rem = n - k0 - kb;
      int grid = rem > 0 ? (rem + TRSV_ROWS - 1) / TRSV_ROWS : 1;
      CVXB_LAUNCH(h, trsv_fwd_step_kernel, grid, 256, 0, n, k0, kb, L, ldl, invD + (size_t)k * NB * NB, b, out);
    }
  } else {
    for (int k = nblk - 1; k >= 0; --k) {
      int k0 = k * NB, kb = n - k0 < NB ? n - k0 : NB;
      int grid = k0 > 0 ? (k0 + 63) / 64 : 1;
      CVXB_LAUNCH(h, trsv_bwd_step_kernel, grid, 256, 0, k0, kb, L, ldl, invD + (size_t)k * NB * NB, b, out);
    }
  }
  CVXB_LAUNCH(h, copy_vec_kernel, (n + 255) / 256, 256, 0, n, out, b);
  return CVXB_OK;
}

bool leaf_attr_set = false;
int leaf_init(bool force = false) {
  if (leaf_attr_set && !force) return CVXB_OK;
  CVXB_CUDA_OK(cudaFuncSetAttribute(leaf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
  CVXB_CUDA_OK(cudaFuncSetAttribute(leaf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
  CVXB_CUDA_OK(cudaFuncSetAttribute(trsv_fwd_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WAVE_SMEM));
  CVXB_CUDA_OK(cudaFuncSetAttribute(trsv_bwd_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WAVE_SMEM));
  leaf_attr_set = true;
  return CVXB_OK;
}

}  // namespace

int factor_init() { return leaf_init(true); }     // per handle: attributes belong to the current device

// the diagonal-block starts of the tile-DAG schedule for an n x n matrix (host logic only; tests/test_abi_cpu.py)
int dag_block_starts(int n, int nbk, int* out, int cap) {
  if (n < 1 || nbk < NB) return -1;
  std::vector<int> st;
  dag_blocks(n, nbk / NB * NB, st);
  for (int i = 0; i < (int)st.size() && i < cap; ++i) out[i] = st[i];
  return (int)st.size();
}

int ruiz_equilibrate(Handle& h, int n, const double* Hm, int ldh, double* d, double* colsq, int max_sweeps, double tol,
                     double* big_scratch, size_t big_doubles) {
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, ruiz_init_kernel, (n + 255) / 256, 256, 0, n, d, h.d_flag, h.d_scal, h.d_ticket);
  if (max_sweeps <= 0) return CVXB_OK;
  // beyond L2: symmetric-half sweeps (needs 2 * ceil(n/128) * pad(n) doubles of scratch; the callers hand over the factor's
  // buffer, which is not in use yet)
  static const int sym_min = getenv("CVXB_RUIZ_SYM_MIN") ? atoi(getenv("CVXB_RUIZ_SYM_MIN")) : 4096;
  static int sym_occ = -1;
  if (sym_occ < 0) {
    int occ = 0;
    if (getenv("CVXB_NO_FUSED_RUIZ") || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ruiz_sym_kernel, 256, 0) != cudaSuccess) {
      cudaGetLastError();
      occ = 0;
    }
    sym_occ = occ;
  }
  const int ldp = pad_ld(n), nt_sym = (n + RS_T - 1) / RS_T;
  if (sym_occ > 0 && n >= sym_min && h.wave_ready && max_sweeps <= RUIZ_MAX_FUSED && big_scratch &&
      big_doubles >= (size_t)2 * nt_sym * ldp && !((uintptr_t)d & 15) && !((uintptr_t)colsq & 15)) {
    int grid = sym_occ * h.sm_count;
    const int T = nt_sym * (nt_sym + 1) / 2;
    if (grid > T) grid = T;
    unsigned long long* rho_bits = (unsigned long long*)(h.d_ticket + 16);
    int nn = n, ld = ldh, ms = max_sweeps, lp = ldp;
    int* flag = h.d_flag;
    double* scal = h.d_scal;
    double tl = tol;
    void* args[] = {&nn, (void*)&Hm, &ld, &d, &colsq, &big_scratch, &lp, &rho_bits, &flag, &scal, &ms, &tl};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)ruiz_sym_kernel, dim3(grid), dim3(256), args, 0, h.stream);
    if (e == cudaSuccess) {
      h.launches++;
      return CVXB_OK;
    }
    cudaGetLastError();       // not launchable cooperatively here: the full-column kernels below
  }
  static int fused_occ = -1;          // CTAs of the fused kernel per SM (0: not usable)
  if (fused_occ < 0) {
    int occ = 0;
    if (getenv("CVXB_NO_FUSED_RUIZ") ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ruiz_fused_kernel, 128, 0) != cudaSuccess) {
      cudaGetLastError();
      occ = 0;
    }
    fused_occ = occ;
  }
  if (fused_occ > 0 && h.wave_ready && max_sweeps <= RUIZ_MAX_FUSED && !(ldh & 1) && !((uintptr_t)Hm & 15) &&
      !((uintptr_t)d & 15) && !((uintptr_t)colsq & 15)) {
    int grid = fused_occ * h.sm_count;
    if (grid > n) grid = n;
    unsigned long long* rho_bits = (unsigned long long*)(h.d_ticket + 16);
    int nn = n, ld = ldh, ms = max_sweeps;
    int* flag = h.d_flag;
    double* scal = h.d_scal;
    double tl = tol;
    void* args[] = {&nn, (void*)&Hm, &ld, &d, &colsq, &rho_bits, &flag, &scal, &ms, &tl};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)ruiz_fused_kernel, dim3(grid), dim3(128), args, 0, h.stream);
    if (e == cudaSuccess) {
      h.launches++;
      return CVXB_OK;
    }
    cudaGetLastError();       // not launchable cooperatively here: per-sweep kernels below
  }
  for (int s = 0; s < max_sweeps; ++s)
    CVXB_LAUNCH(h, ruiz_sweep_kernel, (n + 7) / 8, 256, 0, n, Hm, ldh, d, colsq, h.d_flag, h.d_scal, h.d_ticket,
                max_sweeps, tol);
  return CVXB_OK;
}

static inline int ygrid(int n) { return n < 1 ? 1 : (n > 1024 ? 1024 : n); }

int scaled_lower(Handle& h, int n, const double* Hm, int ldh, const double* d, double delta, double* L, int ldl) {
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, scaled_lower_kernel, dim3((n + 127) / 128, ygrid(n)), 128, 0, n, Hm, ldh, d, delta, L, ldl, 0);
  return CVXB_OK;
}

int scaled_full(Handle& h, int n, const double* Hm, int ldh, const double* d, double* Q, int ldq) {
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, scaled_lower_kernel, dim3((n + 127) / 128, ygrid(n)), 128, 0, n, Hm, ldh, d, 0.0, Q, ldq, 1);
  return CVXB_OK;
}

// Right-looking blocked Cholesky with look-ahead for matrices that stay in L2 (n <= RL_MAX_N): per 128-column
// step the critical chain is  leaf -> panel solve -> update of the NEXT column block  on the main stream, while
// the bulk of the trailing update runs on a second stream, overlapped with the next leaf (which occupies a
// single SM).  Fork / join by events, so the whole schedule is capturable into the per-step CUDA graph.
// Programmatic dependent launch of the two kernels that directly follow another kernel on the critical chain (panel
// GEMM behind the leaf, next leaf behind the look-ahead GEMM): their launch and prologue overlap the predecessor's
// tail (potrf n = 2000 outside a graph: 1.09 -> 1.01 ms).  Inside a captured step the graph's own edges already
// cost no more, so plain launches are kept there.
static bool use_pdl(const Handle& h) {
  static int v = -1;
  if (v < 0) v = getenv("CVXB_NO_PDL") ? 0 : 1;
  return v == 1 && !h.capturing;
}

int potrf_lookahead(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, int col0,
                    double* B, int ldb, int r) {
  const int nblk = (n + NB - 1) / NB;
  const size_t half = (PART_DOUBLES - 65536) / 2;
  const size_t ys_doubles = B ? (size_t)NB * r : 0;     // Y_k = invD_k B_k, kept beside the panel in the same half
  if (2 * nblk + 2 > (int)h.la_events.size() || (size_t)pad_ld(n) * NB + ys_doubles > half)
    return potrf_rec_plain(h, n, A, lda, invD, flag_slot, mindiag_slot, col0, B, ldb, r);
  cudaStream_t sa = h.stream, sb = h.stream2;
  int pending = -1;      // step whose second-stream work (copy-back + bulk update) the main stream has not yet waited for
  int last_k0 = 0, last_kb = 0;
  for (int k = 0; k < nblk; ++k) {
    const int k0 = k * NB, kb = n - k0 < NB ? n - k0 : NB;
    const int rem = n - k0 - kb;
    double* Akk = A + (size_t)k0 * lda + k0;
    if (k > 0 && use_pdl(h)) {
      // leaf k follows the look-ahead GEMM directly on this stream: launch it as its programmatic dependent
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1);
      cfg.blockDim = dim3(LEAF_THREADS);
      cfg.dynamicSmemBytes = LEAF_SMEM;
      cfg.stream = h.stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      CVXB_CUDA_OK(cudaLaunchKernelEx(&cfg, leaf_kernel<true>, kb, kb, Akk, lda, invD + (size_t)k * NB * NB, h.d_flag,
                                      h.d_scal, flag_slot, mindiag_slot, col0 + k0));
      h.launches++;
    } else {
      CVXB_LAUNCH(h, leaf_kernel<true>, 1, LEAF_THREADS, LEAF_SMEM, kb, kb, Akk, lda, invD + (size_t)k * NB * NB, h.d_flag,
                  h.d_scal, flag_slot, mindiag_slot, col0 + k0);
    }
    last_k0 = k0; last_kb = kb;
    if (rem <= 0) break;
    double* A21 = Akk + kb;                                   // rows below the diagonal block, this column block
    // panel solve X = A21 L11^-T into one half of the scratch block (alternating halves): the critical chain reads
    // the panel from there, the copy back into A rides on the second stream
    double* Xs = h.d_part + (size_t)(k & 1) * half;
    const int lds = pad_ld(rem);
    GemmArgs gp{rem, kb, kb, A21, lda, false, invD + (size_t)k * NB * NB, NB, false, Xs, lds, 1.0, 0.0, 0};
    CVXB_TRY(use_pdl(h) ? gemm_dmma_pdl(h, gp) : gemm_dmma(h, gp));      // directly behind the leaf
    const int kn = rem < NB ? rem : NB;                       // width of the next column block
    const int rem2 = rem - kn;
    cudaEvent_t evP = h.la_events[2 * k], evB = h.la_events[2 * k + 1];
    CVXB_CUDA_OK(cudaEventRecord(evP, sa));                   // panel k ready
    if (pending >= 0) CVXB_CUDA_OK(cudaStreamWaitEvent(sa, h.la_events[2 * pending + 1], 0));   // bulk update k-1 done
    // look-ahead: column block k+1 only (rows and columns share the origin k0+kb: keep the upper triangle untouched)
    double* Anext = A + (size_t)(k0 + kb) * lda + (k0 + kb);
    GemmArgs gl{rem, kn, kb, Xs, lds, false, Xs, lds, false, Anext, lda, -1.0, 1.0, 0};
    gl.lower_only = true;
    CVXB_TRY(gemm_dmma(h, gl));
    // second stream: bulk of the trailing update, then the panel's way back into A
    CVXB_CUDA_OK(cudaStreamWaitEvent(sb, evP, 0));
    if (rem2 > 0) {
      const double* X31 = Xs + kn;
      double* A33 = A + (size_t)(k0 + kb + kn) * lda + (k0 + kb + kn);
      GemmArgs gb{rem2, rem2, kb, X31, lds, false, X31, lds, false, A33, lda, -1.0, 1.0, 1};
      CVXB_TRY(gemm_dmma_on(h, gb, sb));
    }
    cudaStream_t keep = h.stream;
    h.stream = sb;
    int st = copy_matrix(h, rem, kb, Xs, lds, A21, lda);
    if (st == CVXB_OK && B) {
      // forward substitution riding along: Y_k = invD_k B_k (B_k is final: every earlier update ran on this
      // stream), B_rest -= L_rest,k Y_k with the panel still in scratch, Y_k back into B
      double* Ys = Xs + (half - ys_doubles);
      GemmArgs gy{kb, r, kb, invD + (size_t)k * NB * NB, NB, false, B + k0, ldb, true, Ys, NB, 1.0, 0.0, 0};
      st = gemm_dmma_on(h, gy, sb);
      if (st == CVXB_OK) {
        GemmArgs gu{rem, r, kb, Xs, lds, false, Ys, NB, true, B + k0 + kb, ldb, -1.0, 1.0, 0};
        st = gemm_dmma_on(h, gu, sb);
      }
      if (st == CVXB_OK) st = copy_matrix(h, kb, r, Ys, NB, B + k0, ldb);
    }
    h.stream = keep;
    if (st != CVXB_OK) return st;
    CVXB_CUDA_OK(cudaEventRecord(evB, sb));
    pending = k;
  }
  if (pending >= 0) CVXB_CUDA_OK(cudaStreamWaitEvent(sa, h.la_events[2 * pending + 1], 0));   // join
  if (B)      // last diagonal block: Y = invD B on the main stream (everything else has joined)
    CVXB_TRY(trsm_rec(h, last_kb, r, A + (size_t)last_k0 * lda + last_k0, lda, invD + (size_t)(last_k0 / NB) * NB * NB,
                      B + last_k0, ldb, false));
  return CVXB_OK;
}

int rl_max_n() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CVXB_RL_MAX_N");
    v = e ? atoi(e) : 4608;
  }
  return v;
}

int potrf_lower(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot) {
  CVXB_TRY(leaf_init());
  CVXB_LAUNCH(h, potrf_reset_kernel, 1, 1, 0, h.d_flag, h.d_scal, flag_slot, mindiag_slot);
  if (n <= 0) return CVXB_OK;
  return potrf_rec(h, n, A, lda, invD, flag_slot, mindiag_slot, 0);
}

int potrf_lower_rhs(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, double* B, int ldb,
                    int r) {
  if (!B || r <= 0 || (ldb & 1) || ((uintptr_t)B & 15) || getenv("CVXB_NO_FUSED_TRSM")) {
    CVXB_TRY(potrf_lower(h, n, A, lda, invD, flag_slot, mindiag_slot));
    if (B && r > 0) return trsm_lower(h, n, r, A, lda, invD, B, ldb, false);
    return CVXB_OK;
  }
  CVXB_TRY(leaf_init());
  CVXB_LAUNCH(h, potrf_reset_kernel, 1, 1, 0, h.d_flag, h.d_scal, flag_slot, mindiag_slot);
  if (n <= 0) return CVXB_OK;
  return potrf_rec(h, n, A, lda, invD, flag_slot, mindiag_slot, 0, false, B, ldb, r);
}

int trsm_lower(Handle& h, int n, int r, const double* L, int ldl, const double* invD, double* B, int ldb, bool trans) {
  if (n <= 0 || r <= 0) return CVXB_OK;
  if (r == 1 && n > NB) return trsv_lower(h, n, L, ldl, invD, B, trans);
  return trsm_rec(h, n, r, L, ldl, invD, B, ldb, trans);
}

int leaf_clocks(long long* out, bool reset) {
  if (out) CVXB_CUDA_OK(cudaMemcpyFromSymbol(out, g_leaf_clk, 8 * sizeof(long long)));
  if (reset) {
    long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CVXB_CUDA_OK(cudaMemcpyToSymbol(g_leaf_clk, z, sizeof(z)));
  }
  return CVXB_OK;
}

int invert_diag_blocks(Handle& h, int n, const double* L, int ldl, double* invD) {
  CVXB_TRY(leaf_init());
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, leaf_kernel<false>, (n + NB - 1) / NB, LEAF_THREADS, LEAF_SMEM, n, 0, const_cast<double*>(L), ldl, invD,
              h.d_flag, h.d_scal, 0, 0, 0);
  return CVXB_OK;
}

}  // namespace cvxb

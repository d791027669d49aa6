// HBM-bound BLAS-2 and matrix-elementwise kernels of the Newton step (SURVEY.md K1, K3, K5 parts):
//   gemv_n   y = alpha*A x + beta*y     G x (slacks, BarrierSolver.scala:283), A x (b - Ax), P x, Y w
//   gemv_t   y = alpha*A'x + beta*y     G'(1/d) (gradient, BarrierSolver.scala:291-301), A'w, Y'y
//   scale_rows / fill_matrix / transpose_scale / add_diag / copy_matrix
// All matrices are column-major, so consecutive threads walk DOWN a column (coalesced 16-byte loads).
// Reductions are deterministic (fixed trees, no atomics): bit-reproducible run to run.
#include "common.cuh"

namespace cvxb {
namespace {

constexpr int GN_ROWS = 64;     // rows per CTA in gemv_n (each lane owns 2 consecutive rows)
constexpr int GN_WARPS = 8;

// partial[s][i] = sum_{j in split s} A(i,j) x_j ; S == 1 writes alpha*sum + beta*y directly
__global__ void __launch_bounds__(256) gemv_n_kernel(int m, int n, double alpha, const double* __restrict__ A, int lda,
                                                     const double* __restrict__ x, double beta, double* __restrict__ y,
                                                     double* __restrict__ part, int cols_per_split, int S) {
  __shared__ double red[GN_WARPS][GN_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * GN_ROWS + 2 * lane;
  const int c0 = blockIdx.y * cols_per_split;
  int c1 = c0 + cols_per_split;
  if (c1 > n) c1 = n;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  if (r + 1 < m) {
    int j = c0 + warp;
    // four columns (16-byte loads) in flight per lane: the kernel is a latency-bound stream otherwise
    double e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    for (; j + 3 * GN_WARPS < c1; j += 4 * GN_WARPS) {
      double2 v = *reinterpret_cast<const double2*>(A + (size_t)j * lda + r);
      double2 w = *reinterpret_cast<const double2*>(A + (size_t)(j + GN_WARPS) * lda + r);
      double2 u = *reinterpret_cast<const double2*>(A + (size_t)(j + 2 * GN_WARPS) * lda + r);
      double2 z = *reinterpret_cast<const double2*>(A + (size_t)(j + 3 * GN_WARPS) * lda + r);
      double xj = x[j], xk = x[j + GN_WARPS], xl = x[j + 2 * GN_WARPS], xm = x[j + 3 * GN_WARPS];
      a0 = fma(v.x, xj, a0); a1 = fma(v.y, xj, a1);
      b0 = fma(w.x, xk, b0); b1 = fma(w.y, xk, b1);
      e0 = fma(u.x, xl, e0); e1 = fma(u.y, xl, e1);
      f0 = fma(z.x, xm, f0); f1 = fma(z.y, xm, f1);
    }
    a0 += e0; a1 += e1; b0 += f0; b1 += f1;
    for (; j + GN_WARPS < c1; j += 2 * GN_WARPS) {
      double2 v = *reinterpret_cast<const double2*>(A + (size_t)j * lda + r);
      double2 w = *reinterpret_cast<const double2*>(A + (size_t)(j + GN_WARPS) * lda + r);
      double xj = x[j], xk = x[j + GN_WARPS];
      a0 = fma(v.x, xj, a0); a1 = fma(v.y, xj, a1);
      b0 = fma(w.x, xk, b0); b1 = fma(w.y, xk, b1);
    }
    for (; j < c1; j += GN_WARPS) {
      double2 v = *reinterpret_cast<const double2*>(A + (size_t)j * lda + r);
      double xj = x[j];
      a0 = fma(v.x, xj, a0); a1 = fma(v.y, xj, a1);
    }
  } else if (r < m) {
    for (int j = c0 + warp; j < c1; j += GN_WARPS) a0 = fma(A[(size_t)j * lda + r], x[j], a0);
  }
  red[warp][2 * lane] = a0 + b0;
  red[warp][2 * lane + 1] = a1 + b1;
  __syncthreads();
  if (threadIdx.x < GN_ROWS) {
    int i = blockIdx.x * GN_ROWS + threadIdx.x;
    if (i < m) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < GN_WARPS; ++w) s += red[w][threadIdx.x];
      if (S == 1) y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * s;
      else part[(size_t)blockIdx.y * m + i] = s;
    }
  }
}

__global__ void gemv_n_finish(int m, int S, double alpha, const double* __restrict__ part, double beta,
                              double* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = 0;
  for (int k = 0; k < S; ++k) s += part[(size_t)k * m + i];
  y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * s;
}

// one warp per column
__global__ void __launch_bounds__(256) gemv_t_kernel(int m, int n, double alpha, const double* __restrict__ A, int lda,
                                                     const double* __restrict__ x, double beta, double* __restrict__ y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  if (j >= n) return;
  const double* col = A + (size_t)j * lda;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const int m2 = m & ~1;
  int i = 2 * lane;
  {   // four 16-byte loads of the column (and of x) in flight per lane
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (; i + 192 < m2; i += 256) {
      double2 v = *reinterpret_cast<const double2*>(col + i);
      double2 w = *reinterpret_cast<const double2*>(col + i + 64);
      double2 u = *reinterpret_cast<const double2*>(col + i + 128);
      double2 z = *reinterpret_cast<const double2*>(col + i + 192);
      double2 xv = *reinterpret_cast<const double2*>(x + i);
      double2 xw = *reinterpret_cast<const double2*>(x + i + 64);
      double2 xu = *reinterpret_cast<const double2*>(x + i + 128);
      double2 xz = *reinterpret_cast<const double2*>(x + i + 192);
      s0 = fma(v.x, xv.x, s0); s1 = fma(v.y, xv.y, s1);
      s2 = fma(w.x, xw.x, s2); s3 = fma(w.y, xw.y, s3);
      t0 = fma(u.x, xu.x, t0); t1 = fma(u.y, xu.y, t1);
      t2 = fma(z.x, xz.x, t2); t3 = fma(z.y, xz.y, t3);
    }
    s0 += t0; s1 += t1; s2 += t2; s3 += t3;
  }
  for (; i + 64 < m2; i += 128) {
    double2 v = *reinterpret_cast<const double2*>(col + i);
    double2 w = *reinterpret_cast<const double2*>(col + i + 64);
    double2 xv = *reinterpret_cast<const double2*>(x + i);
    double2 xw = *reinterpret_cast<const double2*>(x + i + 64);
    s0 = fma(v.x, xv.x, s0); s1 = fma(v.y, xv.y, s1);
    s2 = fma(w.x, xw.x, s2); s3 = fma(w.y, xw.y, s3);
  }
  for (; i < m2; i += 64) {
    double2 v = *reinterpret_cast<const double2*>(col + i);
    double2 xv = *reinterpret_cast<const double2*>(x + i);
    s0 = fma(v.x, xv.x, s0); s1 = fma(v.y, xv.y, s1);
  }
  if (lane == 0 && (m & 1)) s2 = fma(col[m - 1], x[m - 1], s2);
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[j] = (beta == 0.0 ? 0.0 : beta * y[j]) + alpha * s;
}

__global__ void scale_rows_kernel(int m, int n, const double* __restrict__ G, int ldg, const double* __restrict__ s,
                                  double* __restrict__ Gs, int ldgs, int sqrt_of_s) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= m) return;
  double s0 = s[i], s1 = (i + 1 < m) ? s[i + 1] : 0.0;
  if (sqrt_of_s) { s0 = sqrt(s0); s1 = sqrt(s1); }
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    if (i + 1 < m) {
      double2 v = *reinterpret_cast<const double2*>(G + (size_t)j * ldg + i);
      v.x *= s0; v.y *= s1;
      *reinterpret_cast<double2*>(Gs + (size_t)j * ldgs + i) = v;
    } else {
      Gs[(size_t)j * ldgs + i] = G[(size_t)j * ldg + i] * s0;
    }
  }
}

// C = alpha*A (A NULL -> 0); diagonal += diag_scale / diag_den[i]  (KL: t * diag(1/x), Dist_KL.scala:236-239)
__global__ void fill_matrix_kernel(int n, double alpha, const double* __restrict__ A, int lda,
                                   const double* __restrict__ diag_den, double diag_scale, double* __restrict__ C,
                                   int ldc, const double* __restrict__ mul_dev, double diag_exp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mul_dev) { alpha *= *mul_dev; diag_scale *= *mul_dev; }   // barrier parameter t read on the device (graph replay)
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    double v = A ? alpha * A[(size_t)j * lda + i] : 0.0;
    if (i == j && diag_den) v += (diag_exp == -1.0) ? diag_scale / diag_den[i] : diag_scale * pow(fabs(diag_den[i]), diag_exp);
    C[(size_t)j * ldc + i] = v;
  }
}

__global__ void transpose_scale_kernel(int p, int n, const double* __restrict__ A, int lda, const double* __restrict__ s,
                                       double* __restrict__ Bt, int ldbt) {
  __shared__ double tile[32][33];
  int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;   // j over p (rows of A), i over n (cols of A)
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    int j = j0 + threadIdx.x, i = i0 + k;
    tile[k][threadIdx.x] = (j < p && i < n) ? A[(size_t)i * lda + j] : 0.0;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    int i = i0 + threadIdx.x, j = j0 + k;
    if (i < n && j < p) Bt[(size_t)j * ldbt + i] = tile[threadIdx.x][k] * (s ? s[i] : 1.0);
  }
}

__global__ void add_diag_kernel(int n, double alpha, double* C, int ldc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) C[(size_t)i * ldc + i] += alpha;
}

__global__ void copy_matrix_kernel(int m, int n, const double* __restrict__ A, int lda, double* __restrict__ B, int ldb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int j = blockIdx.y; j < n; j += gridDim.y) B[(size_t)j * ldb + i] = A[(size_t)j * lda + i];
}

}  // namespace

int gemv_n(Handle& h, int m, int n, double alpha, const double* A, int lda, const double* x, double beta, double* y) {
  if (m <= 0) return CVXB_OK;
  if (n <= 0) {
    n = 0;
  }
  if ((lda & 1) || ((uintptr_t)A & 15)) {
    set_last_error("gemv_n: A must be 16-byte aligned with an even leading dimension");
    return CVXB_EINVAL;
  }
  int rb = (m + GN_ROWS - 1) / GN_ROWS;
  int S = 1;
  int want = 2 * h.sm_count;
  if (rb < want && n > 128) {
    S = (want + rb - 1) / rb;
    int maxS = (n + 63) / 64;
    if (S > maxS) S = maxS;
    if ((size_t)S * m > PART_DOUBLES) S = (int)(PART_DOUBLES / m);
    if (S < 1) S = 1;
  }
  int cps = ((n + S - 1) / S + GN_WARPS - 1) / GN_WARPS * GN_WARPS;
  if (cps < GN_WARPS) cps = GN_WARPS;
  S = n > 0 ? (n + cps - 1) / cps : 1;
  CVXB_LAUNCH(h, gemv_n_kernel, dim3(rb, S), 256, 0, m, n, alpha, A, lda, x, beta, y, h.d_part, cps, S);
  if (S > 1) CVXB_LAUNCH(h, gemv_n_finish, (m + 255) / 256, 256, 0, m, S, alpha, h.d_part, beta, y);
  return CVXB_OK;
}

int gemv_t(Handle& h, int m, int n, double alpha, const double* A, int lda, const double* x, double beta, double* y) {
  if (n <= 0) return CVXB_OK;
  if ((lda & 1) || ((uintptr_t)A & 15) || ((uintptr_t)x & 15)) {
    set_last_error("gemv_t: A and x must be 16-byte aligned, lda even");
    return CVXB_EINVAL;
  }
  CVXB_LAUNCH(h, gemv_t_kernel, (n + 7) / 8, 256, 0, m, n, alpha, A, lda, x, beta, y);
  return CVXB_OK;
}

static inline int ygrid(int n) { return n < 1 ? 1 : (n > 1024 ? 1024 : n); }

int scale_rows(Handle& h, int m, int n, const double* G, int ldg, const double* s, double* Gs, int ldgs, bool sqrt_of_s) {
  if (m <= 0 || n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, scale_rows_kernel, dim3((m / 2 + 1 + 127) / 128, ygrid(n)), 128, 0, m, n, G, ldg, s, Gs, ldgs,
              sqrt_of_s ? 1 : 0);
  return CVXB_OK;
}

int fill_matrix(Handle& h, int n, double alpha, const double* A, int lda, const double* diag_den, double diag_scale,
                double* C, int ldc, const double* mul_dev, double diag_exp) {
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, fill_matrix_kernel, dim3((n + 127) / 128, ygrid(n)), 128, 0, n, alpha, A, lda, diag_den, diag_scale, C,
              ldc, mul_dev, diag_exp);
  return CVXB_OK;
}

int transpose_scale(Handle& h, int p, int n, const double* A, int lda, const double* s, double* Bt, int ldbt) {
  if (p <= 0 || n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, transpose_scale_kernel, dim3((p + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, p, n, A, lda, s, Bt, ldbt);
  return CVXB_OK;
}

int add_diag(Handle& h, int n, double alpha, double* C, int ldc) {
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, add_diag_kernel, (n + 255) / 256, 256, 0, n, alpha, C, ldc);
  return CVXB_OK;
}

int copy_matrix(Handle& h, int m, int n, const double* A, int lda, double* B, int ldb) {
  if (m <= 0 || n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, copy_matrix_kernel, dim3((m + 127) / 128, ygrid(n)), 128, 0, m, n, A, lda, B, ldb);
  return CVXB_OK;
}

}  // namespace cvxb

// Equilibration, Cholesky and triangular solves of the KKT path.
//   ruiz_equilibrate   MatrixUtils.ruizEquilibrate        MatrixUtils.scala:240-268
//   potrf_lower        Breeze cholesky -> LAPACK dpotrf   MatrixUtils.scala:452-461, KKTSystem.scala:140
//   trsm_lower         LAPACK dtrtrs / forwardSolve / backSolve   MatrixUtils.scala:362-430
// The factorisation is the recursive (cache-oblivious) right-looking Cholesky: every flop outside the
// NB x NB diagonal leaves is a large-K DMMA GEMM/SYRK (gemm_dmma.cu).  A leaf CTA factors its block in
// shared memory and also writes the block's triangular inverse, so that every leaf triangular solve is
// a DMMA GEMM with the inverse too (same device as MAGMA's trsm with inverted diagonal blocks).
#include "common.cuh"

namespace cvxb {
namespace {

// ------------------------------------------------------------------------------------------- Ruiz
__global__ void ruiz_init_kernel(int n, double* d, int* flag, double* scal, unsigned* ticket) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = 1.0;
  if (i == 0) {
    flag[F_RUIZ_DONE] = 0;
    flag[F_RUIZ_SWEEPS] = 0;
    scal[S_RUIZ_RHO] = 1.0;
    ticket[0] = 0;
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
    double s0 = 0, s1 = 0;
    int i = lane;
    for (; i + 32 < n; i += 64) {
      double q0 = (d[i] * dj) * col[i];
      double q1 = (d[i + 32] * dj) * col[i + 32];
      s0 = fma(q0, q0, s0);
      s1 = fma(q1, q1, s1);
    }
    for (; i < n; i += 32) {
      double q0 = (d[i] * dj) * col[i];
      s0 = fma(q0, q0, s0);
    }
    double s = s0 + s1;
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
constexpr int LDW = NB + 1;                                   // odd stride: conflict-free rows and columns
constexpr int LEAF_SMEM = (NB + 1) * LDW * (int)sizeof(double);   // L (lower) + inverse (transposed, upper, shifted)
constexpr int LEAF_THREADS = 1024;

// W holds L(i,j), i >= j, at W[i + j*LDW]; the inverse X(i,j), i >= j, at W[j + (i+1)*LDW].
#define LW(i, j) W[(i) + (j) * LDW]
#define XW(i, j) W[(j) + ((i) + 1) * LDW]

// blockIdx.x selects the diagonal block when only inverting (FACTOR == false).
template <bool FACTOR>
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_kernel(int n_total, int nb_first, double* A, int lda, double* invD, int* flag, double* scal, int flag_slot,
            int mindiag_slot, int col0) {
  extern __shared__ double W[];
  __shared__ double s_piv;
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;        // 32 x 32
  int nb, off;
  if (FACTOR) { nb = nb_first; off = 0; }
  else {
    off = blockIdx.x * NB;
    nb = n_total - off;
    if (nb > NB) nb = NB;
  }
  double* Ab = A + (size_t)off * lda + off;
  double* Xg = invD + (size_t)(FACTOR ? 0 : blockIdx.x) * NB * NB;
  for (int j = ty; j < nb; j += 32)
    for (int i = tx; i < nb; i += 32)
      if (i >= j) LW(i, j) = Ab[(size_t)j * lda + i];
  if (tid == 0) s_fail = 0;
  __syncthreads();

  if (FACTOR) {
    double mind = 1e300;
    for (int j = 0; j < nb; ++j) {
      if (tid == 0) {
        double pv = LW(j, j);
        if (!(pv > 0.0)) {       // also catches NaN, as dpotrf's  ajj <= 0 || isnan(ajj)
          if (!s_fail) s_fail = col0 + j + 1;
          pv = 1.0;
        }
        double l = sqrt(pv);
        LW(j, j) = l;
        s_piv = 1.0 / l;
        if (l < mind) mind = l;
      }
      __syncthreads();
      const double r = s_piv;
      for (int i = j + 1 + tid; i < nb; i += LEAF_THREADS) LW(i, j) *= r;
      __syncthreads();
      // trailing update of the lower triangle: (i,c), c > j, i >= c
      for (int c = j + 1 + ty; c < nb; c += 32) {
        const double lc = LW(c, j);
        for (int i = c + tx; i < nb; i += 32) LW(i, c) = fma(-LW(i, j), lc, LW(i, c));
      }
      __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
      if (s_fail && flag[flag_slot] == 0) flag[flag_slot] = s_fail;
      if (mind < scal[mindiag_slot]) scal[mindiag_slot] = mind;
    }
  }

  // ---- triangular inverse: right-looking forward substitution on R = I
  for (int j = ty; j < nb; j += 32)
    for (int i = tx; i < nb; i += 32)
      if (i >= j) XW(i, j) = (i == j) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    if (tid == 0) {
      double pv = LW(k, k);
      if (pv == 0.0) { flag[F_ZERO_DIAG] = 1; pv = 1.0; }
      s_piv = 1.0 / pv;
    }
    __syncthreads();
    const double r = s_piv;
    for (int j = tid; j <= k; j += LEAF_THREADS) XW(k, j) *= r;
    __syncthreads();
    for (int i = k + 1 + ty; i < nb; i += 32) {
      const double lik = LW(i, k);
      for (int j = tx; j <= k; j += 32) XW(i, j) = fma(-lik, XW(k, j), XW(i, j));
    }
    __syncthreads();
  }
  __syncthreads();
  for (int j = ty; j < NB; j += 32)
    for (int i = tx; i < NB; i += 32) {
      double x = (i < nb && j < nb && i >= j) ? XW(i, j) : 0.0;
      Xg[(size_t)j * NB + i] = x;
      if (FACTOR && i < nb && j < nb && i >= j) Ab[(size_t)j * lda + i] = LW(i, j);
    }
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
    // in place: the tile grid has a single column (N = n1 <= 128), so a CTA reads only the rows it writes
    GemmArgs g{M, n1, n1, A21, lda, false, invD, NB, false, A21, lda, 1.0, 0.0, 0};
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

int potrf_rec(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, int col0) {
  if (n <= NB) {
    CVXB_LAUNCH(h, leaf_kernel<true>, 1, LEAF_THREADS, LEAF_SMEM, n, n, A, lda, invD, h.d_flag, h.d_scal, flag_slot,
                mindiag_slot, col0);
    return CVXB_OK;
  }
  int a = split_point(n), b = n - a;
  CVXB_TRY(potrf_rec(h, a, A, lda, invD, flag_slot, mindiag_slot, col0));
  double* A21 = A + a;
  double* A22 = A + (size_t)a * lda + a;
  CVXB_TRY(trsm_right_lt(h, b, a, A, lda, invD, A21, lda));
  // A22 -= A21 A21'  lower: A(m,k) = A21[k*lda + m] (M contiguous), B(k,n) = A21(n,k) (N contiguous)
  GemmArgs g{b, b, a, A21, lda, false, A21, lda, false, A22, lda, -1.0, 1.0, 1};
  CVXB_TRY(gemm_dmma(h, g));
  return potrf_rec(h, b, A22, lda, invD + (size_t)(a / NB) * NB * NB, flag_slot, mindiag_slot, col0 + a);
}

int trsm_rec(Handle& h, int n, int r, const double* L, int ldl, const double* invD, double* B, int ldb, bool trans) {
  if (n <= NB) {
    // in place: single tile row (M = n <= 128): a CTA reads only the columns it writes
    GemmArgs g{n, r, n, invD, NB, trans, B, ldb, true, B, ldb, 1.0, 0.0, 0};
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

bool leaf_attr_set = false;
int leaf_init() {
  if (leaf_attr_set) return CVXB_OK;
  CVXB_CUDA_OK(cudaFuncSetAttribute(leaf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
  CVXB_CUDA_OK(cudaFuncSetAttribute(leaf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
  leaf_attr_set = true;
  return CVXB_OK;
}

}  // namespace

int ruiz_equilibrate(Handle& h, int n, const double* Hm, int ldh, double* d, double* colsq, int max_sweeps, double tol) {
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, ruiz_init_kernel, (n + 255) / 256, 256, 0, n, d, h.d_flag, h.d_scal, h.d_ticket);
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

int potrf_lower(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot) {
  CVXB_TRY(leaf_init());
  CVXB_LAUNCH(h, potrf_reset_kernel, 1, 1, 0, h.d_flag, h.d_scal, flag_slot, mindiag_slot);
  if (n <= 0) return CVXB_OK;
  return potrf_rec(h, n, A, lda, invD, flag_slot, mindiag_slot, 0);
}

int trsm_lower(Handle& h, int n, int r, const double* L, int ldl, const double* invD, double* B, int ldb, bool trans) {
  if (n <= 0 || r <= 0) return CVXB_OK;
  return trsm_rec(h, n, r, L, ldl, invD, B, ldb, trans);
}

int invert_diag_blocks(Handle& h, int n, const double* L, int ldl, double* invD) {
  CVXB_TRY(leaf_init());
  if (n <= 0) return CVXB_OK;
  CVXB_LAUNCH(h, leaf_kernel<false>, (n + NB - 1) / NB, LEAF_THREADS, LEAF_SMEM, n, 0, const_cast<double*>(L), ldl, invD,
              h.d_flag, h.d_scal, 0, 0, 0);
  return CVXB_OK;
}

}  // namespace cvxb

// Householder QR of a tall matrix and the solution space of an underdetermined system:
//   MatrixUtils.solveUnderdetermined (MatrixUtils.scala:536-550):  A' = QR (Breeze qr -> LAPACK dgeqrf + dorgqr),
//   F = Q(:, p..n-1) (orthonormal basis of ker A), y = forwardSolve(R', b), z0 = Q(:, 0..p-1) y (minimum-norm solution);
//   SolutionSpace (SolutionSpace.scala:20-33): x = z0 + F u, parameter(x0) = F'(x0 - z0).
//
// Blocked compact-WY Householder (LAPACK's dlarfg / dlarft / dlarfb conventions, so Q agrees with dorgqr's up to
// rounding): a one-CTA panel kernel factors 32 columns and builds the explicit reflector block V (unit lower
// trapezoidal) and its triangular factor T; the trailing update and the accumulation of Q are DMMA GEMMs
// (gemm_dmma).  This is set-up work (once per equality system), not part of the Newton step; the panel kernel is
// latency-bound by its 32 column sweeps and is sized accordingly (one CTA, operands L2-resident).
#include "kkt.cuh"

namespace cvxb {

namespace {

constexpr int QR_NB = 32;
constexpr int QR_THREADS = 1024;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Panel = rows 0..r-1, columns 0..jb-1 at Mp (leading dimension ldm).  On exit: R in the upper triangle, the
// scaled reflectors below it (as dgeqr2 leaves them), V (r x jb) = explicit unit-lower-trapezoidal reflector block,
// T (QR_NB x QR_NB, leading dimension QR_NB) = upper-triangular block-reflector factor with zeros below.
__global__ void __launch_bounds__(QR_THREADS) qr_panel_kernel(int r, int jb, double* __restrict__ Mp, int ldm,
                                                              double* __restrict__ V, int ldv, double* __restrict__ T) {
  __shared__ double red[32];
  __shared__ double Ts[QR_NB][QR_NB + 1];
  __shared__ double z[QR_NB];
  __shared__ double bc[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < QR_NB * QR_NB; i += QR_THREADS) Ts[i / QR_NB][i % QR_NB] = 0.0;
  __syncthreads();
  for (int k = 0; k < jb; ++k) {
    double* colk = Mp + (size_t)k * ldm;
    // dlarfg: sigma = sum_{i>k} x_i^2
    double s = 0.0;
    for (int i = k + 1 + tid; i < r; i += QR_THREADS) { double v = colk[i]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
      double t = red[lane];
      t = warp_sum(t);
      if (lane == 0) {
        const double alpha = colk[k], xnorm = sqrt(t);
        double tau = 0.0, scale = 0.0, beta = alpha;
        if (xnorm != 0.0) {
          beta = -copysign(hypot(alpha, xnorm), alpha);
          tau = (beta - alpha) / beta;
          scale = 1.0 / (alpha - beta);
        }
        bc[0] = tau;
        bc[1] = scale;
        colk[k] = beta;
      }
    }
    __syncthreads();
    const double tau = bc[0], scale = bc[1];
    double* vk = V + (size_t)k * ldv;
    for (int i = tid; i < r; i += QR_THREADS) {
      double v = 0.0;
      if (i == k) v = 1.0;
      else if (i > k) { v = colk[i] * scale; colk[i] = v; }
      vk[i] = v;
    }
    __syncthreads();
    // warp c > k: apply H_k to column c.  warp c < k: z_c = V(:,c)' v_k  (for T)
    if (warp < jb && warp != k) {
      if (warp > k) {
        double* cc = Mp + (size_t)warp * ldm;
        double d = 0.0;
#pragma unroll 4
        for (int i = k + lane; i < r; i += 32) d = fma(vk[i], cc[i], d);
        d = warp_sum(d) * tau;
#pragma unroll 4
        for (int i = k + lane; i < r; i += 32) cc[i] = fma(-d, vk[i], cc[i]);
      } else {
        const double* vc = V + (size_t)warp * ldv;
        double d = 0.0;
#pragma unroll 4
        for (int i = k + lane; i < r; i += 32) d = fma(vc[i], vk[i], d);
        d = warp_sum(d);
        if (lane == 0) z[warp] = d;
      }
    }
    __syncthreads();
    // dlarft (forward, columnwise): T(0:k,k) = -tau * T(0:k,0:k) * z ; T(k,k) = tau
    if (warp == 0) {
      if (lane < k) {
        double acc = 0.0;
        for (int l = lane; l < k; ++l) acc = fma(Ts[lane][l], z[l], acc);
        Ts[lane][k] = -tau * acc;
      } else if (lane == k) {
        Ts[k][k] = tau;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < QR_NB * QR_NB; i += QR_THREADS) T[i] = Ts[i % QR_NB][i / QR_NB];     // column-major
}

__global__ void qr_identity_kernel(int n, double* __restrict__ Q, int ldq) {
  const int j = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) Q[(size_t)j * ldq + i] = (i == j) ? 1.0 : 0.0;
}

// forwardSolve(R', b) (MatrixUtils.scala:383-403): R' lower triangular, (R')(k,i) = R(i,k), column k of R contiguous.
// One CTA; flag = 1 when a diagonal entry is zero (the reference's assert).
__global__ void __launch_bounds__(QR_THREADS) rt_forward_kernel(int p, const double* __restrict__ R, int ldr,
                                                                const double* __restrict__ b, double* __restrict__ y, int* flag) {
  __shared__ double red[32];
  __shared__ double yk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = 0; k < p; ++k) {
    const double* col = R + (size_t)k * ldr;
    double s = 0.0;
    for (int i = tid; i < k; i += QR_THREADS) s = fma(col[i], y[i], s);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
      double t = warp_sum(red[lane]);
      if (lane == 0) {
        const double d = col[k];
        if (d == 0.0) *flag = 1;
        yk = (b[k] - t) / d;
        y[k] = yk;
      }
    }
    __syncthreads();
  }
}

__global__ void sub_vec_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = a[i] - b[i];
}
__global__ void add_vec_kernel(int n, const double* __restrict__ a, double* __restrict__ c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] += a[i];
}

}  // namespace

void solution_space_free(SolutionSpaceDev* S) {
  if (!S) return;
  if (S->Q) cudaFree(S->Q);
  if (S->Fext) cudaFree(S->Fext);
  if (S->z0) cudaFree(S->z0);
  if (S->tmp) cudaFree(S->tmp);
  delete S;
}

// QR of A' (n x p) and the explicit Q (n x n), z0.  A: p x n column-major on the device (leading dimension lda).
int solution_space_build(Handle& h, int p, int n, const double* A, int lda, const double* b, SolutionSpaceDev** out) {
  if (!(p >= 1 && p < n)) { set_last_error("SolutionSpace: need 1 <= A.rows < A.cols (got %d x %d)", p, n); return CVXB_EDIM; }
  SolutionSpaceDev* S = new SolutionSpaceDev();
  S->n = n; S->p = p; S->ldq = pad_ld(n);
  S->device = h.device; S->stream = h.stream;
  const int ldm = pad_ld(n);
  double *M = nullptr, *V = nullptr, *T = nullptr, *W = nullptr;
  int* d_flag = nullptr;
  const int wcols = n > p ? n : p;
  auto cleanup = [&]() { cudaFree(M); cudaFree(V); cudaFree(T); cudaFree(W); cudaFree(d_flag); };
  auto fail = [&](int st) { cleanup(); solution_space_free(S); return st; };
#define QR_OK(call)                                                                                          \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess) { set_last_error("CUDA error %s in SolutionSpace", cudaGetErrorString(e_)); return fail(CVXB_ECUDA); } \
  } while (0)
#define QR_TRY(call)                          \
  do {                                        \
    int s_ = (call);                          \
    if (s_ != CVXB_OK) return fail(s_);       \
  } while (0)
  QR_OK(cudaMalloc((void**)&M, sizeof(double) * (size_t)ldm * p));
  QR_OK(cudaMalloc((void**)&V, sizeof(double) * (size_t)ldm * p));
  QR_OK(cudaMalloc((void**)&T, sizeof(double) * QR_NB * (size_t)round_up(p, QR_NB)));
  QR_OK(cudaMalloc((void**)&W, sizeof(double) * 2 * QR_NB * (size_t)wcols));
  QR_OK(cudaMalloc((void**)&d_flag, sizeof(int)));
  QR_OK(cudaMalloc((void**)&S->Q, sizeof(double) * (size_t)S->ldq * n));
  QR_OK(cudaMalloc((void**)&S->z0, sizeof(double) * (size_t)pad_ld(n)));
  QR_OK(cudaMalloc((void**)&S->tmp, sizeof(double) * 2 * (size_t)pad_ld(n)));
  QR_OK(cudaMemsetAsync(M, 0, sizeof(double) * (size_t)ldm * p, h.stream));
  QR_OK(cudaMemsetAsync(V, 0, sizeof(double) * (size_t)ldm * p, h.stream));
  QR_OK(cudaMemsetAsync(d_flag, 0, sizeof(int), h.stream));
  QR_OK(cudaMemsetAsync(S->tmp, 0, sizeof(double) * 2 * (size_t)pad_ld(n), h.stream));
  QR_TRY(transpose_scale(h, p, n, A, lda, nullptr, M, ldm));          // M = A'
  double* W2 = W + (size_t)QR_NB * wcols;
  // ---- factorisation
  for (int j0 = 0; j0 < p; j0 += QR_NB) {
    const int jb = (p - j0 < QR_NB) ? p - j0 : QR_NB, r = n - j0, c2 = p - j0 - jb;
    double* Mp = M + j0 + (size_t)j0 * ldm;
    double* Vp = V + j0 + (size_t)j0 * ldm;
    double* Tp = T + (size_t)j0 * QR_NB;
    qr_panel_kernel<<<1, QR_THREADS, 0, h.stream>>>(r, jb, Mp, ldm, Vp, ldm, Tp);
    h.launches++;
    QR_OK(cudaGetLastError());
    if (c2 > 0) {       // M2 <- (I - V T' V') M2
      double* M2 = Mp + (size_t)jb * ldm;
      GemmArgs g1{jb, c2, r, Vp, ldm, true, M2, ldm, true, W, QR_NB, 1.0, 0.0, 0};
      QR_TRY(gemm_dmma(h, g1));
      GemmArgs g2{jb, c2, jb, Tp, QR_NB, true, W, QR_NB, true, W2, QR_NB, 1.0, 0.0, 0};
      QR_TRY(gemm_dmma(h, g2));
      GemmArgs g3{r, c2, jb, Vp, ldm, false, W2, QR_NB, true, M2, ldm, -1.0, 1.0, 0};
      QR_TRY(gemm_dmma(h, g3));
    }
  }
  // ---- Q = H_1 ... H_p applied to I, panels in reverse (dorgqr)
  qr_identity_kernel<<<n, 256, 0, h.stream>>>(n, S->Q, S->ldq);
  h.launches++;
  QR_OK(cudaGetLastError());
  const int last = ((p - 1) / QR_NB) * QR_NB;
  for (int j0 = last; j0 >= 0; j0 -= QR_NB) {
    const int jb = (p - j0 < QR_NB) ? p - j0 : QR_NB, r = n - j0;
    double* Vp = V + j0 + (size_t)j0 * ldm;
    double* Tp = T + (size_t)j0 * QR_NB;
    double* Qs = S->Q + j0 + (size_t)j0 * S->ldq;                    // Q(j0:, j0:) <- (I - V T V') Q(j0:, j0:)
    GemmArgs g1{jb, r, r, Vp, ldm, true, Qs, S->ldq, true, W, QR_NB, 1.0, 0.0, 0};
    QR_TRY(gemm_dmma(h, g1));
    GemmArgs g2{jb, r, jb, Tp, QR_NB, false, W, QR_NB, true, W2, QR_NB, 1.0, 0.0, 0};
    QR_TRY(gemm_dmma(h, g2));
    GemmArgs g3{r, r, jb, Vp, ldm, false, W2, QR_NB, true, Qs, S->ldq, -1.0, 1.0, 0};
    QR_TRY(gemm_dmma(h, g3));
  }
  // ---- z0 = Q(:, 0:p) * forwardSolve(R', b)
  double* y = S->tmp;
  rt_forward_kernel<<<1, QR_THREADS, 0, h.stream>>>(p, M, ldm, b, y, d_flag);
  h.launches++;
  QR_OK(cudaGetLastError());
  QR_TRY(gemv_n(h, n, p, 1.0, S->Q, S->ldq, y, 0.0, S->z0));
  int flag = 0;
  QR_OK(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h.stream));
  QR_OK(cudaStreamSynchronize(h.stream));
  if (flag) { set_last_error("forwardSolve: zero on the diagonal of R' (A not of full rank)"); return fail(CVXB_ELINSOLVE); }
  cleanup();
#undef QR_OK
#undef QR_TRY
  *out = S;
  return CVXB_OK;
}

// u = F'(x - z0)      (SolutionSpace.parameter, SolutionSpace.scala:32)
int solution_space_parameter(Handle& h, SolutionSpaceDev* S, const double* x, double* u) {
  const int n = S->n, k = n - S->p;
  sub_vec_kernel<<<(n + 255) / 256, 256, 0, h.stream>>>(n, x, S->z0, S->tmp);
  h.launches++;
  CVXB_CUDA_OK(cudaGetLastError());
  return gemv_t(h, n, k, 1.0, S->F(), S->ldq, S->tmp, 0.0, u);
}

// x = z0 + F u
int solution_space_map(Handle& h, SolutionSpaceDev* S, const double* u, double* x) {
  const int n = S->n, k = n - S->p;
  CVXB_TRY(gemv_n(h, n, k, 1.0, S->F(), S->ldq, u, 0.0, x));
  add_vec_kernel<<<(n + 255) / 256, 256, 0, h.stream>>>(n, S->z0, x);
  h.launches++;
  CVXB_CUDA_OK(cudaGetLastError());
  return CVXB_OK;
}

}  // namespace cvxb

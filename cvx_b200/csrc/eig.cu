// Last-resort solves of the reference's fallback chain:
//   MatrixUtils.diagonalizationSolve   MatrixUtils.scala:603-707  (with defect D3: once the straightforward
//                                      solution misses the tolerance, the regularisation loop cannot succeed)
//   MatrixUtils.symSolve / svdSolve    MatrixUtils.scala:712-751  (Breeze eigSym / svd -> LAPACK dsyev / dgesdd)
//   KKTSystem.kktSymSolve              KKTSystem.scala:283-310    (symSolve on the (n+p)^2 KKT matrix)
// The decomposition is a one-sided (Hestenes) Jacobi SVD, A V = U S: column pairs of W = A V are rotated
// until mutually orthogonal, n/2 disjoint pairs per launch in round-robin order, V accumulates the
// rotations.  It serves both branches: for a symmetric matrix u_i = sign(lambda_i) v_i and
// s_i = |lambda_i|, so  V S^-1 U' b  is exactly the eigen-decomposition solve.  HBM-bound (each round
// streams W and V once); a rarely taken path, so simplicity and accuracy win over speed.
#include "kkt.cuh"
#include "vecops.cuh"

namespace cvxb {
namespace {

__global__ void identity_kernel(int n, double* V, int ldv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int j = blockIdx.y; j < n; j += gridDim.y) V[(size_t)j * ldv + i] = (i == j) ? 1.0 : 0.0;
}

// round-robin pair (i, j) number k of round r over np (even) players; player np-1 is fixed
__device__ __forceinline__ void rr_pair(int np, int r, int k, int* i, int* j) {
  const int m = np - 1;
  *i = (r + k) % m;
  *j = (k == 0) ? m : (r + m - k) % m;
}

__global__ void __launch_bounds__(256) jacobi_round_kernel(int n, int np, int round, double* __restrict__ W, int ldw,
                                                           double* __restrict__ V, int ldv, double thresh,
                                                           unsigned* rotations) {
  __shared__ double buf[33];
  int ci, cj;
  rr_pair(np, round, blockIdx.x, &ci, &cj);
  if (ci >= n || cj >= n) return;            // dummy column of an odd-sized problem
  double* wi = W + (size_t)ci * ldw;
  double* wj = W + (size_t)cj * ldw;
  double a = 0.0, b = 0.0, g = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) {
    const double x = wi[k], y = wj[k];
    a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
  }
  a = block_sum(a, buf);
  b = block_sum(b, buf);
  g = block_sum(g, buf);
  if (!(fabs(g) > thresh * sqrt(a * b))) return;      // already orthogonal (also when a column is zero)
  const double zeta = (b - a) / (2.0 * g);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  for (int k = threadIdx.x; k < n; k += 256) {
    const double x = wi[k], y = wj[k];
    wi[k] = c * x - s * y;
    wj[k] = s * x + c * y;
  }
  double* vi = V + (size_t)ci * ldv;
  double* vj = V + (size_t)cj * ldv;
  for (int k = threadIdx.x; k < n; k += 256) {
    const double x = vi[k], y = vj[k];
    vi[k] = c * x - s * y;
    vj[k] = s * x + c * y;
  }
  if (threadIdx.x == 0) atomicAdd(rotations, 1u);
}

// c_i = (w_i . b) / s_i  (0 when s_i == 0), and the coefficient vectors of the two products
//   a_over_s_i = c_i / s_i  (for  b0 = U a = W (a / s)  and  z = c / s)
__global__ void __launch_bounds__(256) svd_coeff_kernel(int n, const double* __restrict__ W, int ldw,
                                                        const double* __restrict__ b, double sign,
                                                        double* __restrict__ z, double* __restrict__ sig) {
  __shared__ double buf[33];
  const int j = blockIdx.x;
  const double* w = W + (size_t)j * ldw;
  double s2 = 0.0, wb = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) {
    s2 = fma(w[k], w[k], s2);
    wb = fma(w[k], sign * b[k], wb);
  }
  s2 = block_sum(s2, buf);
  wb = block_sum(wb, buf);
  if (threadIdx.x == 0) {
    const double s = sqrt(s2);
    sig[j] = s;
    // (u_j . b) / d_j with u_j = w_j / s  ->  (w_j . b) / s^2 ; exact-zero test as in the reference
    z[j] = (s > 0.0) ? (wb / s) / s : 0.0;
  }
}

// Symmetric matrices (symSolve / kktSymSolve, MatrixUtils.scala:734-751: eigSym, U = V = eigenvectors, d = eigenvalues): the
// coefficients are taken from V, which is orthonormal to rounding by construction (a product of rotations), and from the
// Rayleigh quotient lambda_j = v_j . (A v_j) = v_j . w_j.  The left vectors u_j = w_j / s_j of a one-sided Jacobi SVD are
// only orthonormal for singular values well above rounding: for a (numerically) singular matrix -- dependent equality rows
// make the KKT matrix singular -- u_null is a unit vector of pure noise, not orthogonal to the range, and the projection
// U U'b of a perfectly consistent right-hand side missed it by percents (found by tools/gpu_fuzz_kkt.py).
//   c_j = v_j . b ;  a_j = c_j (lambda_j != 0, exact-zero test as in the reference) ;  z_j = c_j / lambda_j (sym_finish_kernel)
__global__ void __launch_bounds__(256) sym_coeff_kernel(int n, const double* __restrict__ W, int ldw, const double* __restrict__ V,
                                                        int ldv, const double* __restrict__ b, double sign,
                                                        double* __restrict__ z, double* __restrict__ a) {
  __shared__ double buf[33];
  const int j = blockIdx.x;
  const double* w = W + (size_t)j * ldw;
  const double* v = V + (size_t)j * ldv;
  double lam = 0.0, vb = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) {
    lam = fma(v[k], w[k], lam);
    vb = fma(v[k], sign * b[k], vb);
  }
  lam = block_sum(lam, buf);
  vb = block_sum(vb, buf);
  if (threadIdx.x == 0) {
    z[j] = vb;        // c_j, finished by sym_finish_kernel
    a[j] = lam;
  }
}

// LAPACK's dsyev returns the eigenvalues of a singular matrix as noise of size eps * ||A||, and the reference divides by
// them: the component c_j / lambda_j v_j it adds lies in the null space and is harmless.  The Rayleigh quotient of a Jacobi
// null column is noise too, but of ANY size down to 1e-150 (a column that small keeps shrinking under the rotations), and
// c_j / lambda_j then overflows the solution.  Eigenvalues at the rounding level of the largest one are therefore solved
// as the zeros they stand for (z_j = 0: the minimum-norm solution; the x-part of a KKT solution with dependent equality
// rows is the same); they stay in the projection of the range test as in the reference unless exactly zero.
__global__ void __launch_bounds__(1024) sym_finish_kernel(int n, double* __restrict__ z, double* __restrict__ a) {
  __shared__ double buf[33];
  double mx = 0.0;
  for (int j = threadIdx.x; j < n; j += 1024) mx = fmax(mx, fabs(a[j]));
  mx = -block_min(-mx, buf);
  const double thr = 64.0 * 2.220446049250313e-16 * (double)n * mx;
  for (int j = threadIdx.x; j < n; j += 1024) {
    const double lam = a[j], c = z[j];
    z[j] = (fabs(lam) > thr) ? c / lam : 0.0;
    a[j] = (lam != 0.0) ? c : 0.0;
  }
}

// General matrices (svdSolve): singular values at the rounding level of the largest one are treated as the zeros they
// stand for (their left vectors w_j / s_j are noise, see above; LAPACK's U is orthonormal there, the reference then
// divides rounding noise by rounding noise)
__global__ void __launch_bounds__(1024) svd_threshold_kernel(int n, const double* __restrict__ sig, double* __restrict__ z) {
  __shared__ double buf[33];
  double mx = 0.0;
  for (int j = threadIdx.x; j < n; j += 1024) mx = fmax(mx, sig[j]);
  mx = -block_min(-mx, buf);
  const double thr = 64.0 * 2.220446049250313e-16 * (double)n * mx;
  for (int j = threadIdx.x; j < n; j += 1024)
    if (!(sig[j] > thr)) z[j] = 0.0;
}

// relative sizes  ||b - b0|| / relsize(b)  and  ||A w - b|| / relsize(b)
__global__ void __launch_bounds__(VT) svd_check_kernel(int n, double sign, const double* __restrict__ b,
                                                       const double* __restrict__ b0, const double* __restrict__ aw,
                                                       double tol, double* scal) {
  __shared__ double buf[33];
  double nb = 0.0, d0 = 0.0, d1 = 0.0;
  for (int i = threadIdx.x; i < n; i += VT) {
    const double bi = sign * b[i];
    nb = fma(bi, bi, nb);
    double r0 = bi - b0[i], r1 = aw[i] - bi;
    d0 = fma(r0, r0, d0);
    d1 = fma(r1, r1, d1);
  }
  nb = block_sum(nb, buf);
  d0 = block_sum(d0, buf);
  d1 = block_sum(d1, buf);
  if (threadIdx.x == 0) {
    scal[S_TMP0] = relative_size(sqrt(d0), sqrt(nb), tol);
    scal[S_TMP1] = relative_size(sqrt(d1), sqrt(nb), tol);
  }
}

// M = [H A'; A 0]  ((n+p)^2),  rhs = [-q; b]      KKTSystem.scala:253-260, 283-290
__global__ void kkt_matrix_kernel(int n, int p, const double* __restrict__ Hm, int ldh, const double* __restrict__ A, int lda,
                                  double* __restrict__ M, int ldm) {
  const int N = n + p;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int j = blockIdx.y; j < N; j += gridDim.y) {
    double v;
    if (i < n && j < n) v = Hm[(size_t)j * ldh + i];
    else if (i >= n && j < n) v = A[(size_t)j * lda + (i - n)];
    else if (i < n && j >= n) v = A[(size_t)i * lda + (j - n)];
    else v = 0.0;
    M[(size_t)j * ldm + i] = v;
  }
}
__global__ void kkt_rhs2_kernel(int n, int p, const double* __restrict__ q, const double* __restrict__ b, double* rhs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rhs[i] = -q[i];
  else if (i < n + p) rhs[i] = b[i - n];
}

}  // namespace

// x = pseudo-inverse solve of A x = sign*b through the Jacobi SVD, with the reference's two acceptance
// tests.  Returns CVXB_EUNSOLVABLE (UnsolvableSystemException) when either fails.
int svd_solve_device(Handle& h, int n, const double* A, int lda, const double* b, double sign, double tol, double* x,
                     int* sweeps_out, bool symmetric) {
  const int ld = pad_ld(n);
  const int np = n + (n & 1);
  double *W = nullptr, *V = nullptr, *vec = nullptr;
  unsigned* rot = nullptr;
  CVXB_CUDA_OK(cudaMalloc((void**)&W, (size_t)ld * n * sizeof(double)));
  CVXB_CUDA_OK(cudaMalloc((void**)&V, (size_t)ld * n * sizeof(double)));
  CVXB_CUDA_OK(cudaMalloc((void**)&vec, (size_t)ld * 4 * sizeof(double)));
  CVXB_CUDA_OK(cudaMalloc((void**)&rot, sizeof(unsigned)));
  auto cleanup = [&]() { cudaFree(W); cudaFree(V); cudaFree(vec); cudaFree(rot); };
  int st = copy_matrix(h, n, n, A, lda, W, ld);
  if (st == CVXB_OK) {
    identity_kernel<<<dim3((n + 127) / 128, n > 1024 ? 1024 : n), 128, 0, h.stream>>>(n, V, ld);
    h.launches++;
  }
  int sweeps = 0;
  unsigned hrot = 1;
  while (st == CVXB_OK && hrot != 0 && sweeps < 60 && n > 1) {
    if (cudaMemsetAsync(rot, 0, sizeof(unsigned), h.stream) != cudaSuccess) { st = CVXB_ECUDA; break; }
    for (int r = 0; r < np - 1; ++r) {
      jacobi_round_kernel<<<np / 2, 256, 0, h.stream>>>(n, np, r, W, ld, V, ld, 1e-15, rot);
      h.launches++;
    }
    if (cudaMemcpyAsync(&hrot, rot, sizeof(unsigned), cudaMemcpyDeviceToHost, h.stream) != cudaSuccess ||
        cudaStreamSynchronize(h.stream) != cudaSuccess) { st = CVXB_ECUDA; break; }
    ++sweeps;
  }
  if (sweeps_out) *sweeps_out = sweeps;
  if (st != CVXB_OK) { cleanup(); if (st == CVXB_ECUDA) set_last_error("svd_solve_device: CUDA failure"); return st; }
  double *z = vec, *sig = vec + ld, *b0 = vec + 2 * ld, *aw = vec + 3 * ld;
  if (symmetric) {
    sym_coeff_kernel<<<n, 256, 0, h.stream>>>(n, W, ld, V, ld, b, sign, z, sig);      // sig := a (projection coefficients)
    sym_finish_kernel<<<1, 1024, 0, h.stream>>>(n, z, sig);
    h.launches += 2;
    st = gemv_n(h, n, n, 1.0, V, ld, sig, 0.0, b0);        // U a with U = V
  } else {
    svd_coeff_kernel<<<n, 256, 0, h.stream>>>(n, W, ld, b, sign, z, sig);
    svd_threshold_kernel<<<1, 1024, 0, h.stream>>>(n, sig, z);
    h.launches += 2;
    st = gemv_n(h, n, n, 1.0, W, ld, z, 0.0, b0);          // U a = W (c / s)
  }
  if (st == CVXB_OK) st = gemv_n(h, n, n, 1.0, V, ld, z, 0.0, x);       // w = V z
  if (st == CVXB_OK) st = gemv_n(h, n, n, 1.0, A, lda, x, 0.0, aw);     // A w
  if (st == CVXB_OK) {
    svd_check_kernel<<<1, VT, 0, h.stream>>>(n, sign, b, b0, aw, tol, h.d_scal);
    h.launches++;
    st = fetch_status(h);
  }
  cleanup();
  if (st != CVXB_OK) return st;
  const double relDist = h.h_scal[S_TMP0], relErr = h.h_scal[S_TMP1];
  if (!(relDist <= tol)) {
    set_last_error("diagonalizationSolve: min_x||Ax-b||/||b|| = %.5g > tol = %g (UnsolvableSystemException)", relDist, tol);
    return CVXB_EUNSOLVABLE;
  }
  if (!(relErr <= tol)) {
    set_last_error("diagonalizationSolve: system not solvable within tolerance tol = %g, error ||Ax-b||/||b|| = %.5g "
                   "(UnsolvableSystemException; the reference's regularisation loop cannot recover, defect D3)", tol, relErr);
    return CVXB_EUNSOLVABLE;
  }
  return CVXB_OK;
}

// KKTSystem.kktSymSolve: eigen/SVD solve of the full KKT matrix with right-hand side (-q, b)
int kkt_sym_solve_device(Handle& h, int n, int p, const double* Hm, int ldh, const double* A, int lda, const double* q,
                         const double* b, double tol, double* x, double* w) {
  const int N = n + p, ld = pad_ld(N);
  double *M = nullptr, *rhs = nullptr, *sol = nullptr;
  CVXB_CUDA_OK(cudaMalloc((void**)&M, (size_t)ld * N * sizeof(double)));
  CVXB_CUDA_OK(cudaMalloc((void**)&rhs, (size_t)ld * 2 * sizeof(double)));
  sol = rhs + ld;
  kkt_matrix_kernel<<<dim3((N + 127) / 128, N > 1024 ? 1024 : N), 128, 0, h.stream>>>(n, p, Hm, ldh, A, lda, M, ld);
  kkt_rhs2_kernel<<<(N + 255) / 256, 256, 0, h.stream>>>(n, p, q, b, rhs);
  h.launches += 2;
  int st = svd_solve_device(h, N, M, ld, rhs, 1.0, tol, sol, nullptr, true);
  if (st == CVXB_OK) {
    if (cudaMemcpyAsync(x, sol, n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream) != cudaSuccess ||
        cudaMemcpyAsync(w, sol + n, p * sizeof(double), cudaMemcpyDeviceToDevice, h.stream) != cudaSuccess ||
        cudaStreamSynchronize(h.stream) != cudaSuccess)
      st = CVXB_ECUDA;
  }
  cudaFree(M);
  cudaFree(rhs);
  return st;
}

}  // namespace cvxb

// Device-resident KKT / Cholesky solves (KKTSystem.scala, MatrixUtils.choleskySolve).
#pragma once
#include "common.cuh"

namespace cvxb {

struct KktWork {
  int n = 0, p = 0, ldn = 0, ldp = 0;
  double* L = nullptr;      // n x n: lower((dd')oH) -> Cholesky factor (zeros above the diagonal)
  double* invD = nullptr;   // inverse diagonal blocks of L
  double* Y = nullptr;      // n x (p+1): [d o A', d o q] -> L^-1 [..]
  double* S = nullptr;      // p x p Schur complement -> its factor
  double* invDs = nullptr;
  double* Xlit = nullptr;   // n x (p+1): H^-1 [DA', Dq]   } the reference's literal block elimination (bugCompat & 2),
  double* Blit = nullptr;   // n x (p+1): copy of [DA', Dq]  } allocated with the rest when p > 0
  double* Hk = nullptr;     // path 1: H + A'A (allocated on first use)
  double* Q = nullptr;      // symmetric-solve scratch (allocated on first use)
  double *dr = nullptr, *colsq = nullptr, *qs = nullptr, *xs = nullptr, *t1 = nullptr, *t2 = nullptr, *t3 = nullptr,
         *qk = nullptr, *dr2 = nullptr;
  double *z = nullptr, *ax = nullptr, *tp = nullptr;
  std::vector<void*> owned;
  Arena* arena = nullptr;
};

// SolutionSpace of A x = b (SolutionSpace.scala:20-33), device resident (qr.cu): x = z0 + F u
struct SolutionSpaceDev {
  int n = 0, p = 0, ldq = 0;
  int device = 0;           // the creating handle's device and stream (destroy synchronises that stream only)
  cudaStream_t stream = nullptr;
  double* Q = nullptr;      // n x n orthogonal factor of A' = QR; F = Q(:, p..n-1)
  double* z0 = nullptr;     // minimum-norm solution
  double* tmp = nullptr;    // 2 * pad_ld(n) scratch
  double* Fext = nullptr;   // basis handed in by the caller (cvxb_solution_space_from_basis): n x k, leading dimension ldq; Q unused
  double* F() const { return Fext ? Fext : Q + (size_t)p * ldq; }
  int k() const { return n - p; }
};
int solution_space_build(Handle& h, int p, int n, const double* A, int lda, const double* b, SolutionSpaceDev** out);
void solution_space_free(SolutionSpaceDev* S);
int solution_space_parameter(Handle& h, SolutionSpaceDev* S, const double* x, double* u);   // u = F'(x - z0)
int solution_space_map(Handle& h, SolutionSpaceDev* S, const double* u, double* x);         // x = z0 + F u

int kkt_work_alloc(Handle& h, KktWork& W, int n, int p, Arena* arena = nullptr);
size_t kkt_work_bytes(int n, int p);
void kkt_work_free(KktWork& W);

// Enqueue one solvePD (KKTSystem.scala:200-246) without host synchronisation:
//   Ruiz(H) -> regularizedCholesky attempt (delta = 0 or cholRegDelta) -> Schur block solve ->
//   residual check -> x = d o y.  Outcome in d_flag[F_CHOL_H, F_CHOL_S, F_RUIZ_SWEEPS, F_BAD] and
//   d_scal[S_MINDIAG_H, S_ERR1, S_ERR2].  F_BAD != 0 <=> this attempt must not be used.
int kkt_enqueue(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* A, int lda,
                const double* q, const double* b, double tol, bool regularize, bool skip_ruiz, double* x, double* w,
                bool prefactored = false);   // prefactored: Hm holds a Cholesky factor L (KKTSystem.solveWithCholFactor)

// KKTSystem.solve (KKTSystem.scala:43-66) on device pointers, with the host reading the status
// words between attempts: path 0 -> (regularised retry) -> path 1 (H + A'A) -> path 2.
int kkt_solve_device(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* A, int lda,
                     const double* q, const double* b, double tol, double* x, double* w, cvxb_kkt_info* info);
// called after kkt_enqueue(regularize=false) + sync flagged F_BAD: continue the chain from there
int kkt_solve_fallbacks(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* A,
                        int lda, const double* q, const double* b, double tol, double* x, double* w,
                        cvxb_kkt_info* info);

// MatrixUtils.choleskySolve (MatrixUtils.scala:468-516): one attempt, no host sync.  Solves H x = rhs_sign * b.
int chol_enqueue(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* b,
                 double rhs_sign, double tol, bool regularize, bool skip_ruiz, double* x);
// choleskySolve with its regularisation retry decided on the host; LinSolveException -> CVXB_ELINSOLVE
int chol_solve_device(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* b,
                      double rhs_sign, double tol, double* x, cvxb_kkt_info* info);
// continue after a flagged optimistic chol_enqueue(regularize=false)
int chol_solve_retry(Handle& h, KktWork& W, const cvxb_params& P, const double* Hm, int ldh, const double* b,
                     double rhs_sign, double tol, double* x, cvxb_kkt_info* info);

// MatrixUtils.symSolve / svdSolve (eig.cu): pseudo-inverse solve of A x = sign*b by one-sided Jacobi SVD with the
// reference's acceptance tests; CVXB_EUNSOLVABLE = UnsolvableSystemException
// symmetric: A is symmetric (symSolve / kktSymSolve: eigen-decomposition semantics, coefficients from V); else svdSolve
int svd_solve_device(Handle& h, int n, const double* A, int lda, const double* b, double sign, double tol, double* x,
                     int* sweeps_out, bool symmetric);
// KKTSystem.kktSymSolve on the (n+p)^2 KKT matrix
int kkt_sym_solve_device(Handle& h, int n, int p, const double* Hm, int ldh, const double* A, int lda, const double* q,
                         const double* b, double tol, double* x, double* w);

// copies flags + scalars to the pinned mirrors and waits for the stream
int fetch_status(Handle& h);
void fill_info(Handle& h, cvxb_kkt_info* info, int path, int regularized);

}  // namespace cvxb

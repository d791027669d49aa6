/* cvxb.h -- C ABI of libcvxb: the B200-native (sm_100a) Newton/KKT hot path behind the API of the
 * Scala convex solver spyqqqdia/cvx.  Plain C: pointers, sizes and POD structs only.
 *
 * The reference has no FFI of its own (pure Scala + Breeze/netlib JNI); each entry point below
 * names the reference method it stands in for (paths relative to src/main/scala/cvx/).  The JNI
 * binding a maintainer would add on the Scala side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every matrix is column-major FP64 with an explicit leading dimension, exactly Breeze's
 *     DenseMatrix layout (data, offset, majorStride);
 *   - pointers are HOST pointers unless the handle was created with CVXB_FLAG_DEVICE_PTRS, in which
 *     case all array arguments of the seam-B calls are device pointers on the handle's device;
 *   - every function returns a cvxb_status; cvxb_last_error() gives the message of the last failure
 *     on the calling thread;
 *   - a handle owns one CUDA device + stream and is not thread-safe; use one handle per thread.
 */
#ifndef CVXB_H
#define CVXB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cvxb_status {
  CVXB_OK = 0,
  CVXB_ELINSOLVE = 1,     /* LinSolveException            (LinSolveException.scala:11-17)           */
  CVXB_EUNSOLVABLE = 2,   /* UnsolvableSystemException    (UnsolvableSystemException.scala)          */
  CVXB_ELINESEARCH = 3,   /* LineSearchFailedException (PD) / NotConvergedException.Breakdown (BR)   */
  CVXB_ENOTFEASIBLE = 4,  /* IllegalArgumentException "x not strictly feasible" (BarrierSolver.scala:284) */
  CVXB_EINFEASIBLE = 5,   /* InfeasibleProblemException   (InfeasibleProblemException.scala:6)       */
  CVXB_EDIM = 6,          /* AssertionError: dimension mismatch (KKTSystem.scala:31-32)              */
  CVXB_ENOTSYMMETRIC = 7, /* Breeze MatrixNotSymmetricException raised inside cholesky / eigSym       */
  CVXB_ECUDA = 8,         /* CUDA runtime failure (no reference analogue)                            */
  CVXB_EINVAL = 9,        /* bad argument                                                            */
  CVXB_ENOTIMPL = 10      /* path of the reference not built yet (see DESIGN.md "out of scope")      */
} cvxb_status;

typedef struct cvxb_handle_s* cvxb_handle;
typedef struct cvxb_problem_s* cvxb_problem;
typedef struct cvxb_solution_space_s* cvxb_solution_space;

enum { CVXB_FLAG_DEVICE_PTRS = 1 };

/* ---- lifetime ------------------------------------------------------------------------------ */
/* `stream` is a cudaStream_t (or NULL: the handle creates its own non-blocking stream). */
int cvxb_create(int device, void* stream, unsigned flags, cvxb_handle* out);
int cvxb_destroy(cvxb_handle h);
int cvxb_synchronize(cvxb_handle h);
const char* cvxb_last_error(void);
const char* cvxb_version(void);
/* number of kernels launched through this handle since creation (bench.py "gpu_launches") */
long long cvxb_launch_count(cvxb_handle h);
/* number of host round trips so far: reads of the device status block followed by a stream synchronisation.  A barrier
 * stage driven from the device costs two (its initial evaluation and the end of its Newton loop), not one per step. */
long long cvxb_status_read_count(cvxb_handle h);

/* Per-launch timing of the dominant kernel (the Hessian-assembly SYRK) with CUDA events on the handle's
 * stream: enable, run solves, then read {launches, total milliseconds, total algorithmic flops}. */
int cvxb_profile_enable(cvxb_handle h, int on);
int cvxb_profile_read(cvxb_handle h, long long* launches, double* ms_total, double* flops_total);
/* Further timed ranges of the Newton step (same enable switch; plain launches only, i.e. the primal-dual loop and the
 * seam-B calls, not the captured barrier step): {count, total ms, total algorithmic work}.  range: 0 = Hessian SYRK
 * (same as cvxb_profile_read), 1 = Cholesky trailing updates A22 -= A21 A21' of the recursive levels (flops),
 * 2 = factorisation of H with the forward substitution of [DA', Dq] riding along (flops), 3 = Schur complement SYRK
 * (flops), 4 = ruizEquilibrate (bytes of ONE sweep over H per call), 5 = the GEMVs G x and G'(1/f) (bytes),
 * 6 = the chain-bound look-ahead phases of the factorisations (flops), 7 = the A21 L11^-T solves of the recursive
 * levels (flops). */
int cvxb_profile_read_range(cvxb_handle h, int range, long long* count, double* ms_total, double* work_total);

/* ---- parameters: SolverParams.scala:24-46 plus the constants hard-coded in the solvers ------- */
typedef struct cvxb_params {
  int maxIter;          /* SolverParams.maxIter          1000 */
  double alpha;         /* line-search descent factor    0.04 */
  double beta;          /* line-search backtrack factor  0.8  */
  double tolSolver;     /* 1e-8 */
  double tolEqSolve;    /* 1e-1 */
  double tolFeas;       /* 1e-7 */
  double delta;         /* 1e-6 (never used by the reference, KKTSystem.scala:43; kept) */
  double mu;            /* 10   BarrierSolver.scala:73,130; PrimalDualSolver.scala:392,562 */
  double t0;            /* 1    BarrierSolver.scala:74,131 */
  int ruizMaxSweeps;    /* 20   MatrixUtils.scala:247 */
  double ruizTol;       /* 1e-6 MatrixUtils.scala:250 */
  double cholRegDelta;  /* 1e-10 MatrixUtils.scala:454 */
  double cholMinDiag;   /* 1e-7  MatrixUtils.scala:460 */
  double newtonRegDelta;/* 1e-9  UnconstrainedSolver.scala:60 */
  double phase1EqTol;   /* 1e-6  ConstraintSet.scala:342, CvxUtils.scala:86 */
  double pdStepFraction;/* 0.99  PrimalDualSolver.scala:339,509 */
  int bugCompat;        /* bit 0: reproduce reference defects D1/D2 of PrimalDualSolver.solve_withEQs;
                         * bit 1: form the Schur complement literally as KKTSystem.scala:116-139 does (H^-1 A' by two
                         * triangular solves, A (H^-1 A') by a GEMM, symmetrised) instead of Y'Y with Y = L^-1 A': twice
                         * the flops, and the reference's loss of definiteness at cond(H) ~ 1e20 */
  long long stepLimit;  /* 0 (reference behaviour): no budget.  >0: a budget of Newton steps for the whole call (phase I
                         * included) -- the solve returns its current iterate with status OK when it is used up; not a
                         * SolverParams field of the reference (callers with a time budget, and bench.py --steps) */
} cvxb_params;
int cvxb_default_params(cvxb_params* p);

/* ---- seam B: per-step linear algebra on caller-owned matrices -------------------------------- */
typedef struct cvxb_kkt_info {
  int path;            /* 0 solvePD(H); 1 solvePD(H+A'A) (KKTSystem.scala:57-59); 2 kktSymSolve / symSolve
                          (decomposition fallback); 3 svdSolve (asymmetric SymmetricLinearSystem)        */
  int regularized;     /* regularizedCholesky took the Q+1e-10 I branch (MatrixUtils.scala:452-461) */
  int ruiz_sweeps;     /* sweeps taken by ruizEquilibrate (MatrixUtils.scala:240-268)                */
  int chol_info;       /* 0, or 1-based column of the first non-positive pivot of the last attempt   */
  double min_diag;     /* min diag(L) of the factor that was used                                    */
  double err1;         /* ||LL'x + A'w + q|| / (tol+||q||)   (KKTSystem.scala:148-151)               */
  double err2;         /* ||Ax-b|| / (tol+||b||)             (KKTSystem.scala:153-154)               */
} cvxb_kkt_info;

/* KKTSystem(H,A,q,b).solve(delta,logger,tol,debugLevel): (x,w)   KKTSystem.scala:43-66
 * Solves Hx + A'w = -q, Ax = b.  H is n x n symmetric, A is p x n (p may be 0). */
int cvxb_kkt_solve(cvxb_handle h, int n, int p, const double* H, int ldh, const double* A, int lda,
                   const double* q, const double* b, double tol, double* x, double* w,
                   cvxb_kkt_info* info);

/* KKTData(H,A,g,r,...).reduced + KKTSystem.solve + KKTData.paddVector (KKTData.scala:68-127; exercised by
 * KktTest.testKktSystemReduction, KktTest.scala:52-104): variables x_j on which neither H, A nor g depend (row and
 * column j of H, column j of A all zero -- as in phase-I systems with unconstrained variables) are eliminated, the
 * reduced system  Hr xr + Ar'w = -gr, Ar xr = r  is solved with the cvxb_kkt_solve chain, and x is padded back with
 * zeros.  A zero row with |g_j| >= 1e-15 gives CVXB_EUNSOLVABLE.  null_indices (n ints, may be NULL) receives the
 * eliminated indices in increasing order, *n_null their number.  Host pointers only. */
int cvxb_kkt_solve_reduced(cvxb_handle h, int n, int p, const double* H, int ldh, const double* A, int lda, const double* g,
                           const double* r, double tol, double* x, double* w, int* null_indices, int* n_null,
                           cvxb_kkt_info* info);

/* KKTSystem.solveWithCholFactor(L,A,q,b,logger,tol,debugLevel)   KKTSystem.scala:99-167 */
int cvxb_kkt_solve_with_chol_factor(cvxb_handle h, int n, int p, const double* L, int ldl,
                                    const double* A, int lda, const double* q, const double* b,
                                    double tol, double* x, double* w, cvxb_kkt_info* info);

/* MatrixUtils.choleskySolve(H,b,logger,tol,debugLevel)           MatrixUtils.scala:468-516 */
int cvxb_cholesky_solve(cvxb_handle h, int n, const double* H, int ldh, const double* b, double tol,
                        double* x, cvxb_kkt_info* info);

/* SymmetricLinearSystem(H,r,logger).solve(tol,debugLevel)        SymmetricLinearSystem.scala:15-56 */
int cvxb_symmetric_solve(cvxb_handle h, int n, const double* H, int ldh, const double* r, double tol,
                         double* x, cvxb_kkt_info* info);

/* ---- seam B with pinned, double-buffered staging: general (non-closed-form) objectives ---------------------------
 * An objective given only as closures (ObjectiveFunction.valueAt / gradientAt / hessianAt, ObjectiveFunction.scala:12-14;
 * e.g. the reference's Type1Function power problems, src/test/scala/cvx/Type1Function.scala:67-78) has its Hessian
 * assembled on the HOST every Newton step.  A stage owns the device copy of H (n x n) and A (p x n), two pinned host
 * buffers of block_cols columns each, and a copy stream: the caller fills buffer 0 with columns [0, c), pushes it, fills
 * buffer 1 with [c, 2c) while the first block is in flight, waits for buffer 0, ... -- assembly and upload overlap -- and
 * then solves on the device-resident matrix:
 *   cvxb_stage_cholesky_solve   MatrixUtils.choleskySolve(H, b)       (UnconstrainedSolver.scala:50-66)
 *   cvxb_stage_kkt_solve        KKTSystem(H, A, q, b).solve           (EqualityConstrainedSolver.scala:52-58)
 * Vectors are plain host pointers.  Host pointers only (no CVXB_FLAG_DEVICE_PTRS). */
typedef struct cvxb_stage_s* cvxb_stage;
int cvxb_stage_create(cvxb_handle h, int n, int p, int block_cols /* 0 = default */, cvxb_stage* out);
int cvxb_stage_destroy(cvxb_stage stage);
/* pinned host buffer `which` (0 or 1): n x block_cols column-major with leading dimension n */
int cvxb_stage_buffer(cvxb_stage stage, int which, double** host_buffer, int* block_cols);
/* start the upload of buffer `which` into columns [col0, col0 + ncols) of H; returns at once */
int cvxb_stage_push(cvxb_stage stage, int which, int col0, int ncols);
/* block until buffer `which` may be overwritten (its last push has left the host) */
int cvxb_stage_wait(cvxb_stage stage, int which);
/* A (p x n, leading dimension lda): uploaded once, it does not change between Newton steps */
int cvxb_stage_set_equalities(cvxb_stage stage, const double* A, int lda);
int cvxb_stage_cholesky_solve(cvxb_stage stage, const double* b, double tol, double* x, cvxb_kkt_info* info);
int cvxb_stage_kkt_solve(cvxb_stage stage, const double* q, const double* b, double tol, double* x, double* w,
                         cvxb_kkt_info* info);

/* MatrixUtils.ruizEquilibrate(H): (d, Q)                         MatrixUtils.scala:240-268 */
int cvxb_ruiz_equilibrate(cvxb_handle h, int n, const double* H, int ldh, double* d, double* Q,
                          int ldq, int* sweeps);

/* MatrixUtils.regularizedCholesky(Q): L (lower, zeros above)     MatrixUtils.scala:452-461 */
int cvxb_regularized_cholesky(cvxb_handle h, int n, const double* Q, int ldq, double* L, int ldl,
                              cvxb_kkt_info* info);

/* MatrixUtils.triangularSolve(A,"L"|"U",B)                       MatrixUtils.scala:362-376
 * uplo 'L': solves L X = B; 'U': solves U X = B with U given as an upper-triangular matrix.
 * B (n x nrhs) is overwritten by X. */
int cvxb_triangular_solve(cvxb_handle h, char uplo, int n, int nrhs, const double* T, int ldt,
                          double* B, int ldb);

/* ---- seam A: device-resident problems (closed-form families) ------------- */
typedef enum cvxb_objective_kind {
  CVXB_OBJ_LINEAR = 0,    /* r + a'x            LinearObjectiveFunction.scala:5-22     */
  CVXB_OBJ_QUADRATIC = 1, /* r + a'x + x'Px/2   QuadraticObjectiveFunction.scala:11-33 */
  CVXB_OBJ_KL = 2,        /* sum x log(n x)     Dist_KL.scala:223-239                  */
  CVXB_OBJ_PNORM = 4,     /* sum |x_j|^p, p >= 2  ObjectiveFunctions.p_norm_p, ObjectiveFunctions.scala:70-83 (obj_pow = p) */
  CVXB_OBJ_KLDUAL = 3     /* w'z + R'exp(-B'z): the convex dual objective -L_*(z) of Dist_KL (Dist_KL.scala:143-163,
                             Duality.scala:68-75); obj_a = w (n), obj_P = B (n x obj_k, ld obj_ldP), obj_R = R (obj_k).
                             Barrier solver only.                                              */
} cvxb_objective_kind;

typedef struct cvxb_problem_desc {
  int n, m, p;               /* variables, linear inequalities, equalities                        */
  int objective;             /* cvxb_objective_kind                                               */
  const double* obj_a;       /* n   (LINEAR, QUADRATIC)                                           */
  double obj_r;
  const double* obj_P;       /* n x n, ld obj_ldP (QUADRATIC)                                     */
  int obj_ldP;
  const double* G;           /* m x n: rows are LinearConstraint.a  (LinearConstraint.scala:7-13) */
  int ldg;
  const double* g_r;         /* m: LinearConstraint.r (NULL = zeros)                              */
  const double* ub;          /* m: Constraint.ub                                                  */
  const double* A;           /* p x n  EqualityConstraint.A (NULL when p == 0)                    */
  int lda;
  const double* b;           /* p      EqualityConstraint.b                                       */
  const double* x_feasible;  /* n: ConstraintSet with FeasiblePoint .feasiblePoint, or NULL        */
  const double* x_defined;   /* n: ConstraintSet.pointWhereDefined (start of phase I)             */
  /* quadratic constraints  q_r[k] + q_a[:,k]'x + x'P_k x / 2 <= q_ub[k]   (QuadraticConstraint.scala:7-40);
   * they follow the m linear constraints in every per-constraint vector (lambda, stage order).           */
  int mq;                    /* number of quadratic constraints (0 = none)                        */
  const double* q_P;         /* mq symmetric n x n matrices, each column-major with ld n, packed   */
  const double* q_a;         /* n x mq, column k = a_k                                            */
  const double* q_r;         /* mq                                                                */
  const double* q_ub;        /* mq                                                                */
  /* CVXB_OBJ_KLDUAL only */
  int obj_k;                 /* columns of B = dimension of the primal KL problem                 */
  const double* obj_R;       /* obj_k                                                             */
  double obj_pow;            /* CVXB_OBJ_PNORM only: the exponent p >= 2                          */
} cvxb_problem_desc;

/* Duality.primalOptimum for CVXB_OBJ_KLDUAL: x = R o exp(-B'z) at the problem's current iterate (after a solve). */
int cvxb_kldual_primal_optimum(cvxb_handle h, cvxb_problem prob, double* x_primal);

/* ---- equality elimination x = z0 + F u   -------------------------------------------------------
 * MatrixUtils.solveUnderdetermined (MatrixUtils.scala:536-550) / SolutionSpace (SolutionSpace.scala:20-33), which the
 * reference runs eagerly in every EqualityConstraint constructor (EqualityConstraint.scala:21-23): QR of A' (n x p)
 * by blocked Householder reflections on the device, F = Q(:, p..n-1) an orthonormal basis of ker A, z0 = the
 * minimum-norm solution of A x = b.  A is p x n column-major with 1 <= p < n and full rank (not checked, as in the
 * reference; a zero pivot of R gives CVXB_ELINSOLVE).  Host pointers unless the handle has CVXB_FLAG_DEVICE_PTRS. */
int cvxb_solution_space_create(cvxb_handle h, int p, int n, const double* A, int lda, const double* b,
                               cvxb_solution_space* out);
/* The same object from a caller-supplied affine map x = z0 + F u (Solver.affineTransformed(z0, F, u0), Solver.scala:33-46;
 * BarrierSolver.affineTransformed, BarrierSolver.scala:209-247): F is n x k column-major with leading dimension ldf.
 * Nothing is assumed about F (the reference does not check orthonormality either); cvxb_solution_space_parameter
 * computes F'(x - z0), which inverts the map only for an orthonormal F. */
int cvxb_solution_space_from_basis(cvxb_handle h, int n, int k, const double* z0, const double* F, int ldf,
                                   cvxb_solution_space* out);
int cvxb_solution_space_destroy(cvxb_solution_space space);
/* z0 (n) and F (n x (n-p), leading dimension ldf); either may be NULL */
int cvxb_solution_space_get(cvxb_handle h, cvxb_solution_space space, double* z0, double* F, int ldf);
/* SolutionSpace.parameter: u = F'(x - z0) (length n-p) */
int cvxb_solution_space_parameter(cvxb_handle h, cvxb_solution_space space, const double* x, double* u);
/* x = z0 + F u */
int cvxb_solution_space_map(cvxb_handle h, cvxb_solution_space space, const double* u, double* x);
/* one shot: (z0, F) = MatrixUtils.solveUnderdetermined(A, b) */
int cvxb_solve_underdetermined(cvxb_handle h, int p, int n, const double* A, int lda, const double* b, double* z0,
                               double* F, int ldf);
/* BarrierSolver.reduced(sol) / PrimalDualSolver.reduced(sol) (BarrierSolver.scala:249-256, PrimalDualSolver.scala:699):
 * the problem of dimension n-p in the variable u, built on the device from a problem WITHOUT equality constraints
 * (linear or quadratic objective; linear and quadratic constraints): constraints (G F) u <= ub - r - G z0, P -> F'PF,
 * a -> F'(a + P z0); starting points u0 = F'(x0 - z0), with the reference's check ||x0 - (z0 + F u0)|| < tolEqSolve
 * (CVXB_EINVAL).  Solve it with cvxb_barrier_solve / cvxb_pd_solve; the solution's x is u, as in the reference --
 * map it back with cvxb_solution_space_map.  Other objective families: CVXB_ENOTIMPL. */
int cvxb_problem_reduce(cvxb_handle h, cvxb_problem prob, cvxb_solution_space space, const cvxb_params* pars,
                        cvxb_problem* reduced);

/* ConstraintSet.valuesAt-style evaluation: g[i] = g_i(x) for the m linear rows then the mq quadratic constraints
 * (Constraint.valueAt, LinearConstraint.scala:22, QuadraticConstraint.scala:30), and *strictly_satisfied =
 * ConstraintSet.isSatisfiedStrictlyBy(x) (ConstraintSet.scala:28-29: g_i(x)(1+3e-16) < ub_i for all i).  Host pointers.
 * Used by the phase-I drivers (starting values of the SOI variables, violated-constraint reports). */
int cvxb_constraint_values(cvxb_handle h, cvxb_problem prob, const double* x, double* g, int* strictly_satisfied);

/* mirrors Solution.scala:32-43; has_* say which Option fields are Some(...) */
typedef struct cvxb_solution {
  double* x;       /* n, caller-allocated */
  double* lambda;  /* m, caller-allocated or NULL */
  double* nu;      /* p, caller-allocated or NULL */
  int has_lambda, has_nu;
  double newtonDecrement; int has_newtonDecrement;
  double dualityGap;      int has_dualityGap;
  double equalityGap;     int has_equalityGap;
  double normGrad;        int has_normGrad;
  double normDualResidual;int has_normDualResidual;
  int iter;
  int maxedOut;
  /* extras */
  double objective;            /* objF.valueAt(x) at the returned point                   */
  int outer_stages;            /* barrier stages / PD iterations taken                    */
  long long newton_steps;      /* Newton iterations over all stages as the reference counts them (phase I not included) */
  long long executed_newton_steps; /* of those, the ones computed on the device: a stage that can no longer move
                                  (newton decrement <= tol but ||b-Ax|| > tol) makes the reference repeat the same
                                  step until maxIter; those repeats are counted above and skipped here */
  long long phase1_newton_steps;
  long long phase1_executed_steps;
  int phase1_stages;
  double phase1_s;             /* slack s at the phase-I solution (ConstraintSet.scala:369-371) */
  long long linesearch_trials; /* backtracking multiplications by beta, summed            */
  int kkt_fallbacks;           /* Newton steps that left path 0                           */
  int kkt_regularized;         /* Newton steps whose Cholesky was regularised             */
  int stage_newton_steps[128]; /* per outer stage                                         */
  double solve_ms;             /* device time of the solve, CUDA events on the handle's stream */
} cvxb_solution;

int cvxb_problem_create(cvxb_handle h, const cvxb_problem_desc* desc, cvxb_problem* out);
int cvxb_problem_destroy(cvxb_problem prob);

/* ConstraintSet.withFeasiblePoint(eqs,pars,debugLevel): phase I   ConstraintSet.scala:556-575.
 * On success the problem holds a strictly feasible point (also copied to x_feas if non-NULL). */
int cvxb_phase1(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, double* x_feas,
                cvxb_solution* phase1_out);

/* BarrierSolver(...).solve(debugLevel): Solution                   BarrierSolver.scala:184-188
 * Runs phase I first when the problem has no feasible point (OptimizationProblem.scala:174-196). */
int cvxb_barrier_solve(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, cvxb_solution* out);

/* PrimalDualSolver(...).solve(debugLevel): Solution                PrimalDualSolver.scala:628-641 */
int cvxb_pd_solve(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars, cvxb_solution* out);

/* One barrier Newton direction at (x, t): H = hessianBarrierFunction(t,x), g = gradientBarrierFunction
 * (BarrierSolver.scala:291-315), then KKTSystem(H,A,g,b-Ax).solve (EqualityConstrainedSolver.scala:52-58)
 * or choleskySolve(H,-g) when p == 0 (UnconstrainedSolver.scala:50-55).  Any of H_out (n x n, ld n),
 * g_out, dx, nu may be NULL. */
int cvxb_barrier_newton_direction(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars,
                                  const double* x, double t, double* H_out, double* g_out,
                                  double* dx, double* nu, cvxb_kkt_info* info);

/* One primal-dual search direction at (x, lambda, nu, t): H_pd (PrimalDualSolver.scala:216-240),
 * rhs1 (:162-176), KKT solve (:254-285), deltaLambda (:184-209). */
int cvxb_pd_newton_direction(cvxb_handle h, cvxb_problem prob, const cvxb_params* pars,
                             const double* x, const double* lambda, const double* nu, double t,
                             double* H_out, double* dx, double* dlambda, double* dnu,
                             cvxb_kkt_info* info);

/* ---- batched small problems (one CTA per problem) ------------------------------ */
/* B independent problems of identical shape (n <= 64, m <= 128, p in {0,1}); arrays are packed
 * problem after problem, each matrix column-major with ld = its row count.  objective[i] is a
 * cvxb_objective_kind; x0 must be strictly feasible unless phase1[i] is set.  Outputs: x (B*n), per-problem
 * status, Newton step and outer-stage counts.
 * phase1[i] != 0: x0 of problem i is only a point where the problem is defined (ConstraintSet.pointWhereDefined); the CTA
 * first runs the reference's phase-I analysis (ConstraintSet.phase_I_Analysis, ConstraintSet.scala:326-395, 556-575:
 * minimise s subject to g_i(x) - s <= ub_i, +-(a.x - b) - s <= phase1EqTol, from (x0, 1 + max(g(x0) - ub)), until the
 * objective is negative), fails the problem with CVXB_EINFEASIBLE unless s < tolSolver, and then solves from the feasible
 * point found.  The feasibility problem must fit the kernel too: n + 1 <= 64 and m + 2p <= 128. */
typedef struct cvxb_batch_desc {
  int B, n, m, p;
  const int* objective;    /* B */
  const int* pcount;       /* B or NULL: equalities of each problem (0 or 1 <= p); NULL = p for all */
  const double* obj_a;     /* B*n (ignored for KL problems) */
  const double* obj_r;     /* B */
  const double* obj_P;     /* B*n*n (ignored unless QUADRATIC) */
  const double* G;         /* B*m*n */
  const double* ub;        /* B*m   */
  const double* A;         /* B*p*n */
  const double* b;         /* B*p   */
  const double* x0;        /* B*n   */
  const int* phase1;       /* B or NULL (= no problem needs phase I) */
} cvxb_batch_desc;

enum { CVXB_BATCH_STAGES = 16 };
typedef struct cvxb_batch_result {
  double* x;               /* B*n */
  int* status;             /* B: cvxb_status per problem */
  int* newton_steps;       /* B */
  int* outer_stages;       /* B */
  double* objective;       /* B */
  double* duality_gap;     /* B */
  double* equality_gap;    /* B */
  double solve_ms;
  int* stage_newton_steps; /* B * CVXB_BATCH_STAGES or NULL: Newton steps of each of the first 16 outer stages (0 beyond
                              the last stage), as cvxb_solution.stage_newton_steps */
  long long* cycles;       /* B or NULL: SM clock cycles each problem occupied its CTA (load balance / roofline evidence) */
  int* phase1_newton_steps; /* B or NULL: Newton steps / outer stages of the phase-I analysis (0 where none ran) */
  int* phase1_stages;
  double* phase1_s;        /* B or NULL: final phase-I slack s (negative: strictly feasible point found) */
} cvxb_batch_result;

typedef struct cvxb_batch_s* cvxb_batch;
int cvxb_batch_create(cvxb_handle h, const cvxb_batch_desc* desc, cvxb_batch* out);
int cvxb_batch_destroy(cvxb_batch batch);
/* Any array pointer of `out` may be NULL: that result is then not copied to the caller (it stays on the device). */
int cvxb_batch_barrier_solve(cvxb_handle h, cvxb_batch batch, const cvxb_params* pars,
                             cvxb_batch_result* out);
/* The results of the last cvxb_batch_barrier_solve as they lie on the device: B rows of *row_doubles = n +
 * CVXB_BATCH_RECORD_EXTRA doubles, [x(n), objective, duality gap, status, newton steps, outer stages] -- the buffer the
 * multi-GPU driver hands to ONE NCCL all-gather without staging through the host.  Owned by the batch. */
enum { CVXB_BATCH_RECORD_EXTRA = 5 };
int cvxb_batch_device_records(cvxb_batch batch, double** dev_records, int* row_doubles);

/* ---- calibration / measurement helpers (bench.py, not part of the reference surface) ---------- */
/* C = alpha*op(A)op(B) + beta*C through the DMMA kernel on HOST column-major arrays (tests) */
int cvxb_test_dgemm(cvxb_handle h, int a_kc, int b_kc, int M, int N, int K, double alpha,
                    const double* A, int lda, const double* B, int ldb, double beta, double* C,
                    int ldc, int tri);
/* times `reps` launches of one resident kernel; returns average milliseconds per launch.
 * which: 0 = DMMA issue-rate peak (registers only), 1 = SYRK-TN n x n x k, 2 = SYRK-NT trailing
 * update n x n x k, 3 = blocked Cholesky n, 4 = HBM copy of n*k doubles */
int cvxb_bench_kernel(cvxb_handle h, int which, int n, int k, int reps, double* ms_per_launch,
                      double* flops_or_bytes_per_launch);

/* Schedule of the big factorisations (tests and tuning; defaults: blocks of 2048 columns from n = 5120 on, 8 SMs left to
 * the critical chain): the tile-DAG schedule runs the bulk updates of a right-looking blocked Cholesky on a third stream
 * beside the chain of diagonal-block factorisations.  A negative argument keeps the current value;
 * dag_block = 0 switches the schedule off (recursive halving + look-ahead only); all three negative restore the defaults,
 * including the automatic choice of narrower blocks / a smaller reserve when no wide right-hand-side block rides along. */
int cvxb_debug_set_schedule(cvxb_handle h, int dag_block, int dag_min_n, int dag_reserve);
/* The diagonal-block starts that schedule uses for an n x n matrix with blocks of dag_block columns, followed by n itself;
 * returns their number (host logic only: needs no GPU and no handle), -1 on bad arguments. */
int cvxb_debug_dag_blocks(int n, int dag_block, int* starts, int cap);

#ifdef __cplusplus
}
#endif
#endif /* CVXB_H */

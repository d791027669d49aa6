/* JNI shim between the Scala side (scala/cvx/CvxbNative.scala) and libcvxb's C ABI (include/cvxb.h).
 * One function per native method; arrays cross as double[] / int[] (Breeze DenseMatrix.data is a column-major
 * double[] with offset / majorStride, which maps 1:1 onto the (pointer, leading dimension) pairs of cvxb.h).
 *
 * Rules this file keeps (JNI specification, "Critical regions" and "Exceptions"):
 *   - no GetPrimitiveArrayCritical: every cvxb_* call allocates, synchronises a stream and may run for many
 *     milliseconds, and a critical region must not block or call back into JNI (it stalls the GC for every thread);
 *     arrays are taken with Get<Type>ArrayElements and released before the function returns;
 *   - exceptions are constructed with the constructor the reference's class really has:
 *       cvx.LinSolveException(A: DenseMatrix, b: DenseVector, L: DenseMatrix, message: String)   LinSolveException.scala:11-17
 *       cvx.UnsolvableSystemException(msg: String)                                               UnsolvableSystemException.scala
 *       cvx.LineSearchFailedException(message: String)                                           LineSearchFailedException.scala
 *       cvx.CvxbInfeasibleException(msg: String)   (shim-owned; GpuSolver rethrows it as InfeasibleProblemException(report, tol),
 *                                                   InfeasibleProblemException.scala:6, which needs a FeasibilityReport)
 *     a class or constructor that cannot be found falls back to java.lang.RuntimeException with the same message.
 *
 * The build image has no JDK.  tests/c/jni_stub/jni.h declares the subset of the JNI this file uses, so that
 *   gcc -std=c11 -Wall -Werror -shared -fPIC -Itests/c/jni_stub -Iinclude jni/cvxb_jni.c -Lcvx_b200/lib -lcvxb
 * compiles it in the CPU test run (tests/test_boundary_cpu.py), and tests/c/fake_jvm.c implements that subset over
 * plain C arrays so that the GPU test run drives every native method below through a real libcvxb.so
 * (tests/test_boundary_gpu.py).  With a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/cvxb_jni.c \
 *       -Lcvx_b200/lib -lcvxb -o libcvxb_jni.so
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>
#include "cvxb.h"

#define HND(h) ((cvxb_handle)(intptr_t)(h))
#define PRB(p) ((cvxb_problem)(intptr_t)(p))

/* ---- exceptions ------------------------------------------------------------------------------------------------ */
static void throw_runtime(JNIEnv* env, const char* msg) {
  jclass rc = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (rc) (*env)->ThrowNew(env, rc, msg);
}

static void throw_with_string_ctor(JNIEnv* env, const char* cls_name, const char* msg) {
  jclass cls = (*env)->FindClass(env, cls_name);
  if (!cls) {                         /* NoClassDefFoundError pending: replace it by something the caller can catch */
    (*env)->ExceptionClear(env);
    throw_runtime(env, msg);
    return;
  }
  if ((*env)->ThrowNew(env, cls, msg) != 0) {       /* no (String) constructor */
    (*env)->ExceptionClear(env);
    throw_runtime(env, msg);
  }
}

/* LinSolveException is a case class whose only constructor is (DenseMatrix, DenseVector, DenseMatrix, String): the
 * matrices stay on the device, so the exception carries nulls and the message (the reference's handlers only catch
 * the type: KKTSystem.scala:54-63, UnconstrainedSolver.scala:57-66). */
static void throw_linsolve(JNIEnv* env, const char* msg) {
  jclass cls = (*env)->FindClass(env, "cvx/LinSolveException");
  jmethodID ctor = 0;
  if (cls)
    ctor = (*env)->GetMethodID(env, cls, "<init>",
                               "(Lbreeze/linalg/DenseMatrix;Lbreeze/linalg/DenseVector;Lbreeze/linalg/DenseMatrix;Ljava/lang/String;)V");
  if (!cls || !ctor) {
    (*env)->ExceptionClear(env);
    throw_runtime(env, msg);
    return;
  }
  jstring jmsg = (*env)->NewStringUTF(env, msg);
  jobject ex = jmsg ? (*env)->NewObject(env, cls, ctor, (jobject)0, (jobject)0, (jobject)0, jmsg) : 0;
  if (!ex) {
    (*env)->ExceptionClear(env);
    throw_runtime(env, msg);
    return;
  }
  (*env)->Throw(env, (jthrowable)ex);
}

/* java.lang.AssertionError has no public (String) constructor: use (Object) */
static void throw_assertion(JNIEnv* env, const char* msg) {
  jclass cls = (*env)->FindClass(env, "java/lang/AssertionError");
  jmethodID ctor = cls ? (*env)->GetMethodID(env, cls, "<init>", "(Ljava/lang/Object;)V") : 0;
  jstring jmsg = ctor ? (*env)->NewStringUTF(env, msg) : 0;
  jobject ex = jmsg ? (*env)->NewObject(env, cls, ctor, (jobject)jmsg) : 0;
  if (!ex) {
    (*env)->ExceptionClear(env);
    throw_runtime(env, msg);
    return;
  }
  (*env)->Throw(env, (jthrowable)ex);
}

static void throw_for(JNIEnv* env, int status) {
  const char* msg = cvxb_last_error();
  if ((*env)->ExceptionCheck(env)) return;          /* keep an exception that is already pending (e.g. OutOfMemoryError) */
  switch (status) {
    case CVXB_ELINSOLVE: throw_linsolve(env, msg); break;
    case CVXB_EUNSOLVABLE: throw_with_string_ctor(env, "cvx/UnsolvableSystemException", msg); break;
    case CVXB_ELINESEARCH: throw_with_string_ctor(env, "cvx/LineSearchFailedException", msg); break;
    case CVXB_ENOTFEASIBLE: throw_with_string_ctor(env, "java/lang/IllegalArgumentException", msg); break;
    case CVXB_EINFEASIBLE: throw_with_string_ctor(env, "cvx/CvxbInfeasibleException", msg); break;
    case CVXB_EDIM: throw_assertion(env, msg); break;
    case CVXB_ENOTIMPL: throw_with_string_ctor(env, "java/lang/UnsupportedOperationException", msg); break;
    default: throw_runtime(env, msg); break;
  }
}

/* ---- array access ------------------------------------------------------------------------------------------------ */
#define DGET(a) ((a) ? (*env)->GetDoubleArrayElements(env, (a), 0) : (jdouble*)0)
#define DPUT(a, p, mode) do { if ((a) && (p)) (*env)->ReleaseDoubleArrayElements(env, (a), (p), (mode)); } while (0)
#define IGET(a) ((a) ? (*env)->GetIntArrayElements(env, (a), 0) : (jint*)0)
#define IPUT(a, p, mode) do { if ((a) && (p)) (*env)->ReleaseIntArrayElements(env, (a), (p), (mode)); } while (0)
/* a non-null array whose elements could not be obtained: OutOfMemoryError is pending */
#define MISSING(a, p) ((a) && !(p))

static void params_from(JNIEnv* env, jdoubleArray params, cvxb_params* P) {
  cvxb_default_params(P);
  if (!params) return;
  jdouble pv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  jsize len = (*env)->GetArrayLength(env, params);
  (*env)->GetDoubleArrayRegion(env, params, 0, len < 8 ? len : 8, pv);
  /* maxIter alpha beta tolSolver tolEqSolve tolFeas delta [bugCompat]   (SolverParams.scala:24-32) */
  if (len >= 7) {
    P->maxIter = (int)pv[0]; P->alpha = pv[1]; P->beta = pv[2]; P->tolSolver = pv[3]; P->tolEqSolve = pv[4];
    P->tolFeas = pv[5]; P->delta = pv[6];
  }
  if (len >= 8) P->bugCompat = pv[7] != 0.0;
}

static void stats_out(JNIEnv* env, jdoubleArray stats, const cvxb_solution* s) {
  /* [0] newtonDecrement [1] dualityGap [2] equalityGap [3] normGrad [4] normDualResidual [5] iter [6] maxedOut
   * [7] has-flags bitmask (1 nd, 2 gap, 4 eqGap, 8 normGrad, 16 normDualResidual, 32 lambda, 64 nu)
   * [8] newton steps [9] outer stages [10] objective [11] device ms [12] phase-I newton steps [13] phase-I stages
   * [14] phase-I slack s [15] line-search trials */
  if (!stats) return;
  jdouble out[16] = {s->newtonDecrement, s->dualityGap, s->equalityGap, s->normGrad, s->normDualResidual, (double)s->iter,
                     (double)s->maxedOut,
                     (double)(s->has_newtonDecrement | s->has_dualityGap << 1 | s->has_equalityGap << 2 | s->has_normGrad << 3 |
                              s->has_normDualResidual << 4 | s->has_lambda << 5 | s->has_nu << 6),
                     (double)s->newton_steps, (double)s->outer_stages, s->objective, s->solve_ms,
                     (double)s->phase1_newton_steps, (double)s->phase1_stages, s->phase1_s, (double)s->linesearch_trials};
  jsize len = (*env)->GetArrayLength(env, stats);
  (*env)->SetDoubleArrayRegion(env, stats, 0, len < 16 ? len : 16, out);
}

static void info_out(JNIEnv* env, jintArray info, const cvxb_kkt_info* ki) {
  if (!info) return;
  jint v[4] = {ki->path, ki->regularized, ki->ruiz_sweeps, ki->chol_info};
  jsize len = (*env)->GetArrayLength(env, info);
  (*env)->SetIntArrayRegion(env, info, 0, len < 4 ? len : 4, v);
}

/* ---- lifetime ------------------------------------------------------------------------------------------------------ */
JNIEXPORT jlong JNICALL Java_cvx_CvxbNative_create(JNIEnv* env, jclass c, jint device) {
  (void)c;
  cvxb_handle h = 0;
  int st = cvxb_create(device, 0, 0, &h);
  if (st != CVXB_OK) { throw_for(env, st); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_cvx_CvxbNative_destroy(JNIEnv* env, jclass c, jlong h) {
  (void)env; (void)c;
  cvxb_destroy(HND(h));
}

/* ---- seam B ---------------------------------------------------------------------------------------------------------- */
/* KKTSystem.solve (KKTSystem.scala:43-66): fills x (n) and w (p); info = {path, regularized, ruizSweeps, cholInfo} */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_kktSolve(JNIEnv* env, jclass c, jlong h, jint n, jint p, jdoubleArray H,
                                                    jint hOff, jint ldh, jdoubleArray A, jint aOff, jint lda,
                                                    jdoubleArray q, jdoubleArray b, jdouble tol, jdoubleArray x,
                                                    jdoubleArray w, jintArray info) {
  (void)c;
  jdouble *pH = DGET(H), *pA = DGET(A), *pq = DGET(q), *pb = DGET(b), *px = DGET(x), *pw = DGET(w);
  cvxb_kkt_info ki;
  memset(&ki, 0, sizeof ki);
  int st = CVXB_EINVAL;
  int ok = !(MISSING(H, pH) || MISSING(A, pA) || MISSING(q, pq) || MISSING(b, pb) || MISSING(x, px) || MISSING(w, pw));
  if (ok) st = cvxb_kkt_solve(HND(h), n, p, pH ? pH + hOff : 0, ldh, pA ? pA + aOff : 0, lda, pq, pb, tol, px, pw, &ki);
  DPUT(w, pw, 0); DPUT(x, px, 0);
  DPUT(b, pb, JNI_ABORT); DPUT(q, pq, JNI_ABORT); DPUT(A, pA, JNI_ABORT); DPUT(H, pH, JNI_ABORT);
  if (!ok) return;
  info_out(env, info, &ki);
  if (st != CVXB_OK) throw_for(env, st);
}

/* KKTSystem.solveWithCholFactor (KKTSystem.scala:99-167) */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_kktSolveWithCholFactor(JNIEnv* env, jclass c, jlong h, jint n, jint p,
                                                                  jdoubleArray L, jint lOff, jint ldl, jdoubleArray A,
                                                                  jint aOff, jint lda, jdoubleArray q, jdoubleArray b,
                                                                  jdouble tol, jdoubleArray x, jdoubleArray w) {
  (void)c;
  jdouble *pL = DGET(L), *pA = DGET(A), *pq = DGET(q), *pb = DGET(b), *px = DGET(x), *pw = DGET(w);
  int st = CVXB_EINVAL;
  int ok = !(MISSING(L, pL) || MISSING(A, pA) || MISSING(q, pq) || MISSING(b, pb) || MISSING(x, px) || MISSING(w, pw));
  if (ok)
    st = cvxb_kkt_solve_with_chol_factor(HND(h), n, p, pL ? pL + lOff : 0, ldl, pA ? pA + aOff : 0, lda, pq, pb, tol, px, pw, 0);
  DPUT(w, pw, 0); DPUT(x, px, 0);
  DPUT(b, pb, JNI_ABORT); DPUT(q, pq, JNI_ABORT); DPUT(A, pA, JNI_ABORT); DPUT(L, pL, JNI_ABORT);
  if (ok && st != CVXB_OK) throw_for(env, st);
}

/* MatrixUtils.choleskySolve (MatrixUtils.scala:468-516) */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_choleskySolve(JNIEnv* env, jclass c, jlong h, jint n, jdoubleArray H, jint hOff,
                                                         jint ldh, jdoubleArray b, jdouble tol, jdoubleArray x) {
  (void)c;
  jdouble *pH = DGET(H), *pb = DGET(b), *px = DGET(x);
  int st = CVXB_EINVAL;
  int ok = !(MISSING(H, pH) || MISSING(b, pb) || MISSING(x, px));
  if (ok) st = cvxb_cholesky_solve(HND(h), n, pH ? pH + hOff : 0, ldh, pb, tol, px, 0);
  DPUT(x, px, 0); DPUT(b, pb, JNI_ABORT); DPUT(H, pH, JNI_ABORT);
  if (ok && st != CVXB_OK) throw_for(env, st);
}

/* SymmetricLinearSystem.solve (SymmetricLinearSystem.scala:15-56) */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_symmetricSolve(JNIEnv* env, jclass c, jlong h, jint n, jdoubleArray H, jint hOff,
                                                          jint ldh, jdoubleArray r, jdouble tol, jdoubleArray x) {
  (void)c;
  jdouble *pH = DGET(H), *pr = DGET(r), *px = DGET(x);
  int st = CVXB_EINVAL;
  int ok = !(MISSING(H, pH) || MISSING(r, pr) || MISSING(x, px));
  if (ok) st = cvxb_symmetric_solve(HND(h), n, pH ? pH + hOff : 0, ldh, pr, tol, px, 0);
  DPUT(x, px, 0); DPUT(r, pr, JNI_ABORT); DPUT(H, pH, JNI_ABORT);
  if (ok && st != CVXB_OK) throw_for(env, st);
}

/* SolutionSpace(A, b) / MatrixUtils.solveUnderdetermined: fills z0 (n) and F (n x (n-p), column-major, ld n) */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_solveUnderdetermined(JNIEnv* env, jclass c, jlong h, jint p, jint n,
                                                                jdoubleArray A, jint aOff, jint lda, jdoubleArray b,
                                                                jdoubleArray z0, jdoubleArray F) {
  (void)c;
  jdouble *pA = DGET(A), *pb = DGET(b), *pz = DGET(z0), *pF = DGET(F);
  int st = CVXB_EINVAL;
  int ok = !(MISSING(A, pA) || MISSING(b, pb) || MISSING(z0, pz) || MISSING(F, pF));
  if (ok) st = cvxb_solve_underdetermined(HND(h), p, n, pA ? pA + aOff : 0, lda, pb, pz, pF, n);
  DPUT(F, pF, 0); DPUT(z0, pz, 0); DPUT(b, pb, JNI_ABORT); DPUT(A, pA, JNI_ABORT);
  if (ok && st != CVXB_OK) throw_for(env, st);
}

/* KKTData.reduced -> KKTSystem.solve -> paddVector; nullIdx (n ints) receives the eliminated indices, returns their number */
JNIEXPORT jint JNICALL Java_cvx_CvxbNative_kktSolveReduced(JNIEnv* env, jclass c, jlong h, jint n, jint p, jdoubleArray H,
                                                           jint ldh, jdoubleArray A, jint lda, jdoubleArray g,
                                                           jdoubleArray r, jdouble tol, jdoubleArray x, jdoubleArray w,
                                                           jintArray nullIdx) {
  (void)c;
  jdouble *pH = DGET(H), *pA = DGET(A), *pg = DGET(g), *pr = DGET(r), *px = DGET(x), *pw = DGET(w);
  jint* pi = IGET(nullIdx);
  int nn = 0, st = CVXB_EINVAL;
  int ok = !(MISSING(H, pH) || MISSING(A, pA) || MISSING(g, pg) || MISSING(r, pr) || MISSING(x, px) || MISSING(w, pw) ||
             MISSING(nullIdx, pi));
  if (ok) st = cvxb_kkt_solve_reduced(HND(h), n, p, pH, ldh, pA, lda, pg, pr, tol, px, pw, (int*)pi, &nn, 0);
  IPUT(nullIdx, pi, 0);
  DPUT(w, pw, 0); DPUT(x, px, 0);
  DPUT(r, pr, JNI_ABORT); DPUT(g, pg, JNI_ABORT); DPUT(A, pA, JNI_ABORT); DPUT(H, pH, JNI_ABORT);
  if (ok && st != CVXB_OK) throw_for(env, st);
  return nn;
}

/* ---- seam A ---------------------------------------------------------------------------------------------------------- */
/* cvxb_problem_create for the closed-form families; kind: cvxb_objective_kind (0 linear, 1 quadratic, 2 KL, 4 p-norm).
 * Null arrays = absent.  Quadratic constraints (QuadraticConstraint.scala:7-40): mq of them, qP = mq packed n x n
 * column-major matrices, qA = n x mq column-major, qR, qUb of length mq. */
JNIEXPORT jlong JNICALL Java_cvx_CvxbNative_problemCreate(JNIEnv* env, jclass c, jlong h, jint n, jint m, jint p, jint kind,
                                                          jdoubleArray objA, jdouble objR, jdoubleArray objP, jdouble objPow,
                                                          jdoubleArray G, jdoubleArray gR, jdoubleArray ub,
                                                          jdoubleArray A, jdoubleArray b, jdoubleArray xFeasible,
                                                          jdoubleArray xDefined, jint mq, jdoubleArray qP, jdoubleArray qA,
                                                          jdoubleArray qR, jdoubleArray qUb) {
  (void)c;
  cvxb_problem_desc d;
  memset(&d, 0, sizeof d);
  jdouble *pa = DGET(objA), *pP = DGET(objP), *pG = DGET(G), *pgr = DGET(gR), *pub = DGET(ub), *pA = DGET(A), *pb = DGET(b),
          *pxf = DGET(xFeasible), *pxd = DGET(xDefined), *pqP = DGET(qP), *pqA = DGET(qA), *pqR = DGET(qR), *pqU = DGET(qUb);
  int ok = !(MISSING(objA, pa) || MISSING(objP, pP) || MISSING(G, pG) || MISSING(gR, pgr) || MISSING(ub, pub) || MISSING(A, pA) ||
             MISSING(b, pb) || MISSING(xFeasible, pxf) || MISSING(xDefined, pxd) || MISSING(qP, pqP) || MISSING(qA, pqA) ||
             MISSING(qR, pqR) || MISSING(qUb, pqU));
  d.n = n; d.m = m; d.p = p; d.objective = kind; d.obj_a = pa; d.obj_r = objR; d.obj_P = pP; d.obj_ldP = n; d.obj_pow = objPow;
  d.G = pG; d.ldg = m; d.g_r = pgr; d.ub = pub; d.A = pA; d.lda = p; d.b = pb; d.x_feasible = pxf; d.x_defined = pxd;
  d.mq = mq; d.q_P = pqP; d.q_a = pqA; d.q_r = pqR; d.q_ub = pqU;
  cvxb_problem prob = 0;
  int st = CVXB_EINVAL;
  if (ok) st = cvxb_problem_create(HND(h), &d, &prob);   /* copies everything to the device */
  DPUT(objA, pa, JNI_ABORT); DPUT(objP, pP, JNI_ABORT); DPUT(G, pG, JNI_ABORT); DPUT(gR, pgr, JNI_ABORT);
  DPUT(ub, pub, JNI_ABORT); DPUT(A, pA, JNI_ABORT); DPUT(b, pb, JNI_ABORT); DPUT(xFeasible, pxf, JNI_ABORT);
  DPUT(xDefined, pxd, JNI_ABORT); DPUT(qP, pqP, JNI_ABORT); DPUT(qA, pqA, JNI_ABORT); DPUT(qR, pqR, JNI_ABORT);
  DPUT(qUb, pqU, JNI_ABORT);
  if (!ok) return 0;
  if (st != CVXB_OK) { throw_for(env, st); return 0; }
  return (jlong)(intptr_t)prob;
}

JNIEXPORT void JNICALL Java_cvx_CvxbNative_problemDestroy(JNIEnv* env, jclass c, jlong prob) {
  (void)env; (void)c;
  cvxb_problem_destroy(PRB(prob));
}

/* solver: 0 = BarrierSolver.solve, 1 = PrimalDualSolver.solve.  stats (double[16]): see stats_out. */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_solve(JNIEnv* env, jclass c, jlong h, jlong prob, jint solver,
                                                 jdoubleArray params, jdoubleArray x, jdoubleArray lambda,
                                                 jdoubleArray nu, jdoubleArray stats) {
  (void)c;
  cvxb_params P;
  params_from(env, params, &P);
  cvxb_solution s;
  memset(&s, 0, sizeof s);
  s.x = DGET(x); s.lambda = DGET(lambda); s.nu = DGET(nu);
  int ok = !(MISSING(x, s.x) || MISSING(lambda, s.lambda) || MISSING(nu, s.nu));
  int st = CVXB_EINVAL;
  if (ok) st = solver == 0 ? cvxb_barrier_solve(HND(h), PRB(prob), &P, &s) : cvxb_pd_solve(HND(h), PRB(prob), &P, &s);
  DPUT(x, s.x, 0); DPUT(lambda, s.lambda, 0); DPUT(nu, s.nu, 0);
  if (!ok) return;
  stats_out(env, stats, &s);
  if (st != CVXB_OK) throw_for(env, st);
}

/* ConstraintSet.withFeasiblePoint / phase_I_Analysis (ConstraintSet.scala:326-395, 556-575): xFeasible (n) and the
 * phase-I iterate xs = (x, s) (n + 1); CVXB_EINFEASIBLE -> CvxbInfeasibleException */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_phase1(JNIEnv* env, jclass c, jlong h, jlong prob, jdoubleArray params,
                                                  jdoubleArray xFeasible, jdoubleArray xs, jdoubleArray stats) {
  (void)c;
  cvxb_params P;
  params_from(env, params, &P);
  cvxb_solution s;
  memset(&s, 0, sizeof s);
  jdouble* pxf = DGET(xFeasible);
  s.x = DGET(xs);
  int ok = !(MISSING(xFeasible, pxf) || MISSING(xs, s.x));
  int st = CVXB_EINVAL;
  if (ok) st = cvxb_phase1(HND(h), PRB(prob), &P, pxf, &s);
  DPUT(xs, s.x, 0); DPUT(xFeasible, pxf, 0);
  if (!ok) return;
  stats_out(env, stats, &s);
  if (st != CVXB_OK) throw_for(env, st);
}

/* One barrier Newton direction at (x, t) (BarrierSolver.scala:291-315 + KKTSystem.solve / choleskySolve); H (n*n) may be null */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_barrierNewtonDirection(JNIEnv* env, jclass c, jlong h, jlong prob,
                                                                  jdoubleArray params, jdoubleArray x, jdouble t,
                                                                  jdoubleArray H, jdoubleArray g, jdoubleArray dx,
                                                                  jdoubleArray nu, jintArray info) {
  (void)c;
  cvxb_params P;
  params_from(env, params, &P);
  jdouble *px = DGET(x), *pH = DGET(H), *pg = DGET(g), *pdx = DGET(dx), *pnu = DGET(nu);
  cvxb_kkt_info ki;
  memset(&ki, 0, sizeof ki);
  int ok = !(MISSING(x, px) || MISSING(H, pH) || MISSING(g, pg) || MISSING(dx, pdx) || MISSING(nu, pnu));
  int st = CVXB_EINVAL;
  if (ok) st = cvxb_barrier_newton_direction(HND(h), PRB(prob), &P, px, t, pH, pg, pdx, pnu, &ki);
  DPUT(nu, pnu, 0); DPUT(dx, pdx, 0); DPUT(g, pg, 0); DPUT(H, pH, 0); DPUT(x, px, JNI_ABORT);
  if (!ok) return;
  info_out(env, info, &ki);
  if (st != CVXB_OK) throw_for(env, st);
}

/* One primal-dual search direction at (x, lambda, nu, t) (PrimalDualSolver.scala:162-285) */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_pdNewtonDirection(JNIEnv* env, jclass c, jlong h, jlong prob, jdoubleArray params,
                                                             jdoubleArray x, jdoubleArray lambda, jdoubleArray nu, jdouble t,
                                                             jdoubleArray H, jdoubleArray dx, jdoubleArray dlambda,
                                                             jdoubleArray dnu, jintArray info) {
  (void)c;
  cvxb_params P;
  params_from(env, params, &P);
  jdouble *px = DGET(x), *pl = DGET(lambda), *pn = DGET(nu), *pH = DGET(H), *pdx = DGET(dx), *pdl = DGET(dlambda),
          *pdn = DGET(dnu);
  cvxb_kkt_info ki;
  memset(&ki, 0, sizeof ki);
  int ok = !(MISSING(x, px) || MISSING(lambda, pl) || MISSING(nu, pn) || MISSING(H, pH) || MISSING(dx, pdx) ||
             MISSING(dlambda, pdl) || MISSING(dnu, pdn));
  int st = CVXB_EINVAL;
  if (ok) st = cvxb_pd_newton_direction(HND(h), PRB(prob), &P, px, pl, pn, t, pH, pdx, pdl, pdn, &ki);
  DPUT(dnu, pdn, 0); DPUT(dlambda, pdl, 0); DPUT(dx, pdx, 0); DPUT(H, pH, 0);
  DPUT(nu, pn, JNI_ABORT); DPUT(lambda, pl, JNI_ABORT); DPUT(x, px, JNI_ABORT);
  if (!ok) return;
  info_out(env, info, &ki);
  if (st != CVXB_OK) throw_for(env, st);
}

/* g_i(x) for every constraint of an uploaded problem; returns 1 when ConstraintSet.isSatisfiedStrictlyBy(x) */
JNIEXPORT jint JNICALL Java_cvx_CvxbNative_constraintValues(JNIEnv* env, jclass c, jlong h, jlong problem, jdoubleArray x,
                                                            jdoubleArray g) {
  (void)c;
  jdouble *px = DGET(x), *pg = DGET(g);
  int okv = 0, st = CVXB_EINVAL;
  int ok = !(MISSING(x, px) || MISSING(g, pg));
  if (ok) st = cvxb_constraint_values(HND(h), PRB(problem), px, pg, &okv);
  DPUT(g, pg, 0); DPUT(x, px, JNI_ABORT);
  if (ok && st != CVXB_OK) throw_for(env, st);
  return okv;
}

/* ---- equality elimination (Solver.affineTransformed / reduced) --------------------------------------------------------- */
JNIEXPORT jlong JNICALL Java_cvx_CvxbNative_solutionSpaceCreate(JNIEnv* env, jclass c, jlong h, jint p, jint n,
                                                                jdoubleArray A, jint aOff, jint lda, jdoubleArray b) {
  (void)c;
  jdouble *pA = DGET(A), *pb = DGET(b);
  cvxb_solution_space sp = 0;
  int st = CVXB_EINVAL;
  int ok = !(MISSING(A, pA) || MISSING(b, pb));
  if (ok) st = cvxb_solution_space_create(HND(h), p, n, pA ? pA + aOff : 0, lda, pb, &sp);
  DPUT(b, pb, JNI_ABORT); DPUT(A, pA, JNI_ABORT);
  if (!ok) return 0;
  if (st != CVXB_OK) { throw_for(env, st); return 0; }
  return (jlong)(intptr_t)sp;
}

/* the space x = z0 + F u given by the caller's own (z0, F) (Solver.affineTransformed, Solver.scala:46): F is n x k */
JNIEXPORT jlong JNICALL Java_cvx_CvxbNative_solutionSpaceFromBasis(JNIEnv* env, jclass c, jlong h, jint n, jint k,
                                                                   jdoubleArray z0, jdoubleArray F, jint fOff, jint ldf) {
  (void)c;
  jdouble *pz = DGET(z0), *pF = DGET(F);
  cvxb_solution_space sp = 0;
  int st = CVXB_EINVAL;
  int ok = !(MISSING(z0, pz) || MISSING(F, pF));
  if (ok) st = cvxb_solution_space_from_basis(HND(h), n, k, pz, pF ? pF + fOff : 0, ldf, &sp);
  DPUT(F, pF, JNI_ABORT); DPUT(z0, pz, JNI_ABORT);
  if (!ok) return 0;
  if (st != CVXB_OK) { throw_for(env, st); return 0; }
  return (jlong)(intptr_t)sp;
}

JNIEXPORT void JNICALL Java_cvx_CvxbNative_solutionSpaceDestroy(JNIEnv* env, jclass c, jlong space) {
  (void)env; (void)c;
  cvxb_solution_space_destroy((cvxb_solution_space)(intptr_t)space);
}

/* x = z0 + F u */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_solutionSpaceMap(JNIEnv* env, jclass c, jlong h, jlong space, jdoubleArray u,
                                                            jdoubleArray x) {
  (void)c;
  jdouble *pu = DGET(u), *px = DGET(x);
  int st = CVXB_EINVAL;
  int ok = !(MISSING(u, pu) || MISSING(x, px));
  if (ok) st = cvxb_solution_space_map(HND(h), (cvxb_solution_space)(intptr_t)space, pu, px);
  DPUT(x, px, 0); DPUT(u, pu, JNI_ABORT);
  if (ok && st != CVXB_OK) throw_for(env, st);
}

/* BarrierSolver.reduced / PrimalDualSolver.reduced: the problem in the variable u; returns the reduced problem */
JNIEXPORT jlong JNICALL Java_cvx_CvxbNative_problemReduce(JNIEnv* env, jclass c, jlong h, jlong prob, jlong space,
                                                          jdoubleArray params) {
  (void)c;
  cvxb_params P;
  params_from(env, params, &P);
  cvxb_problem red = 0;
  int st = cvxb_problem_reduce(HND(h), PRB(prob), (cvxb_solution_space)(intptr_t)space, &P, &red);
  if (st != CVXB_OK) { throw_for(env, st); return 0; }
  return (jlong)(intptr_t)red;
}

/* ---- batched small problems (cvxb_batch_*): returns the device milliseconds of the solve ------------------------------- */
JNIEXPORT jdouble JNICALL Java_cvx_CvxbNative_batchSolve(JNIEnv* env, jclass c, jlong h, jint B, jint n, jint m, jint p,
                                                         jintArray objective, jintArray pcount, jdoubleArray objA,
                                                         jdoubleArray objR, jdoubleArray objP, jdoubleArray G,
                                                         jdoubleArray ub, jdoubleArray A, jdoubleArray b, jdoubleArray x0,
                                                         jdoubleArray params, jdoubleArray x, jintArray status,
                                                         jintArray newtonSteps, jintArray outerStages,
                                                         jdoubleArray objectiveOut, jdoubleArray dualityGap,
                                                         jdoubleArray equalityGap, jintArray phase1,
                                                         jintArray phase1NewtonSteps) {
  (void)c;
  cvxb_params P;
  params_from(env, params, &P);
  jint *po = IGET(objective), *pc = IGET(pcount), *pst = IGET(status), *pns = IGET(newtonSteps), *pos = IGET(outerStages);
  /* phase1 (may be null): problems whose x0 is only a point where they are defined run the phase-I analysis first
   * (ConstraintSet.withFeasiblePoint, ConstraintSet.scala:556-575); phase1NewtonSteps (may be null) receives its steps */
  jint *pp1 = IGET(phase1), *pp1n = IGET(phase1NewtonSteps);
  jdouble *pa = DGET(objA), *pr = DGET(objR), *pP = DGET(objP), *pG = DGET(G), *pub = DGET(ub), *pA = DGET(A), *pb = DGET(b),
          *px0 = DGET(x0), *px = DGET(x), *pov = DGET(objectiveOut), *pgap = DGET(dualityGap), *peq = DGET(equalityGap);
  int ok = !(MISSING(objective, po) || MISSING(pcount, pc) || MISSING(status, pst) || MISSING(newtonSteps, pns) ||
             MISSING(outerStages, pos) || MISSING(objA, pa) || MISSING(objR, pr) || MISSING(objP, pP) || MISSING(G, pG) ||
             MISSING(ub, pub) || MISSING(A, pA) || MISSING(b, pb) || MISSING(x0, px0) || MISSING(x, px) ||
             MISSING(objectiveOut, pov) || MISSING(dualityGap, pgap) || MISSING(equalityGap, peq) || MISSING(phase1, pp1) ||
             MISSING(phase1NewtonSteps, pp1n));
  cvxb_batch_desc d;
  memset(&d, 0, sizeof d);
  d.B = B; d.n = n; d.m = m; d.p = p; d.objective = (const int*)po; d.pcount = (const int*)pc; d.obj_a = pa; d.obj_r = pr;
  d.obj_P = pP; d.G = pG; d.ub = pub; d.A = pA; d.b = pb; d.x0 = px0; d.phase1 = (const int*)pp1;
  cvxb_batch_result r;
  memset(&r, 0, sizeof r);
  r.x = px; r.status = (int*)pst; r.newton_steps = (int*)pns; r.outer_stages = (int*)pos; r.objective = pov;
  r.duality_gap = pgap; r.equality_gap = peq; r.phase1_newton_steps = (int*)pp1n;
  cvxb_batch bt = 0;
  int st = CVXB_EINVAL;
  if (ok) {
    st = cvxb_batch_create(HND(h), &d, &bt);
    if (st == CVXB_OK) st = cvxb_batch_barrier_solve(HND(h), bt, &P, &r);
    cvxb_batch_destroy(bt);
  }
  IPUT(phase1NewtonSteps, pp1n, 0); IPUT(phase1, pp1, JNI_ABORT);
  IPUT(outerStages, pos, 0); IPUT(newtonSteps, pns, 0); IPUT(status, pst, 0);
  DPUT(equalityGap, peq, 0); DPUT(dualityGap, pgap, 0); DPUT(objectiveOut, pov, 0); DPUT(x, px, 0);
  IPUT(pcount, pc, JNI_ABORT); IPUT(objective, po, JNI_ABORT);
  DPUT(x0, px0, JNI_ABORT); DPUT(b, pb, JNI_ABORT); DPUT(A, pA, JNI_ABORT); DPUT(ub, pub, JNI_ABORT); DPUT(G, pG, JNI_ABORT);
  DPUT(objP, pP, JNI_ABORT); DPUT(objR, pr, JNI_ABORT); DPUT(objA, pa, JNI_ABORT);
  if (!ok) return 0.0;
  if (st != CVXB_OK) { throw_for(env, st); return 0.0; }
  return r.solve_ms;
}

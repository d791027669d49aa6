/* JNI shim between the Scala side (scala/cvx/CvxbNative.scala) and libcvxb's C ABI (include/cvxb.h).
 * One function per native method; arrays cross as double[] (Breeze DenseMatrix.data is a column-major
 * double[] with offset / majorStride, which maps 1:1 onto the (pointer, leading dimension) pairs of cvxb.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD: the build image has no JDK (no jni.h).  Build where one exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/cvxb_jni.c \
 *       -Lcvx_b200/lib -lcvxb -o libcvxb_jni.so
 */
#include <jni.h>
#include <string.h>
#include "cvxb.h"

static void throw_for(JNIEnv* env, int status) {
  const char* cls;
  switch (status) {
    case CVXB_ELINSOLVE: cls = "cvx/LinSolveException"; break;
    case CVXB_EUNSOLVABLE: cls = "cvx/UnsolvableSystemException"; break;
    case CVXB_ELINESEARCH: cls = "cvx/LineSearchFailedException"; break;
    case CVXB_ENOTFEASIBLE: cls = "java/lang/IllegalArgumentException"; break;
    case CVXB_EINFEASIBLE: cls = "cvx/CvxbInfeasibleException"; break;   /* rethrown as InfeasibleProblemException */
    case CVXB_EDIM: cls = "java/lang/AssertionError"; break;
    default: cls = "java/lang/RuntimeException"; break;
  }
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), cvxb_last_error());
}

JNIEXPORT jlong JNICALL Java_cvx_CvxbNative_create(JNIEnv* env, jclass c, jint device) {
  cvxb_handle h = 0;
  int st = cvxb_create(device, 0, 0, &h);
  if (st != CVXB_OK) { throw_for(env, st); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_cvx_CvxbNative_destroy(JNIEnv* env, jclass c, jlong h) {
  cvxb_destroy((cvxb_handle)(intptr_t)h);
}

/* KKTSystem.solve: returns 0, fills x (n) and w (p); info = {path, regularized, ruizSweeps} */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_kktSolve(JNIEnv* env, jclass c, jlong h, jint n, jint p, jdoubleArray H,
                                                    jint hOff, jint ldh, jdoubleArray A, jint aOff, jint lda,
                                                    jdoubleArray q, jdoubleArray b, jdouble tol, jdoubleArray x,
                                                    jdoubleArray w, jintArray info) {
  jdouble* pH = (*env)->GetPrimitiveArrayCritical(env, H, 0);
  jdouble* pA = (*env)->GetPrimitiveArrayCritical(env, A, 0);
  jdouble* pq = (*env)->GetPrimitiveArrayCritical(env, q, 0);
  jdouble* pb = (*env)->GetPrimitiveArrayCritical(env, b, 0);
  jdouble* px = (*env)->GetPrimitiveArrayCritical(env, x, 0);
  jdouble* pw = (*env)->GetPrimitiveArrayCritical(env, w, 0);
  cvxb_kkt_info ki;
  memset(&ki, 0, sizeof ki);
  int st = cvxb_kkt_solve((cvxb_handle)(intptr_t)h, n, p, pH + hOff, ldh, pA + aOff, lda, pq, pb, tol, px, pw, &ki);
  (*env)->ReleasePrimitiveArrayCritical(env, w, pw, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, x, px, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, b, pb, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, q, pq, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, A, pA, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, H, pH, JNI_ABORT);
  if (info) { jint v[3] = {ki.path, ki.regularized, ki.ruiz_sweeps}; (*env)->SetIntArrayRegion(env, info, 0, 3, v); }
  if (st != CVXB_OK) throw_for(env, st);
}

JNIEXPORT void JNICALL Java_cvx_CvxbNative_choleskySolve(JNIEnv* env, jclass c, jlong h, jint n, jdoubleArray H, jint hOff,
                                                         jint ldh, jdoubleArray b, jdouble tol, jdoubleArray x) {
  jdouble* pH = (*env)->GetPrimitiveArrayCritical(env, H, 0);
  jdouble* pb = (*env)->GetPrimitiveArrayCritical(env, b, 0);
  jdouble* px = (*env)->GetPrimitiveArrayCritical(env, x, 0);
  int st = cvxb_cholesky_solve((cvxb_handle)(intptr_t)h, n, pH + hOff, ldh, pb, tol, px, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, x, px, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, b, pb, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, H, pH, JNI_ABORT);
  if (st != CVXB_OK) throw_for(env, st);
}

/* cvxb_problem_create for the closed-form families; kind: 0 linear, 1 quadratic, 2 KL.  Null arrays = absent. */
JNIEXPORT jlong JNICALL Java_cvx_CvxbNative_problemCreate(JNIEnv* env, jclass c, jlong h, jint n, jint m, jint p, jint kind,
                                                          jdoubleArray objA, jdouble objR, jdoubleArray objP,
                                                          jdoubleArray G, jdoubleArray gR, jdoubleArray ub,
                                                          jdoubleArray A, jdoubleArray b, jdoubleArray xFeasible,
                                                          jdoubleArray xDefined) {
#define PIN(a) ((a) ? (*env)->GetDoubleArrayElements(env, (a), 0) : 0)
#define UNPIN(a, ptr) if (a) (*env)->ReleaseDoubleArrayElements(env, (a), (ptr), JNI_ABORT)
  cvxb_problem_desc d;
  memset(&d, 0, sizeof d);
  jdouble *pa = PIN(objA), *pP = PIN(objP), *pG = PIN(G), *pgr = PIN(gR), *pub = PIN(ub), *pA = PIN(A), *pb = PIN(b),
          *pxf = PIN(xFeasible), *pxd = PIN(xDefined);
  d.n = n; d.m = m; d.p = p; d.objective = kind; d.obj_a = pa; d.obj_r = objR; d.obj_P = pP; d.obj_ldP = n;
  d.G = pG; d.ldg = m; d.g_r = pgr; d.ub = pub; d.A = pA; d.lda = p; d.b = pb; d.x_feasible = pxf; d.x_defined = pxd;
  cvxb_problem prob = 0;
  int st = cvxb_problem_create((cvxb_handle)(intptr_t)h, &d, &prob);   /* copies everything to the device */
  UNPIN(objA, pa); UNPIN(objP, pP); UNPIN(G, pG); UNPIN(gR, pgr); UNPIN(ub, pub); UNPIN(A, pA); UNPIN(b, pb);
  UNPIN(xFeasible, pxf); UNPIN(xDefined, pxd);
  if (st != CVXB_OK) { throw_for(env, st); return 0; }
  return (jlong)(intptr_t)prob;
}

JNIEXPORT void JNICALL Java_cvx_CvxbNative_problemDestroy(JNIEnv* env, jclass c, jlong prob) {
  cvxb_problem_destroy((cvxb_problem)(intptr_t)prob);
}

/* solver: 0 = BarrierSolver.solve, 1 = PrimalDualSolver.solve.  stats (double[16]) receives the Solution fields:
 * [0] newtonDecrement [1] dualityGap [2] equalityGap [3] normGrad [4] normDualResidual [5] iter [6] maxedOut
 * [7] has-flags bitmask (1 nd, 2 gap, 4 eqGap, 8 normGrad, 16 normDualResidual, 32 lambda, 64 nu)
 * [8] newton steps [9] outer stages [10] objective [11] device ms */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_solve(JNIEnv* env, jclass c, jlong h, jlong prob, jint solver,
                                                 jdoubleArray params, jdoubleArray x, jdoubleArray lambda,
                                                 jdoubleArray nu, jdoubleArray stats) {
  cvxb_params P;
  cvxb_default_params(&P);
  jdouble pv[7];
  (*env)->GetDoubleArrayRegion(env, params, 0, 7, pv);   /* maxIter alpha beta tolSolver tolEqSolve tolFeas delta */
  P.maxIter = (int)pv[0]; P.alpha = pv[1]; P.beta = pv[2]; P.tolSolver = pv[3]; P.tolEqSolve = pv[4]; P.tolFeas = pv[5];
  P.delta = pv[6];
  cvxb_solution s;
  memset(&s, 0, sizeof s);
  s.x = (*env)->GetDoubleArrayElements(env, x, 0);
  s.lambda = lambda ? (*env)->GetDoubleArrayElements(env, lambda, 0) : 0;
  s.nu = nu ? (*env)->GetDoubleArrayElements(env, nu, 0) : 0;
  int st = solver == 0 ? cvxb_barrier_solve((cvxb_handle)(intptr_t)h, (cvxb_problem)(intptr_t)prob, &P, &s)
                       : cvxb_pd_solve((cvxb_handle)(intptr_t)h, (cvxb_problem)(intptr_t)prob, &P, &s);
  (*env)->ReleaseDoubleArrayElements(env, x, s.x, 0);
  if (lambda) (*env)->ReleaseDoubleArrayElements(env, lambda, s.lambda, 0);
  if (nu) (*env)->ReleaseDoubleArrayElements(env, nu, s.nu, 0);
  jdouble out[12] = {s.newtonDecrement, s.dualityGap, s.equalityGap, s.normGrad, s.normDualResidual, s.iter, s.maxedOut,
                     (double)(s.has_newtonDecrement | s.has_dualityGap << 1 | s.has_equalityGap << 2 | s.has_normGrad << 3 |
                              s.has_normDualResidual << 4 | s.has_lambda << 5 | s.has_nu << 6),
                     (double)s.newton_steps, s.outer_stages, s.objective, s.solve_ms};
  (*env)->SetDoubleArrayRegion(env, stats, 0, 12, out);
  if (st != CVXB_OK) throw_for(env, st);
}

/* SolutionSpace(A, b) / MatrixUtils.solveUnderdetermined: fills z0 (n) and F (n x (n-p), column-major, ld n) */
JNIEXPORT void JNICALL Java_cvx_CvxbNative_solveUnderdetermined(JNIEnv* env, jclass c, jlong h, jint p, jint n,
                                                                jdoubleArray A, jint aOff, jint lda, jdoubleArray b,
                                                                jdoubleArray z0, jdoubleArray F) {
  jdouble* pA = (*env)->GetPrimitiveArrayCritical(env, A, 0);
  jdouble* pb = (*env)->GetPrimitiveArrayCritical(env, b, 0);
  jdouble* pz = (*env)->GetPrimitiveArrayCritical(env, z0, 0);
  jdouble* pF = (*env)->GetPrimitiveArrayCritical(env, F, 0);
  int st = cvxb_solve_underdetermined((cvxb_handle)(intptr_t)h, p, n, pA + aOff, lda, pb, pz, pF, n);
  (*env)->ReleasePrimitiveArrayCritical(env, F, pF, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, z0, pz, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, b, pb, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, A, pA, JNI_ABORT);
  if (st != CVXB_OK) throw_for(env, st);
}

/* KKTData.reduced -> KKTSystem.solve -> paddVector; nullIdx (n ints) receives the eliminated indices, returns their number */
JNIEXPORT jint JNICALL Java_cvx_CvxbNative_kktSolveReduced(JNIEnv* env, jclass c, jlong h, jint n, jint p, jdoubleArray H,
                                                           jint ldh, jdoubleArray A, jint lda, jdoubleArray g,
                                                           jdoubleArray r, jdouble tol, jdoubleArray x, jdoubleArray w,
                                                           jintArray nullIdx) {
  jdouble* pH = (*env)->GetPrimitiveArrayCritical(env, H, 0);
  jdouble* pA = (*env)->GetPrimitiveArrayCritical(env, A, 0);
  jdouble* pg = (*env)->GetPrimitiveArrayCritical(env, g, 0);
  jdouble* pr = (*env)->GetPrimitiveArrayCritical(env, r, 0);
  jdouble* px = (*env)->GetPrimitiveArrayCritical(env, x, 0);
  jdouble* pw = (*env)->GetPrimitiveArrayCritical(env, w, 0);
  jint* pi = (*env)->GetPrimitiveArrayCritical(env, nullIdx, 0);
  int nn = 0;
  int st = cvxb_kkt_solve_reduced((cvxb_handle)(intptr_t)h, n, p, pH, ldh, pA, lda, pg, pr, tol, px, pw, (int*)pi, &nn, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, nullIdx, pi, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, w, pw, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, x, px, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, r, pr, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, g, pg, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, A, pA, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, H, pH, JNI_ABORT);
  if (st != CVXB_OK) throw_for(env, st);
  return nn;
}

/* g_i(x) for every constraint of an uploaded problem; returns 1 when ConstraintSet.isSatisfiedStrictlyBy(x) */
JNIEXPORT jint JNICALL Java_cvx_CvxbNative_constraintValues(JNIEnv* env, jclass c, jlong h, jlong problem, jdoubleArray x,
                                                            jdoubleArray g) {
  jdouble* px = (*env)->GetPrimitiveArrayCritical(env, x, 0);
  jdouble* pg = (*env)->GetPrimitiveArrayCritical(env, g, 0);
  int ok = 0;
  int st = cvxb_constraint_values((cvxb_handle)(intptr_t)h, (cvxb_problem)(intptr_t)problem, px, pg, &ok);
  (*env)->ReleasePrimitiveArrayCritical(env, g, pg, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, x, px, JNI_ABORT);
  if (st != CVXB_OK) throw_for(env, st);
  return ok;
}

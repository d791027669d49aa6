/* Executes the JNI shim (jni/cvxb_jni.c) through the fake JNIEnv of fake_jvm.c against a real libcvxb.so.  TEST
 * INFRASTRUCTURE (tests/test_boundary_gpu.py).  Checks the results of the native methods and, for the failure paths,
 * WHICH exception class and WHICH constructor the shim used -- the defect the round-1 shim had
 * (ThrowNew on cvx.LinSolveException, which has no (String) constructor, LinSolveException.scala:11-17). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fake_jvm.h"

jlong Java_cvx_CvxbNative_create(JNIEnv*, jclass, jint);
void Java_cvx_CvxbNative_destroy(JNIEnv*, jclass, jlong);
void Java_cvx_CvxbNative_kktSolve(JNIEnv*, jclass, jlong, jint, jint, jdoubleArray, jint, jint, jdoubleArray, jint, jint,
                                  jdoubleArray, jdoubleArray, jdouble, jdoubleArray, jdoubleArray, jintArray);
void Java_cvx_CvxbNative_choleskySolve(JNIEnv*, jclass, jlong, jint, jdoubleArray, jint, jint, jdoubleArray, jdouble, jdoubleArray);
jlong Java_cvx_CvxbNative_problemCreate(JNIEnv*, jclass, jlong, jint, jint, jint, jint, jdoubleArray, jdouble, jdoubleArray, jdouble,
                                        jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray,
                                        jdoubleArray, jint, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray);
void Java_cvx_CvxbNative_problemDestroy(JNIEnv*, jclass, jlong);
void Java_cvx_CvxbNative_solve(JNIEnv*, jclass, jlong, jlong, jint, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray);
void Java_cvx_CvxbNative_phase1(JNIEnv*, jclass, jlong, jlong, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray);
jint Java_cvx_CvxbNative_constraintValues(JNIEnv*, jclass, jlong, jlong, jdoubleArray, jdoubleArray);
jlong Java_cvx_CvxbNative_solutionSpaceCreate(JNIEnv*, jclass, jlong, jint, jint, jdoubleArray, jint, jint, jdoubleArray);
void Java_cvx_CvxbNative_solutionSpaceDestroy(JNIEnv*, jclass, jlong);
void Java_cvx_CvxbNative_solutionSpaceMap(JNIEnv*, jclass, jlong, jlong, jdoubleArray, jdoubleArray);
jlong Java_cvx_CvxbNative_problemReduce(JNIEnv*, jclass, jlong, jlong, jlong, jdoubleArray);
jdouble Java_cvx_CvxbNative_batchSolve(JNIEnv*, jclass, jlong, jint, jint, jint, jint, jintArray, jintArray, jdoubleArray,
                                       jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray,
                                       jdoubleArray, jdoubleArray, jdoubleArray, jintArray, jintArray, jintArray, jdoubleArray,
                                       jdoubleArray, jdoubleArray, jintArray, jintArray);

static unsigned long long g_seed = 1234567891011ull;
static double urand(double lo, double hi) {
  g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17;
  return lo + (hi - lo) * (double)(g_seed >> 11) / 9007199254740992.0;
}
#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "drive_jni FAILED: " __VA_ARGS__); \
  fprintf(stderr, " [pending %s: %s]\n", fake_pending_class(), fake_pending_message()); return 1; } } while (0)

int main(void) {
  JNIEnv* env = fake_env();
  jlong h = Java_cvx_CvxbNative_create(env, 0, 0);
  CHECK(h != 0 && !fake_pending_class()[0], "create");

  /* ---- kktSolve on a planted system, H handed over with an offset and a larger leading dimension (a Breeze view) ---- */
  const int n = 64, p = 8, ldh = n + 3, off = 5;
  jdoubleArray H = fake_new_double_array(off + ldh * n, 0), A = fake_new_double_array(p * n, 0);
  double *Hd = fake_doubles(H) + off, *Ad = fake_doubles(A);
  double* L = calloc((size_t)n * n, 8);
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) L[(size_t)j * n + i] = urand(-1, 1) + (i == j ? 12.0 : 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int k = 0; k <= (i < j ? i : j); ++k) s += L[(size_t)k * n + i] * L[(size_t)k * n + j];
      Hd[(size_t)j * ldh + i] = s;
    }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < p; ++i) Ad[(size_t)j * p + i] = urand(-1, 1) + (i == j ? 4.0 : 0.0);
  double xs[64], ws[8], qv[64], bv[8];
  for (int j = 0; j < n; ++j) xs[j] = urand(-1, 1);
  for (int i = 0; i < p; ++i) ws[i] = urand(-2, 2);
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += Hd[(size_t)j * ldh + i] * xs[j];
    for (int k = 0; k < p; ++k) s += Ad[(size_t)i * p + k] * ws[k];
    qv[i] = -s;
  }
  for (int k = 0; k < p; ++k) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += Ad[(size_t)j * p + k] * xs[j];
    bv[k] = s;
  }
  jdoubleArray q = fake_new_double_array(n, qv), b = fake_new_double_array(p, bv), x = fake_new_double_array(n, 0),
               w = fake_new_double_array(p, 0);
  jintArray info = fake_new_int_array(4, 0);
  Java_cvx_CvxbNative_kktSolve(env, 0, h, n, p, H, off, ldh, A, 0, p, q, b, 1e-10, x, w, info);
  CHECK(!fake_pending_class()[0], "kktSolve threw");
  double ex = 0;
  for (int j = 0; j < n; ++j) ex = fmax(ex, fabs(fake_doubles(x)[j] - xs[j]));
  for (int k = 0; k < p; ++k) ex = fmax(ex, fabs(fake_doubles(w)[k] - ws[k]));
  CHECK(ex < 1e-9 && fake_ints(info)[0] == 0, "kktSolve: max error %.3g, path %d", ex, fake_ints(info)[0]);
  CHECK(fake_doubles(q)[0] == qv[0], "kktSolve wrote into an input array");
  printf("kktSolve via JNI: max error %.2e, Ruiz sweeps %d\n", ex, fake_ints(info)[2]);

  /* ---- LinSolveException: must be built with its 4-argument constructor ---- */
  jdoubleArray Hneg = fake_new_double_array(n * n, 0);
  for (int j = 0; j < n; ++j) fake_doubles(Hneg)[(size_t)j * n + j] = -1.0;
  Java_cvx_CvxbNative_choleskySolve(env, 0, h, n, Hneg, 0, n, q, 1e-10, x);
  CHECK(!strcmp(fake_pending_class(), "cvx/LinSolveException"), "choleskySolve(-I) raised '%s'", fake_pending_class());
  CHECK(strstr(fake_pending_ctor(), "DenseMatrix;Lbreeze/linalg/DenseVector;Lbreeze/linalg/DenseMatrix;Ljava/lang/String;)V") != 0,
        "LinSolveException built with constructor %s", fake_pending_ctor());
  CHECK(strlen(fake_pending_message()) > 10, "LinSolveException without message");
  printf("LinSolveException via %s\n  message: %s\n", fake_pending_ctor(), fake_pending_message());
  fake_clear_pending();

  /* ---- AssertionError for a dimension error (KKTSystem.scala:31-32) ---- */
  Java_cvx_CvxbNative_kktSolve(env, 0, h, 0, p, H, off, ldh, A, 0, p, q, b, 1e-10, x, w, info);
  CHECK(!strcmp(fake_pending_class(), "java/lang/AssertionError") && !strcmp(fake_pending_ctor(), "(Ljava/lang/Object;)V"),
        "n = 0 raised '%s' via %s", fake_pending_class(), fake_pending_ctor());
  fake_clear_pending();

  /* ---- seam A: probability-simplex QP  min ||x - c||^2/2  s.t. x >= 0, sum x = 1 (p = 1), phase I from 1/d ---- */
  const int d = 16;
  double Pm[256] = {0}, cv[16], Gm[256] = {0}, ubv[16] = {0}, Am[16], bone[1] = {1.0}, xdef[16];
  for (int j = 0; j < d; ++j) { Pm[j * d + j] = 1.0; cv[j] = -urand(0, 0.05); Gm[j * d + j] = -1.0; Am[j] = 1.0; xdef[j] = 1.0 / d; }
  jdoubleArray jP = fake_new_double_array(d * d, Pm), ja = fake_new_double_array(d, cv), jG = fake_new_double_array(d * d, Gm),
               jub = fake_new_double_array(d, ubv), jA = fake_new_double_array(d, Am), jb = fake_new_double_array(1, bone),
               jxd = fake_new_double_array(d, xdef);
  jlong prob = Java_cvx_CvxbNative_problemCreate(env, 0, h, d, d, 1, 1 /* quadratic */, ja, 0.0, jP, 2.0, jG, 0, jub, jA, jb, 0, jxd,
                                                 0, 0, 0, 0, 0);
  CHECK(prob != 0 && !fake_pending_class()[0], "problemCreate");
  double prm[8] = {1000, 0.04, 0.8, 1e-8, 1e-1, 1e-7, 1e-6, 0};
  jdoubleArray jprm = fake_new_double_array(8, prm), jx = fake_new_double_array(d, 0), jl = fake_new_double_array(d, 0),
               jn = fake_new_double_array(1, 0), jst = fake_new_double_array(16, 0);
  for (int solver = 0; solver < 2; ++solver) {
    Java_cvx_CvxbNative_solve(env, 0, h, prob, solver, jprm, jx, jl, jn, jst);
    CHECK(!fake_pending_class()[0], "solve(%d) threw", solver);
    double s = 0, mn = 1e300;
    for (int j = 0; j < d; ++j) { s += fake_doubles(jx)[j]; mn = fmin(mn, fake_doubles(jx)[j]); }
    /* optimum of the projection onto the simplex: x_j = (-c_j) + tau, all positive here since sum(-c) < 1 */
    double tau = 1.0, err = 0;
    for (int j = 0; j < d; ++j) tau -= -cv[j];
    tau /= d;
    for (int j = 0; j < d; ++j) err = fmax(err, fabs(fake_doubles(jx)[j] - (-cv[j] + tau)));
    CHECK(fabs(s - 1.0) < 1e-8 && mn > 0 && err < 1e-6, "solve(%d): sum %.12g min %.3g err %.3g", solver, s, mn, err);
    printf("solve(%s) via JNI: max |x - x*| %.2e, gap %.2e, %d Newton steps\n", solver ? "PD" : "BR", err, fake_doubles(jst)[1],
           (int)fake_doubles(jst)[8]);
  }
  /* phase1 + constraintValues */
  jdoubleArray jxf = fake_new_double_array(d, 0), jxs = fake_new_double_array(d + 1, 0), jg = fake_new_double_array(d, 0);
  Java_cvx_CvxbNative_phase1(env, 0, h, prob, jprm, jxf, jxs, jst);
  CHECK(!fake_pending_class()[0], "phase1 threw");
  jint strict = Java_cvx_CvxbNative_constraintValues(env, 0, h, prob, jxf, jg);
  CHECK(strict == 1 && fake_doubles(jxs)[d] < 0.0, "phase1: strict %d, s %.3g", strict, fake_doubles(jxs)[d]);

  /* ---- equality elimination: reduce the same problem without its equality, solve in u, map back ---- */
  jlong prob0 = Java_cvx_CvxbNative_problemCreate(env, 0, h, d, d, 0, 1, ja, 0.0, jP, 2.0, jG, 0, jub, 0, 0, jxf, jxd, 0, 0, 0, 0, 0);
  jlong sp = Java_cvx_CvxbNative_solutionSpaceCreate(env, 0, h, 1, d, jA, 0, 1, jb);
  CHECK(prob0 && sp && !fake_pending_class()[0], "solutionSpaceCreate");
  jlong red = Java_cvx_CvxbNative_problemReduce(env, 0, h, prob0, sp, jprm);
  CHECK(red && !fake_pending_class()[0], "problemReduce");
  jdoubleArray ju = fake_new_double_array(d - 1, 0), jl2 = fake_new_double_array(d, 0), jxm = fake_new_double_array(d, 0);
  Java_cvx_CvxbNative_solve(env, 0, h, red, 0, jprm, ju, jl2, 0, jst);
  CHECK(!fake_pending_class()[0], "solve(reduced) threw");
  Java_cvx_CvxbNative_solutionSpaceMap(env, 0, h, sp, ju, jxm);
  double dmax = 0;
  for (int j = 0; j < d; ++j) dmax = fmax(dmax, fabs(fake_doubles(jxm)[j] - fake_doubles(jx)[j]));
  CHECK(dmax < 1e-6, "reduced solve differs from the direct one by %.3g", dmax);
  printf("reduced solve via JNI: max difference to the direct solve %.2e\n", dmax);
  Java_cvx_CvxbNative_problemDestroy(env, 0, red);
  Java_cvx_CvxbNative_problemDestroy(env, 0, prob0);
  Java_cvx_CvxbNative_solutionSpaceDestroy(env, 0, sp);
  Java_cvx_CvxbNative_problemDestroy(env, 0, prob);

  /* ---- batched: two copies of a small box LP with a strictly feasible start ---- */
  const int B = 2, bn = 4, bm = 8;
  int kinds[2] = {0, 0}, pc[2] = {0, 0};
  double oa[8], orr[2] = {0, 0}, bG[64] = {0}, bub[16], bx0[8] = {0};
  for (int k = 0; k < B; ++k)
    for (int j = 0; j < bn; ++j) {
      oa[k * bn + j] = (j % 2 ? 1.0 : -1.0) * (1.0 + k);
      bG[k * bn * bm + j * bm + 2 * j] = 1.0; bG[k * bn * bm + j * bm + 2 * j + 1] = -1.0;
      bub[k * bm + 2 * j] = bub[k * bm + 2 * j + 1] = 1.0;
    }
  jintArray jk = fake_new_int_array(B, kinds), jpc = fake_new_int_array(B, pc), jbs = fake_new_int_array(B, 0),
            jbn = fake_new_int_array(B, 0), jbo = fake_new_int_array(B, 0);
  jdoubleArray joa = fake_new_double_array(B * bn, oa), jor = fake_new_double_array(B, orr), jbG = fake_new_double_array(B * bn * bm, bG),
               jbub = fake_new_double_array(B * bm, bub), jbx0 = fake_new_double_array(B * bn, bx0), jbx = fake_new_double_array(B * bn, 0),
               jbov = fake_new_double_array(B, 0), jbg = fake_new_double_array(B, 0), jbe = fake_new_double_array(B, 0);
  double ms = Java_cvx_CvxbNative_batchSolve(env, 0, h, B, bn, bm, 0, jk, jpc, joa, jor, 0, jbG, jbub, 0, 0, jbx0, jprm, jbx, jbs, jbn, jbo,
                                             jbov, jbg, jbe, 0, 0);
  CHECK(!fake_pending_class()[0] && ms > 0, "batchSolve threw");
  for (int k = 0; k < B; ++k)
    CHECK(fake_ints(jbs)[k] == 0 && fabs(fake_doubles(jbov)[k] + 4.0 * (1.0 + k)) < 1e-6, "batch problem %d: status %d objective %.9g", k,
          fake_ints(jbs)[k], fake_doubles(jbov)[k]);
  printf("batchSolve via JNI: objectives %.9f %.9f in %.3f ms\n", fake_doubles(jbov)[0], fake_doubles(jbov)[1], ms);
  /* the same two problems started outside the box: phase I inside the kernel first (phase1 flags), same optima */
  int ph[2] = {1, 1};
  double bx1[8];
  for (int j = 0; j < B * bn; ++j) bx1[j] = 2.0 + 0.1 * j;
  jintArray jph = fake_new_int_array(B, ph), jphn = fake_new_int_array(B, 0);
  jdoubleArray jbx1 = fake_new_double_array(B * bn, bx1);
  ms = Java_cvx_CvxbNative_batchSolve(env, 0, h, B, bn, bm, 0, jk, jpc, joa, jor, 0, jbG, jbub, 0, 0, jbx1, jprm, jbx, jbs, jbn, jbo, jbov, jbg,
                                      jbe, jph, jphn);
  CHECK(!fake_pending_class()[0] && ms > 0, "batchSolve (phase I) threw");
  for (int k = 0; k < B; ++k)
    CHECK(fake_ints(jbs)[k] == 0 && fake_ints(jphn)[k] > 0 && fabs(fake_doubles(jbov)[k] + 4.0 * (1.0 + k)) < 1e-6,
          "batch problem %d with phase I: status %d, %d phase-I steps, objective %.9g", k, fake_ints(jbs)[k], fake_ints(jphn)[k],
          fake_doubles(jbov)[k]);
  printf("batchSolve via JNI with phase I: %d + %d phase-I Newton steps, objectives %.9f %.9f\n", fake_ints(jphn)[0], fake_ints(jphn)[1],
         fake_doubles(jbov)[0], fake_doubles(jbov)[1]);

  CHECK(fake_outstanding_arrays() == 0, "%d array accesses were never released", fake_outstanding_arrays());
  Java_cvx_CvxbNative_destroy(env, 0, h);
  printf("drive_jni ok\n");
  return 0;
}

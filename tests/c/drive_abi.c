/* Plain C program against include/cvxb.h and libcvxb.so -- what a cgo / JNI / Panama binding sees.  TEST
 * INFRASTRUCTURE (tests/test_boundary_gpu.py runs it on the GPU box).  Designs follow the reference's own tests:
 * a planted KKT system (KktTest.testPositiveDefinite, src/test/scala/cvx/KktTest.scala:197-272) and a known-answer LP
 * (SimpleOptimizationProblems.minDotProduct, SimpleOptimizationProblems.scala:142-169: min -a'x, |x_j| <= |a_j|, x* = a). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cvxb.h"

static unsigned long long g_seed = 88172645463325252ull;
static double urand(double lo, double hi) {
  g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17;
  return lo + (hi - lo) * (double)(g_seed >> 11) / 9007199254740992.0;
}
#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "drive_abi FAILED: " __VA_ARGS__); fprintf(stderr, " [%s]\n", cvxb_last_error()); return 1; } } while (0)

int main(void) {
  cvxb_handle h = 0;
  int st = cvxb_create(0, 0, 0, &h);
  CHECK(st == CVXB_OK, "cvxb_create -> %d", st);

  /* ---- planted KKT system: H = L L' (L lower, diagonal boosted), A = U(-5,5) + 20 I, q = -(Hx + A'w), b = Ax ---- */
  const int n = 96, p = 12;
  double *L = calloc((size_t)n * n, 8), *H = calloc((size_t)n * n, 8), *A = calloc((size_t)p * n, 8);
  double *x0 = calloc(n, 8), *w0 = calloc(p, 8), *q = calloc(n, 8), *b = calloc(p, 8), *x = calloc(n, 8), *w = calloc(p, 8);
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) L[(size_t)j * n + i] = urand(-5, 5) + (i == j ? sqrt((double)n) + 5.0 : 0.0);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int k = 0; k <= (i < j ? i : j); ++k) s += L[(size_t)k * n + i] * L[(size_t)k * n + j];
      H[(size_t)j * n + i] = s;
    }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < p; ++i) A[(size_t)j * p + i] = urand(-5, 5) + (i == j ? 20.0 : 0.0);
  for (int j = 0; j < n; ++j) x0[j] = urand(-1, 1);
  for (int i = 0; i < p; ++i) w0[i] = urand(-2, 2);
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += H[(size_t)j * n + i] * x0[j];
    for (int k = 0; k < p; ++k) s += A[(size_t)i * p + k] * w0[k];
    q[i] = -s;
  }
  for (int k = 0; k < p; ++k) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += A[(size_t)j * p + k] * x0[j];
    b[k] = s;
  }
  cvxb_kkt_info info;
  memset(&info, 0, sizeof info);
  st = cvxb_kkt_solve(h, n, p, H, n, A, p, q, b, 1e-10, x, w, &info);
  CHECK(st == CVXB_OK, "cvxb_kkt_solve -> %d", st);
  double ex = 0, nx = 0, ew = 0, nw = 0;
  for (int j = 0; j < n; ++j) { ex += (x[j] - x0[j]) * (x[j] - x0[j]); nx += x0[j] * x0[j]; }
  for (int k = 0; k < p; ++k) { ew += (w[k] - w0[k]) * (w[k] - w0[k]); nw += w0[k] * w0[k]; }
  CHECK(sqrt(ex / nx) < 1e-9 && sqrt(ew / nw) < 1e-9, "planted KKT solution: rel err x %.3g w %.3g", sqrt(ex / nx), sqrt(ew / nw));
  printf("kkt_solve: path %d, rel err x %.2e w %.2e, err1 %.2e err2 %.2e\n", info.path, sqrt(ex / nx), sqrt(ew / nw), info.err1, info.err2);

  /* ---- error convention: an indefinite matrix must come back as CVXB_ELINSOLVE with a message ---- */
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) H[(size_t)j * n + i] = (i == j) ? -1.0 : 0.0;
  st = cvxb_cholesky_solve(h, n, H, n, q, 1e-10, x, 0);
  CHECK(st == CVXB_ELINSOLVE && strlen(cvxb_last_error()) > 0, "choleskySolve(-I) -> %d", st);

  /* ---- known-answer LP through seam A: min -a'x s.t. |x_j| <= |a_j|, optimum x = a, phase I from x = 2a ---- */
  const int d = 24, m = 2 * d;
  double *a = calloc(d, 8), *c = calloc(d, 8), *G = calloc((size_t)m * d, 8), *ub = calloc(m, 8), *xd = calloc(d, 8), *xs = calloc(d, 8);
  for (int j = 0; j < d; ++j) {
    a[j] = urand(0.5, 2.0) * (j % 2 ? -1.0 : 1.0);
    c[j] = -a[j];
    G[(size_t)j * m + 2 * j] = 1.0; G[(size_t)j * m + 2 * j + 1] = -1.0;
    ub[2 * j] = ub[2 * j + 1] = fabs(a[j]);
    xd[j] = 2.0 * a[j];
  }
  cvxb_problem_desc D;
  memset(&D, 0, sizeof D);
  D.n = d; D.m = m; D.p = 0; D.objective = CVXB_OBJ_LINEAR; D.obj_a = c; D.G = G; D.ldg = m; D.ub = ub; D.x_defined = xd;
  cvxb_problem prob = 0;
  st = cvxb_problem_create(h, &D, &prob);
  CHECK(st == CVXB_OK, "cvxb_problem_create -> %d", st);
  cvxb_solution sol;
  memset(&sol, 0, sizeof sol);
  sol.x = xs;
  st = cvxb_barrier_solve(h, prob, 0, &sol);
  CHECK(st == CVXB_OK, "cvxb_barrier_solve -> %d", st);
  double opt = 0, err = 0;
  for (int j = 0; j < d; ++j) { opt -= a[j] * a[j]; err = fmax(err, fabs(xs[j] - a[j])); }
  CHECK(fabs(sol.objective - opt) < 1e-6 * fabs(opt) && err < 1e-6, "minDotProduct: objective %.12g vs %.12g, max |x - a| %.3g", sol.objective, opt, err);
  CHECK(sol.phase1_newton_steps > 0 && sol.has_dualityGap && sol.dualityGap < 1e-8, "minDotProduct: phase I steps %lld gap %g", sol.phase1_newton_steps, sol.dualityGap);
  printf("barrier_solve: objective %.12f (optimum %.12f), %lld phase-I + %lld Newton steps, %d stages, %.2f ms on the device\n",
         sol.objective, opt, sol.phase1_newton_steps, sol.newton_steps, sol.outer_stages, sol.solve_ms);
  cvxb_problem_destroy(prob);
  CHECK(cvxb_launch_count(h) > 0, "no kernel launches counted");
  cvxb_destroy(h);
  printf("drive_abi ok\n");
  return 0;
}

// Device-resident problems (seam A): closed-form objective / constraint families of SURVEY.md 8a row a7.
#pragma once
#include "kkt.cuh"

enum { CVXB_OBJ_COMPOSED = 100 };     // internal: not part of cvxb_objective_kind

struct cvxb_problem_s {
  cvxb_handle_s* h = nullptr;
  int n = 0, m = 0, p = 0, objective = 0;
  int ldm = 0, ldn = 0, ldp = 0;
  double obj_r = 0.0;
  double obj_pow = 2.0;      // CVXB_OBJ_PNORM exponent
  // problem data (device)
  double *G = nullptr, *gr = nullptr, *ub = nullptr, *A = nullptr, *b = nullptr, *obj_a = nullptr, *obj_P = nullptr;
  double *x_feas = nullptr, *x_def = nullptr;
  bool has_feasible = false;
  // iterate + work vectors (device)
  double *x = nullptr, *gx = nullptr, *inv = nullptr, *Gd = nullptr, *y = nullptr, *gt = nullptr, *dir = nullptr,
         *nu = nullptr, *eqdiff = nullptr, *Px = nullptr, *Pd = nullptr, *axv = nullptr;
  // primal-dual work (allocated on first use)
  double *lam = nullptr, *dlam = nullptr, *dnu = nullptr, *wts = nullptr, *rd0 = nullptr, *rd1 = nullptr,
         *Adx = nullptr, *pres = nullptr, *x0s = nullptr, *lam0s = nullptr, *nu0s = nullptr, *tmpm = nullptr,
         *gx0s = nullptr, *rd00s = nullptr, *Px0s = nullptr, *vvec = nullptr, *qvec = nullptr, *atnu = nullptr,
         *pres0s = nullptr, *negpres = nullptr;
  // quadratic constraints: m = mlin + mq rows in every per-constraint vector; rows mlin.. of G hold the
  // current gradients a_k + P_k x (rewritten at every evaluation)
  int mlin = 0, mq = 0, ldq = 0;
  double *Pq = nullptr;      // (mq*ldq) x n stacked: block k = P_k
  double *qa = nullptr;      // ldq x mq
  double *PX = nullptr, *PDv = nullptr;   // mq*ldq: P_k x and P_k d
  double *qcorr = nullptr, *qq = nullptr; // m: x'P_k x / 2 and d'P_k d / 2 on the quadratic rows, 0 elsewhere
  double *rd2 = nullptr;     // n: PD line search, s^2 coefficient of the dual residual
  // CVXB_OBJ_KLDUAL: f(z) = w'z + sum_j R_j exp(-(B'z)_j); obj_P holds B (n x kd), obj_a holds w
  int kd = 0, ldk = 0;
  double *objR = nullptr, *du = nullptr, *dy = nullptr, *dv = nullptr, *Bs = nullptr;   // B'z, R o exp(-B'z), B'd, B diag(sqrt(t y))
  // objective composed with an affine map, f(u) = f_inner(z0 + F u) (ObjectiveFunction.affineTransformed,
  // ObjectiveFunction.scala:26-40, for the KL and p-norm objectives of a reduced problem): objective ==
  // CVXB_OBJ_COMPOSED, cmp_kind the inner family, cmp_n its dimension; F is cmp_n x n (leading dimension cmp_ld)
  int cmp_n = 0, cmp_ld = 0, cmp_kind = 0;
  double *cmpF = nullptr, *cmpz0 = nullptr;
  double *cmpx = nullptr;     // z0 + F u
  double *cmpd = nullptr;     // F du (line searches)
  double *cmpg = nullptr;     // grad f_inner(z0 + F u); scratch of the primal-dual line search afterwards
  double *cmpw = nullptr;     // t * diag hess f_inner
  double *cmpFs = nullptr;    // diag(sqrt(cmpw)) F
  double *cmpgf = nullptr;    // n: F' grad f_inner at a line-search trial point
  // matrices
  double *Gs = nullptr, *H = nullptr, *Hreg = nullptr;
  cvxb::KktWork kw;
  cudaGraphExec_t step_graph[2] = {nullptr, nullptr};   // captured Newton step: [0] with equalities, [1] without
  long long graph_launches[2] = {0, 0};
  double graph_flops[2] = {0.0, 0.0};
  cvxb_params graph_pars;
  // a whole centering stage as ONE graph launch: a WHILE node whose body is the Newton step followed by a kernel that
  // applies the reference's loop test on the device (EqualityConstrainedSolver.scala:49, UnconstrainedSolver.scala:45)
  cudaGraphExec_t loop_graph[2] = {nullptr, nullptr};
  long long loop_step_launches[2] = {0, 0};
  double loop_flops[2] = {0.0, 0.0};          // algorithmic flops of the timed SYRK of one step of the loop
  bool loop_failed = false;
  cvxb_problem_s* phase1 = nullptr;   // the n+1 dimensional feasibility problem (built on demand)
  std::vector<void*> owned;
  cvxb::Arena arena;
  bool arena_async = false;     // owned[0] came from cudaMallocAsync on the handle's stream
};

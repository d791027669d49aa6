// Device-resident problems (seam A): closed-form objective / constraint families of SURVEY.md 8a row a7.
#pragma once
#include "kkt.cuh"

struct cvxb_problem_s {
  cvxb_handle_s* h = nullptr;
  int n = 0, m = 0, p = 0, objective = 0;
  int ldm = 0, ldn = 0, ldp = 0;
  double obj_r = 0.0;
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
  // matrices
  double *Gs = nullptr, *H = nullptr, *Hreg = nullptr;
  cvxb::KktWork kw;
  cvxb_problem_s* phase1 = nullptr;   // the n+1 dimensional feasibility problem (built on demand)
  std::vector<void*> owned;
  cvxb::Arena arena;
};

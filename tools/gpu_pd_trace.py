import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Primal-dual solve step by step: oracle against device iterates after k = 1, 2, ... iterations (step budget).
usage: python tools/gpu_pd_trace.py n m_half p seed"""
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P
n, mh, p, seed = [int(a) for a in sys.argv[1:5]]
prob = P.slab_lp(n, mh, p, seed)
objF, cnts, eqs = P.to_oracle(prob)
h = cb.default_handle()
full = O.PrimalDual(objF, cnts, eqs, O.SolverParams.standardParams(), False, False).solve()
print("oracle: %d iterations, objective %.12g gap %.3g" % (full.newton_steps, objF.valueAt(full.x), full.dualityGap))
for k in range(1, full.newton_steps + 3):
    s0 = O.PrimalDual(objF, cnts, eqs, O.SolverParams.standardParams(), False, False).solve(max_steps=k)
    pars = cb.SolverParams()
    pars.stepLimit = k
    try:
        s1 = cb.from_dict(prob, "PD", pars, h).solve()
        dx = np.linalg.norm(s1.x - s0.x) / max(np.linalg.norm(s0.x), 1e-300)
        print("k=%2d  oracle trials %s gap %.3e |rdual| %.3e   device gap %.3e |rdual| %.3e  rel dx %.2e  steps %d" %
              (k, s0.linesearch_trials[-1:] , s0.dualityGap, s0.normDualResidual, s1.dualityGap, s1.normDualResidual, dx, s1.newton_steps), flush=True)
    except cb.CvxbError as e:
        print("k=%2d  oracle trials %s gap %.3e |rdual| %.3e   device %s: %s" % (k, s0.linesearch_trials[-1:], s0.dualityGap, s0.normDualResidual, type(e).__name__, str(e)[:90]), flush=True)

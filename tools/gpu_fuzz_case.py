import sys; sys.path.insert(0, "/root/repo")
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P
h = cb.default_handle()
cases = [(32, 56, 344636), (13, 24, 483169), (11, 20, 874895), (17, 26, 588447)]
for n, m, seed in cases:
    pr = P.slab_lp(n, m // 2, 0, seed)
    objF, cnts, eqs = P.to_oracle(pr)
    stats = []
    try:
        s0 = O.barrierSolve(objF, cnts, eqs, O.SolverParams.standardParams(), None, False)
        r0 = ("ok", objF.valueAt(s0.x), s0.outer_stages, s0.stage_newton_steps)
    except Exception as e:
        r0 = (type(e).__name__, str(e)[:80])
    try:
        s1 = cb.from_dict(pr, "BR", None, h).solve()
        r1 = ("ok", s1.objective, s1.outer_stages)
    except cb.CvxbError as e:
        r1 = (type(e).__name__, str(e)[:100])
    sol = cb.BatchedBarrierSolver(cb.pack_problems([pr]), None, h).solve()
    print(n, m, seed, "\n  oracle ", r0, "\n  large  ", r1, "\n  batched", int(sol.status[0]), int(sol.outer_stages[0]), sol.stage_newton_steps[0][:6], flush=True)

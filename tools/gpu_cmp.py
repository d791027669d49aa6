import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sys, time
from oracle import cvx_oracle as O, problems as P
import cvx_b200 as cb
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_barrier_gpu import PROBLEMS
PROBLEMS = dict(PROBLEMS)
PROBLEMS["slab_lp_100_eq"] = lambda: P.slab_lp(100, 100, 20, 0)
PROBLEMS["kl_random_200"] = lambda: P.kl_random(200, 200, 49, 0)
for name in sorted(PROBLEMS):
    prob = PROBLEMS[name]()
    objF, cnts, eqs = P.to_oracle(prob)
    t0 = time.time()
    try:
        sol0, ph0 = O.solveProblem(objF, cnts, eqs, "BR")
    except Exception as e:
        print(name, "oracle failed", repr(e)); continue
    t1 = time.time()
    try:
        sol = cb.from_dict(prob, "BR").solve()
    except Exception as e:
        print(name, "gpu failed", repr(e)); continue
    t2 = time.time()
    print(name, "oracle %.2fs gpu %.2fs (device %.1f ms)" % (t1 - t0, t2 - t1, sol.solve_ms))
    print("   stages gpu   ", sol.stage_newton_steps, "exec", sol.executed_newton_steps)
    print("   stages oracle", sol0.stage_newton_steps)
    if ph0 is not None:
        print("   ph1 gpu", sol.phase1_newton_steps, sol.phase1_stages, "oracle", ph0.stage_newton_steps, sum(ph0.stage_newton_steps))
    o0 = objF.valueAt(sol0.x)
    print("   obj gpu %.12g oracle %.12g rel %.2e  xrel %.2e" % (sol.objective, o0, abs(sol.objective - o0) / max(1, abs(o0)), np.linalg.norm(sol.x - sol0.x) / np.linalg.norm(sol0.x)))

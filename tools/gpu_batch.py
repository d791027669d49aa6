import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, sys
import synthetic as P
import cvx_b200 as cb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
t0=time.time(); probs = [P.batched_problem(i, 64, 128, 1000) for i in range(B)]; packed = cb.pack_problems(probs); print('gen+pack', time.time()-t0)
s = cb.BatchedBarrierSolver(packed)
for r in range(3):
    sol = s.solve()
    print('B', B, 'ms', sol.solve_ms, 'solves/s', B/(sol.solve_ms/1e3), 'ok', int((sol.status==0).sum()), 'steps mean', sol.newton_steps.mean(), 'max', sol.newton_steps.max(), 'steps/s', sol.newton_steps.sum()/(sol.solve_ms/1e3))
print(np.bincount(sol.status))

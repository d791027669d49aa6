import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Randomised differential test of the closure-objective path (seam B staged through pinned buffers; host Newton loops in
cvx_b200/generic.py) on random Type1Function power problems, unconstrained and with equalities, against the oracle's
loops fed with the same closures.
usage: python tools/gpu_fuzz_generic.py [cases] [seed]"""
import time
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P
from tests.test_generic_gpu import _ObjectiveOnly

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = cb.default_handle()
bad = 0
t0 = time.time()
for it in range(N):
    dim = int(rng.integers(3, 260))
    m = int(rng.integers(max(1, dim // 2), dim + 1))
    q = float(rng.choice([1.5, 2.0, 2.5, 3.0]))
    seed = int(rng.integers(0, 10**6))
    with_eq = bool(rng.integers(0, 2)) and dim > 6
    blk = int(rng.choice([16, 50, 64, 128]))
    f, x0 = P.random_power_problem(dim, m, q, seed)
    msgs = []
    try:
        if with_eq:
            p = int(rng.integers(1, max(2, dim // 6)))
            A = np.random.default_rng(seed + 1).uniform(-1, 1, (p, dim))
            b = A @ x0
            s1 = cb.generic.EqualityConstrainedSolver(f, A, b, x0, None, None, h, block_cols=blk).solve()
            s0 = O.equalityConstrainedSolve(_ObjectiveOnly(f), 1.0, x0, A, b, O.SolverParams.standardParams())
            if np.linalg.norm(A @ s1.x - b) > 1e-8 * max(1.0, np.linalg.norm(b)): msgs.append("Ax=b")
        else:
            s1 = cb.generic.UnconstrainedSolver(f, x0, None, None, h, block_cols=blk).solve()
            s0 = O.unconstrainedSolve(_ObjectiveOnly(f), 1.0, x0, O.SolverParams.standardParams())
        if abs(s1.newton_steps - s0.newton_steps) > 1: msgs.append("steps %d %d" % (s1.newton_steps, s0.newton_steps))
        if abs(s1.objective - f.valueAt(s0.x)) > 1e-8 * max(1.0, abs(s1.objective)): msgs.append("objective %.3e %.3e" % (s1.objective, f.valueAt(s0.x)))
    except Exception as e:
        msgs.append("%s %s" % (type(e).__name__, str(e)[:100]))
    if msgs:
        bad += 1
        print("CASE", it, "dim", dim, "m", m, "q", q, "seed", seed, "eq" if with_eq else "uncon", msgs, flush=True)
print("closure-objective fuzz: %d cases, %d disagreements, %.1f s" % (N, bad, time.time() - t0))

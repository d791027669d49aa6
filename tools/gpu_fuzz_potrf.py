import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Blocked Cholesky / KKT solve at sizes around every schedule switch (128-column leaves, look-ahead <= 4608, recursive
halving, tile-DAG >= 5120 with its graded blocks) against LAPACK: regularizedCholesky, choleskySolve and KKTSystem.solve.
usage: python tools/gpu_fuzz_potrf.py [cases] [seed]"""
import time
import numpy as np
import scipy.linalg as sla
import cvx_b200 as cb
from cvx_b200 import MatrixUtils, KKTSystem

N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = cb.default_handle()
rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
marks = [128, 256, 1024, 2048, 2560, 4096, 4608, 5120, 6144, 7168]
bad = 0
t0 = time.time()
for it in range(N):
    if it % 3 == 0:
        n = int(rng.integers(1, 7400))
    else:
        n = int(rng.choice(marks)) + int(rng.integers(-2, 3))
    n = max(1, n)
    k = 48
    M = rng.uniform(-1, 1, (n, k))
    H = M @ M.T + np.diag(rng.uniform(0.5, 2.0, n))
    H = (H + H.T) * 0.5
    L = MatrixUtils.regularizedCholesky(H, h)
    L0 = sla.cholesky(H, lower=True)
    e_l, e_r = rel(L, L0), rel(L @ L.T, H)
    p = int(rng.integers(1, min(n, 300))) if n > 1 else 0
    msgs = []
    if e_l > 1e-11 or e_r > 1e-13 or np.any(np.triu(L, 1) != 0.0): msgs.append("factor %.1e %.1e" % (e_l, e_r))
    x = rng.uniform(-1, 1, n)
    c = MatrixUtils.choleskySolve(H, H @ x, None, 1e-9, 0, h)
    if rel(c, x) > 1e-9: msgs.append("choleskySolve %.1e" % rel(c, x))
    if p:
        A = rng.uniform(-1, 1, (p, n))
        w = rng.uniform(-1, 1, p)
        q, b = -(H @ x + A.T @ w), A @ x
        x1, w1 = KKTSystem(H, A, q, b, h).solve(1e-6, None, 1e-8, 0)
        if rel(x1, x) > 1e-8 or rel(w1, w) > 1e-8: msgs.append("kkt %.1e %.1e" % (rel(x1, x), rel(w1, w)))
    print("n=%5d p=%4d  rel(L) %.1e  rel(LL') %.1e  %s" % (n, p, e_l, e_r, "  ".join(msgs) if msgs else "ok"), flush=True)
    bad += bool(msgs)
print("schedule-boundary fuzz: %d sizes, %d failures, %.1f s" % (N, bad, time.time() - t0))

import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Small end-to-end pass over every kernel family (for compute-sanitizer memcheck)."""
import numpy as np
import cvx_b200 as cb
import synthetic as P
h = cb.default_handle()
s = P.kkt_planted_pd(300, 40, 2)
x, w = cb.KKTSystem(s["H"], s["A"], s["q"], s["b"], h).solve(1e-6, None, 1e-7, 0)
print("kkt", np.linalg.norm(x - s["x"]))
T = np.tril(np.random.default_rng(0).uniform(-5, 5, (257, 257))) + 20 * np.eye(257)
print("trsm", cb.MatrixUtils.triangularSolve(T, "L", T @ np.ones((257, 3)), h)[:2, 0])
print("kl_1A", cb.from_dict(P.kl_1A(20), "BR", None, h).solve().objective)
print("qp pd", cb.from_dict(P.slab_qp(48, 60, 6, 3), "PD", None, h).solve().objective)
print("quad", cb.from_dict(P.lin_quad_set(12, 10, 5, 3, 3, "quadratic", False), "BR", None, h).solve().objective)
print("n=300", cb.from_dict(P.kl_random(300, 300, 49, 1), "BR", cb.SolverParams(stepLimit=12), h).solve().phase1_executed_steps)
bs = cb.BatchedBarrierSolver(cb.pack_problems([P.batched_problem(i, 64, 128, 1000) for i in range(6)]), None, h).solve()
print("batched", bs.status, bs.newton_steps)
H = -np.eye(12) * 50.0
A = np.random.default_rng(1).uniform(-1, 1, (3, 12))
print("eig", cb.KKTSystem(H, A, np.ones(12), np.ones(3), h).solve(1e-6, None, 1e-8, 0)[1])
# round 2 paths: reduced KL (composed objective), primal-dual on the dual KL objective, staged seam B with closures
pr = P.kl_random(24, 24, 11, 1); pr["x0"] = pr["qstar"].copy()
noeq = dict(pr); noeq["A"] = noeq["b"] = None
for st in ("BR", "PD"):
    red = cb.from_dict(noeq, st, None, h).solver.reduced(cb.SolutionSpace(pr["A"], pr["b"], h))
    print("reduced kl", st, red.solve().objective)
k1 = P.kl_1A(20)
print("dual pd", np.max(np.abs(cb.Dist_KL(20, k1["G"][:2], k1["ub"][:2], None, None, "BR", None, None, 0, h).solveDual("PD").x - k1["xopt"])))
f, x0 = P.random_power_problem(40, 30, 2.0, 0)
print("generic", cb.generic.UnconstrainedSolver(f, x0, None, None, h, block_cols=16).solve().objective)

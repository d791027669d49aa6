"""Dist_KL through duality (Duality.solveDual + the dual objective of Dist_KL.scala:107-163) on the device:
min -L_*(z) = w'z + R'exp(-B'z), lambda >= 0, then x = R o exp(-B'z).  The reference's own check is exactly
this: solve the KL problems via the dual and compare with the known minimiser (MinimizationTests.scala:28-83)."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_dual_newton_direction_parity(handle):
    import cvx_b200 as cb
    pr = P.kl_random(80, 50, 11, 3)
    H, u, A, r = pr["G"][:50], pr["ub"][:50], pr["A"][:11], pr["b"][:11]
    objF, cnts, mI = O.dist_KL_dual_problem(80, H, u, A, r)
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.uniform(0.01, 0.5, mI), rng.uniform(-0.3, 0.3, objF.dim - mI)])
    t = 30.0
    bf = O.BarrierFunctions(objF, cnts)
    H0, g0 = bf.hessian(t, z), bf.gradient(t, z)
    dp = cb.Dist_KL(80, H, u, A, r, "BR", None, None, 0, handle).dualProblem("BR")
    Hd, g, dz, _, info = dp.solver.newton_direction(z, t)
    assert np.array_equal(Hd, Hd.T)
    assert rel(Hd, H0) < 1e-13 and rel(g, g0) < 1e-13
    assert rel(dz, O.choleskySolve(H0, -g0, 0.1)) < 1e-9


@pytest.mark.parametrize("name", ["kl_1A", "kl_2A", "kl_random"])
def test_solve_dual_matches_oracle_and_known_minimiser(handle, name):
    import cvx_b200 as cb
    if name == "kl_1A":
        pr = P.kl_1A(20)
        n, H, u, A, r = 20, pr["G"][:2], pr["ub"][:2], None, None
    elif name == "kl_2A":
        pr = P.kl_2A(20)
        n, H, u, A, r = 20, None, None, pr["A"][:2], pr["b"][:2]
    else:
        pr = P.kl_random(120, 80, 19, 2)
        n, H, u, A, r = 120, pr["G"][:80], pr["ub"][:80], pr["A"][:19], pr["b"][:19]
    sol0 = O.solveDual(n, H, u, A, r)
    prob = cb.Dist_KL(n, H, u, A, r, "BR", None, None, 0, handle)
    sol = prob.solveDual("BR")
    assert sol.outer_stages == sol0.outer_stages
    for k in range(min(4, sol.outer_stages)):
        assert abs(sol.stage_newton_steps[k] - sol0.stage_newton_steps[k]) <= 1
    assert rel(sol.x, sol0.x) < 1e-7
    assert abs(sol.x.sum() - 1.0) < 1e-6 and np.all(sol.x > 0)      # one centering only when there are no inequalities
    if H is not None:
        assert np.all(sol.lam >= 0) and rel(sol.lam, sol0.lam) < 1e-5
        assert np.all(H @ sol.x <= u + 1e-7)
    if "xopt" in pr:
        assert np.max(np.abs(sol.x - pr["xopt"])) < 1e-5          # analytic minimiser (OptimizationProblems.scala:136-141,249-251)
    else:
        solp = prob.solve()                                        # primal route on the device: same optimum
        assert rel(sol.x, solp.x) < 1e-6


@pytest.mark.parametrize("name", ["kl_1A", "kl_random"])
def test_solve_dual_primal_dual_solver(handle, name):
    """Duality.solveDual with solverType "PD" (Duality.scala:99-133 passes the solver type through to
    OptimizationProblem; PrimalDualSolver.solve_noEQs, PrimalDualSolver.scala:381-460, on -L_*(z), lambda >= 0).  The
    residual line search needs grad f(z + s dz) = w - B (y o exp(-s B'dz)) at every trial point: one matrix-vector
    product per trial inside the line-search kernel.  Against the oracle's PrimalDual on the same dual problem."""
    import cvx_b200 as cb
    if name == "kl_1A":
        pr = P.kl_1A(20)
        n, H, u, A, r = 20, pr["G"][:2], pr["ub"][:2], None, None
    else:
        pr = P.kl_random(120, 80, 19, 2)
        n, H, u, A, r = 120, pr["G"][:80], pr["ub"][:80], pr["A"][:19], pr["b"][:19]
    objF, cnts, mI = O.dist_KL_dual_problem(n, H, u, A, r)
    s0 = O.PrimalDual(objF, cnts, None, O.SolverParams.standardParams()).solve()
    x0 = objF.primalOptimum(s0.x)
    prob = cb.Dist_KL(n, H, u, A, r, "BR", None, None, 0, handle)
    sol = prob.solveDual("PD")
    assert abs(sol.newton_steps - s0.newton_steps) <= 1
    assert rel(sol.z, s0.x) < 1e-6 and rel(sol.x, x0) < 1e-7
    assert sol.dualityGap < 1e-8 and sol.normDualResidual < 1e-8
    assert abs(sol.objective - objF.valueAt(s0.x)) < 1e-8 * max(1.0, abs(sol.objective))
    assert np.all(sol.lam >= 0) and np.all(H @ sol.x <= u + 1e-6)
    solb = prob.solveDual("BR")                                     # the barrier route reaches the same primal optimum
    assert rel(sol.x, solb.x) < 1e-5
    if "xopt" in pr:
        assert np.max(np.abs(sol.x - pr["xopt"])) < 1e-5

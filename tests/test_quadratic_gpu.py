"""Quadratic constraints r + a'x + x'Px/2 <= ub (QuadraticConstraint.scala) on the device: the barrier Hessian
gains hess g_k / d_k, gradients a_k + P_k x are rows of the constraint Jacobian, the line search is closed
form along the ray; phase I wraps them as Constraint.phase_I does.  Test design: FeasibilityTests.scala:105-117
(random sets of 10 linear + 5 quadratic constraints around a feasible point, +- 3 random equalities)."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n,ml,mq,p,seed", [(12, 10, 5, 0, 0), (12, 10, 5, 3, 1), (60, 40, 7, 8, 2), (30, 0, 4, 0, 3)])
def test_newton_direction_with_quadratic_constraints(handle, n, ml, mq, p, seed):
    import cvx_b200 as cb
    prob = P.lin_quad_set(n, ml, mq, p, seed)
    objF, cnts, eqs = P.to_oracle(prob)
    rng = np.random.default_rng(seed)
    x = prob["x0"] + 0.05 * rng.normal(size=n)
    assert cnts.isSatisfiedStrictlyBy(x)
    t = 7.0
    bf = O.BarrierFunctions(objF, cnts)
    H0, g0 = bf.hessian(t, x), bf.gradient(t, x)
    op = cb.from_dict(prob, "BR", None, handle)
    H, g, dx, nu, info = op.solver.newton_direction(x, t)
    assert rel(H, H0) < 1e-13 and rel(g, g0) < 1e-13
    if p:
        dx0, nu0 = O.kkt_solve(H0, prob["A"], g0, prob["b"] - prob["A"] @ x, 0.1)
        assert rel(dx, dx0) < 1e-9 and rel(nu, nu0) < 1e-9
    else:
        assert rel(dx, O.choleskySolve(H0, -g0, 0.1)) < 1e-9


@pytest.mark.parametrize("n,ml,mq,p,seed,feasible", [(12, 10, 5, 0, 0, True), (12, 10, 5, 3, 1, True), (12, 10, 5, 0, 2, False),
                                                     (12, 10, 5, 3, 3, False), (40, 30, 6, 5, 4, True)])
@pytest.mark.parametrize("solver", ["BR", "PD"])
def test_solve_with_quadratic_constraints(handle, n, ml, mq, p, seed, feasible, solver):
    import cvx_b200 as cb
    prob = P.lin_quad_set(n, ml, mq, p, seed, "quadratic", feasible)
    objF, cnts, eqs = P.to_oracle(prob)
    sol0, ph0 = O.solveProblem(objF, cnts, eqs, solver)
    sol = cb.from_dict(prob, solver, None, handle).solve()
    o0 = objF.valueAt(sol0.x)
    assert abs(sol.objective - o0) <= 1e-8 * max(1.0, abs(o0))
    assert rel(sol.x, sol0.x) < 1e-6
    assert cnts.isSatisfiedStrictlyBy(sol.x)
    if solver == "BR":
        assert sol.outer_stages == sol0.outer_stages
        for k in range(min(4, sol.outer_stages)):
            assert abs(sol.stage_newton_steps[k] - sol0.stage_newton_steps[k]) <= 1
    else:
        assert abs(sol.newton_steps - sol0.newton_steps) <= 1
        assert sol.lam.shape[0] == ml + mq and np.all(sol.lam > 0)
    if ph0 is not None:
        assert sol.phase1_stages == ph0.outer_stages

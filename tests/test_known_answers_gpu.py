"""More of the reference's KnownMinimizer problems (SimpleOptimizationProblems.standardProblems,
src/test/scala/cvx/SimpleOptimizationProblems.scala:579-600) through the device path: the analytic minimiser
(the reference's own acceptance test, Runner.scala:30, tolerance 1e-2 there) and agreement with the CPU oracle."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu

CASES = {
    "rank_one_simplex_10": lambda: P.rank_one_simplex(10),          # rank-one objective Hessian, simplex, phase I
    "jopt_p1_6": lambda: P.jopt_p1(6),                              # linear objective, one quadratic constraint, phase I
    "jopt_p2": lambda: P.jopt_p2(),                                 # docs/OptimizerExamples.pdf example 1.5
    "probability_simplex_8": lambda: P.probability_simplex_problem(8),   # a continuum of minimisers: objective only
    "distance_from_origin0_5": lambda: P.distance_from_origin(5),        # quadratic constraint only, infeasible start
    "distance_from_origin1_5": lambda: P.distance_from_origin(5, True),  # + 2n (duplicated) linear cuts active at x*
}


@pytest.mark.parametrize("solver", ["BR", "PD"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_known_minimisers(handle, name, solver):
    import cvx_b200 as cb
    prob = CASES[name]()
    objF, cnts, eqs = P.to_oracle(prob)
    sol0, _ = O.solveProblem(objF, cnts, eqs, solver)
    sol = cb.from_dict(prob, solver, None, handle).solve()
    f_opt = objF.valueAt(prob["xopt"])
    assert abs(objF.valueAt(sol.x) - f_opt) < 1e-6
    assert abs(sol.objective - objF.valueAt(sol0.x)) <= 1e-8 * max(1.0, abs(f_opt))
    if not name.startswith("probability_simplex"):
        assert np.max(np.abs(sol.x - prob["xopt"])) < 1e-3
    if eqs is not None:
        assert np.linalg.norm(eqs.A @ sol.x - eqs.b) < 1e-8


def test_free_variables_barrier(handle):
    """normSquaredWithFreeVariables (:308-340): one constraint, n-1 variables it does not depend on -- the phase-I
    barrier Hessian has rank one, so the plain and regularised Cholesky attempts fail and the reference's
    decomposition fallback takes over."""
    import cvx_b200 as cb
    prob = P.norm_squared_free_variables(8)
    objF, cnts, eqs = P.to_oracle(prob)
    sol0, _ = O.solveProblem(objF, cnts, eqs, "BR")
    sol = cb.from_dict(prob, "BR", None, handle).solve()
    assert abs(objF.valueAt(sol.x) - 0.5) < 1e-6
    assert abs(sol.objective - objF.valueAt(sol0.x)) <= 1e-8
    assert np.max(np.abs(sol.x - prob["xopt"])) < 1e-3

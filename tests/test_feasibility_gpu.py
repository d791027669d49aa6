"""Feasibility analysis drivers on the device: basic phase I with a FeasibilityReport (ConstraintSet.scala:355-414),
sum-of-infeasibilities phase I (ConstraintSet.scala:233-282, 511-545; Constraint.scala:101-159) and
withFeasiblePoint (:556-575).  Test design: FeasibilityTests.scala:22-48 (simple analysis, then SOI, on the same
set) over the probability simplex / random linear + quadratic sets / an infeasible KL set."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def mirror_set(cb, prob):
    quad = [cb.QuadraticConstraint("q%d" % k, prob["n"], q["ub"], q["r"], q["a"], q["P"]) for k, q in enumerate(prob.get("quad") or [])]
    cnts = cb.ConstraintSet(prob["G"], prob["ub"], prob["xdef"], prob.get("rvec"), quad)
    eqs = cb.EqualityConstraint(prob["A"], prob["b"]) if prob.get("A") is not None else None
    return cnts, eqs


CASES = [("slab", lambda: P.slab_lp(12, 12, 2, seed=3, feasible_start=False)),
         ("slab_noeq", lambda: P.slab_lp(20, 25, 0, seed=4, feasible_start=False)),
         ("kl", lambda: P.kl_random(10, 5, 2, seed=1)),
         ("quad", lambda: P.lin_quad_set(8, 6, 2, 1, 5, "quadratic", False)),
         ("quad_noeq", lambda: P.lin_quad_set(14, 9, 3, 0, 6, "quadratic", False)),
         ("infeasible", lambda: P.infeasible_kl_1(8))]


@pytest.mark.parametrize("name,make", CASES)
def test_constraint_values(handle, name, make):
    import cvx_b200 as cb
    prob = make()
    _, cnts0, _ = P.to_oracle(prob)
    cnts, _ = mirror_set(cb, prob)
    rng = np.random.default_rng(0)
    for x in (prob["xdef"], prob["xdef"] + 0.1 * rng.normal(size=prob["n"])):
        g = cnts.valuesAt(x, handle)
        g0 = cnts0.valuesAt(x)
        assert np.allclose(g, g0, rtol=1e-13, atol=1e-14)
        assert cnts.isSatisfiedStrictlyBy(x, handle) == cnts0.isSatisfiedStrictlyBy(x)


@pytest.mark.parametrize("name,make", CASES)
def test_phase_I_SOI(handle, name, make):
    import cvx_b200 as cb
    prob = make()
    _, cnts0, eqs0 = P.to_oracle(prob)
    rep0, sol0 = O.phase_I_Analysis_SOI(cnts0, eqs0, O.SolverParams())
    cnts, eqs = mirror_set(cb, prob)
    rep = cnts.phase_I_Analysis_SOI(eqs, None, 0, handle)
    n, p = cnts.dim, cnts.numConstraints
    assert rep.s.shape == (p,) and rep.x0.shape == (n,)
    assert np.all(rep.s > 0)                                      # inside the barrier's domain, hence D9 => never "strict"
    assert rep.isStrictlyFeasible is False and rep0.isStrictlyFeasible is False
    assert rep.isFeasible(1e-9) == rep0.isFeasible(1e-9)
    assert abs(rep.s.sum() - rep0.s.sum()) <= 1e-8 * max(1.0, rep0.s.sum())      # the SOI optimum
    assert abs(rep.s.max() - rep0.s.max()) <= 1e-6 * max(1.0, rep0.s.max())
    assert rep.violatedConstraints(1e-9) == rep0.violatedConstraints(1e-9)
    if eqs is not None:
        assert rep.equalityConstraintError < 1e-9
    sol = rep.solution
    assert sol.outer_stages == sol0.outer_stages
    for k in range(min(3, sol.outer_stages)):
        assert abs(sol.stage_newton_steps[k] - sol0.stage_newton_steps[k]) <= 1
    if rep0.isFeasible(1e-9):
        # every original constraint holds up to the remaining infeasibility s_j
        g = cnts.valuesAt(rep.x0, handle)
        assert np.all(g <= cnts.ub_all() + rep.s + 1e-12)


@pytest.mark.parametrize("name,make", CASES)
def test_phase_I_report_and_withFeasiblePoint(handle, name, make):
    import cvx_b200 as cb
    prob = make()
    _, cnts0, eqs0 = P.to_oracle(prob)
    cnts, eqs = mirror_set(cb, prob)
    if name == "infeasible":
        # the reference's own phase I leaves the barrier's domain here (IllegalArgumentException) or ends with s > 0
        # (InfeasibleProblemException); either way no feasible point may come out
        with pytest.raises(Exception):
            O.withFeasiblePoint(cnts0, eqs0, O.SolverParams())
        with pytest.raises(cb.CvxbError):
            cnts.withFeasiblePoint(eqs, None, 0, handle)
        return
    x0, s0, sol0 = O.phase_I_Analysis(cnts0, eqs0, O.SolverParams())
    rep = cnts.phase_I_Analysis(eqs, None, 0, handle)
    assert rep.s.shape == (1,)
    assert (rep.s[0] < 0) == (s0 < 0)
    assert rep.solution.outer_stages == sol0.outer_stages
    assert rep.isFeasible(1e-9)
    c2 = cnts.withFeasiblePoint(eqs, None, 0, handle)
    assert c2.feasiblePoint is not None and cnts0.isSatisfiedStrictlyBy(c2.feasiblePoint)
    assert c2.withFeasiblePoint(eqs, None, 0, handle) is c2


def test_probability_simplex_feasibility(handle):
    """FeasibilityTests.checkFeasibilityProbabilitySimplex (:54-68): x_j >= 0, sum x = 1 from x = 1/n."""
    import cvx_b200 as cb
    n = 10
    cnts = cb.ConstraintSet(-np.eye(n), np.zeros(n), np.full(n, 1.0 / n))
    eqs = cb.EqualityConstraint(np.ones((1, n)), np.array([1.0]))
    rep = cnts.phase_I_Analysis(eqs, None, 0, handle)
    assert rep.isFeasible(1e-9) and np.all(rep.x0 > 0) and abs(rep.x0.sum() - 1) < 2e-6
    soi = cnts.phase_I_Analysis_SOI(eqs, None, 0, handle)
    assert soi.isFeasible(1e-9) and abs(soi.x0.sum() - 1) < 1e-9 and np.all(soi.x0 > -1e-9)

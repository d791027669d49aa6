"""Primal-dual solver (PrimalDualSolver.scala) on the device against the CPU oracle: per-step direction
parity and whole-solve parity.  With equalities the literal reference is defective:
full-solve parity is asserted against the corrected oracle, and the bugCompat switch against the
bug-compatible oracle."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _start(prob):
    objF, cnts, eqs = P.to_oracle(prob)
    pd = O.PrimalDual(objF, cnts, eqs, O.SolverParams())
    x = np.array(cnts.feasiblePoint)
    lam = cnts.lambda0(x)
    nu = np.zeros(eqs.A.shape[0]) if eqs is not None else None
    t = 10.0 * cnts.numConstraints / pd.surrogateDualityGap(x, lam)
    return objF, cnts, eqs, pd, x, lam, nu, t


@pytest.mark.parametrize("maker", [lambda: P.slab_qp(64, 64, 0, 1), lambda: P.slab_qp(96, 100, 12, 2),
                                   lambda: P.kl_small(64, 64, 2), lambda: P.slab_qp(300, 300, 40, 3)])
def test_pd_direction_parity(handle, maker):
    import cvx_b200 as cb
    prob = maker()
    objF, cnts, eqs, pd, x, lam, nu, t = _start(prob)
    H0 = pd.kktMatrix_noEqs(x, lam)
    dx0, dlam0, dnu0 = pd.newton_direction(t, x, lam, nu)
    op = cb.from_dict(prob, "PD", None, handle)
    H, dx, dlam, dnu, info = op.solver.newton_direction(x, lam, nu, t)
    assert np.array_equal(H, H.T)
    assert rel(H, H0) < 1e-13
    assert rel(dx, dx0) < 1e-8 and rel(dlam, dlam0) < 1e-8
    if nu is not None:
        assert rel(dnu, dnu0) < 1e-8
        A = prob["A"]
        v = pd.rhs1(t, x)
        res = np.linalg.norm(np.concatenate([H0 @ dx + A.T @ dnu - (v - A.T @ nu), A @ dx + (A @ x - prob["b"])]))
        assert res / np.linalg.norm(v) < 1e-10
    else:
        assert np.linalg.norm(H0 @ dx - pd.rhs1(t, x)) / np.linalg.norm(pd.rhs1(t, x)) < 1e-10


PD_PROBLEMS = {
    "slab_qp_noeq": lambda: P.slab_qp(64, 64, 0, 1),
    "slab_qp_eq": lambda: P.slab_qp(64, 64, 8, 1),
    "kl_small": lambda: P.kl_small(64, 64, 2),
    "slab_lp_noeq": lambda: P.slab_lp(60, 80, 0, 4),
    "kl_random_phase1": lambda: P.kl_random(60, 60, 9, 1),
    "kl_1A_phase1": lambda: P.kl_1A(20),
}


@pytest.mark.parametrize("name", sorted(PD_PROBLEMS))
def test_pd_solve_matches_oracle(handle, name):
    import cvx_b200 as cb
    prob = PD_PROBLEMS[name]()
    objF, cnts, eqs = P.to_oracle(prob)
    sol0, _ = O.solveProblem(objF, cnts, eqs, "PD")
    sol = cb.from_dict(prob, "PD", None, handle).solve()
    assert abs(sol.newton_steps - sol0.newton_steps) <= 1
    obj0 = objF.valueAt(sol0.x)
    assert abs(sol.objective - obj0) <= 1e-8 * max(1.0, abs(obj0))
    assert sol.dualityGap < 1e-8 and sol.normDualResidual < 1e-8
    assert rel(sol.x, sol0.x) < 1e-6
    assert np.all(sol.lam > 0)
    assert cnts.isSatisfiedStrictlyBy(sol.x)
    if eqs is not None:
        assert sol.nu is not None and sol.equalityGap < 1e-8
    else:
        assert sol.nu is None and sol.equalityGap is None


def test_pd_bug_compat_matches_literal_reference(handle):
    """solve_withEQs as written (defects D1 + D2): oracle(bug_compat) and device(bugCompat) must agree on
    the outcome -- either the same failure class or the same iteration count."""
    import cvx_b200 as cb
    prob = P.slab_qp(32, 40, 4, 5)
    objF, cnts, eqs = P.to_oracle(prob)
    try:
        sol0, _ = O.solveProblem(objF, cnts, eqs, "PD", bug_compat=True)
        out0 = ("ok", sol0.newton_steps)
    except O.LineSearchFailedException:
        out0 = ("linesearch", None)
    try:
        sol = cb.from_dict(prob, "PD", cb.SolverParams(bugCompat=True), handle).solve()
        out = ("ok", sol.newton_steps)
    except cb.LineSearchFailedException:
        out = ("linesearch", None)
    assert out[0] == out0[0]
    if out[0] == "ok":
        assert abs(out[1] - out0[1]) <= 1
        assert sol.maxedOut == sol0.maxedOut


def test_pd_line_search_failure_is_reported(handle):
    """slab LP with equalities: the residual-decrease line search of the reference fails on this problem
    (LineSearchFailedException, PrimalDualSolver.scala:538-541); the device path reports the same error."""
    import cvx_b200 as cb
    prob = P.slab_lp(60, 80, 10, 4)
    objF, cnts, eqs = P.to_oracle(prob)
    with pytest.raises(O.LineSearchFailedException):
        O.solveProblem(objF, cnts, eqs, "PD")
    with pytest.raises(cb.LineSearchFailedException):
        cb.from_dict(prob, "PD", None, handle).solve()

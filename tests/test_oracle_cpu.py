"""CPU-only checks of the oracle (the checker of the CUDA path): the reference's own test designs with
seeds added (KktTest.scala, MatrixUtilsTests.scala), the analytic optima of its known-answer problems,
and the committed golden fixtures (tests/golden/oracle_golden.json, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def test_ruiz_zero_row():
    """MatrixUtilsTests.scala:16-26."""
    A = np.array([[.5, 0, .5], [0, 0, 0], [-.5, 0, -.5]])
    d, Q = O.ruizEquilibrate(A)
    assert d[1] == 1.0
    assert np.allclose(np.linalg.norm(Q[[0, 2]], axis=1), 1.0, atol=1e-5)


def test_ruiz_reduces_condition_number():
    """MatrixUtilsTests.scala:384-404."""
    rng = np.random.default_rng(0)
    M = rng.uniform(-1, 1, (60, 60))
    H = M @ M.T + 0.06 * np.eye(60)
    s = 10.0 ** rng.uniform(-3, 3, 60)
    H = H * np.outer(s, s)
    d, Q = O.ruizEquilibrate(H)
    assert np.linalg.cond(Q) < 1e-3 * np.linalg.cond(H)
    assert np.array_equal(Q, Q.T)


@pytest.mark.parametrize("n,p", [(5, 1), (200, 7)])
def test_triangular_solves_planted(n, p):
    """MatrixUtilsTests.testTriangularSolve / testForwardSolve / testBackSolve (:36-158)."""
    rng = np.random.default_rng(n)
    L = np.tril(rng.uniform(-5, 5, (n, n))) + 20 * np.eye(n)
    X = rng.uniform(0, 1, (n, p))
    assert rel(O.triangularSolve(L, "L", L @ X), X) < 1e-12
    assert rel(O.triangularSolve(L.T, "U", L.T @ X), X) < 1e-12
    assert rel(O.forwardSolve(L, L @ X[:, 0]), X[:, 0]) < 1e-12
    assert rel(O.backSolve(L.T, L.T @ X[:, 0]), X[:, 0]) < 1e-12


def test_forward_solve_zero_diagonal_asserts():
    L = np.tril(np.ones((4, 4)))
    L[2, 2] = 0
    with pytest.raises(AssertionError):
        O.forwardSolve(L, np.ones(4))


def test_cholesky_solve_and_regularisation():
    rng = np.random.default_rng(1)
    A = rng.uniform(-1, 1, (80, 80))
    H = A.T @ A + (A.T @ A).T              # testSolveWithPreconditioning :165-198
    x = rng.uniform(-1, 1, 80)
    assert rel(O.choleskySolve(H, H @ x, 1e-9), x) < 1e-6
    B = rng.uniform(-1, 1, (30, 6))
    Q = B @ B.T
    Q = (Q + Q.T) / 2                        # rank deficient -> regularised branch
    L = O.regularizedCholesky(Q)
    assert O.regularizedCholesky.last_regularized
    assert rel(L @ L.T, Q + 1e-10 * np.eye(30)) < 1e-9
    with pytest.raises(Exception):
        O.choleskySolve(np.diag([1.0, -1.0]), np.ones(2), 0.1)


@pytest.mark.parametrize("n,p,seed", [(10, 2, 0), (100, 20, 1), (400, 60, 2)])
def test_kkt_planted(n, p, seed):
    """KktTest.testPositiveDefinite (:197-272) and testSolutionWithCholFactor (:117-184)."""
    s = P.kkt_planted_pd(n, p, seed)
    x, w = O.kkt_solve(s["H"], s["A"], s["q"], s["b"], 1e-7)
    assert rel(x, s["x"]) < 1e-6 and rel(w, s["w"]) < 1e-6
    s = P.kkt_planted_chol(n, p, seed)
    x, w = O.solveWithCholFactor(s["L"], s["A"], s["q"], s["b"], 1e-7)
    assert rel(x, s["x"]) < 1e-6 and rel(w, s["w"]) < 1e-6


def test_kkt_fallback_chain():
    rng = np.random.default_rng(11)
    n, p = 40, 6
    A = rng.uniform(-1, 1, (p, n))
    Qm, _ = np.linalg.qr(A.T, mode="complete")
    N = Qm[:, p:]
    H = N @ N.T
    H = (H + H.T) / 2
    x, w = rng.uniform(-1, 1, n), rng.uniform(-1, 1, p)
    info = O.KKTInfo()
    x1, w1 = O.kkt_solve(H, A, -(H @ x + A.T @ w), A @ x, 1e-6, info)
    assert info.path >= 1
    assert rel(x1, x) < 1e-5


def test_breeze_cholesky_requires_exact_symmetry():
    H = np.array([[2.0, 1.0], [1.0 + 1e-15, 2.0]])
    with pytest.raises(O.MatrixNotSymmetricException):
        O.breeze_cholesky(H)


@pytest.mark.parametrize("key", sorted(GOLD))
def test_golden(key):
    rec = GOLD[key]
    if key.startswith("kkt_planted_pd"):
        _, n, p = key.rsplit("_", 2)
        s = P.kkt_planted_pd(int(n), int(p), {"10": 0, "100": 1}[n])
        info = O.KKTInfo()
        x, w = O.kkt_solve(s["H"], s["A"], s["q"], s["b"], 1e-7, info)
        assert info.path == rec["path"]
        assert rel(x, rec["x"]) < 1e-9 and rel(w, rec["w"]) < 1e-9
        return
    from tests.golden.make_golden import CASES
    name, solver = key.split(":")
    prob = CASES[name]()
    objF, cnts, eqs = P.to_oracle(prob)
    if "raises" in rec:
        with pytest.raises(AssertionError):
            O.solveProblem(objF, cnts, eqs, solver)
        return
    sol, ph1 = O.solveProblem(objF, cnts, eqs, solver)
    assert abs(objF.valueAt(sol.x) - rec["objective"]) <= 1e-10 * max(1, abs(rec["objective"]))
    assert rel(sol.x, rec["x"]) < 1e-7
    assert sol.outer_stages == rec["outer_stages"]
    if "analytic_objective" in rec:         # KnownMinimizer check, Runner.scala:30 uses tol 1e-2
        assert abs(objF.valueAt(sol.x) - rec["analytic_objective"]) < 1e-6


def test_infeasible_problem_never_yields_a_point():
    objF, cnts, eqs = P.to_oracle(P.infeasible_kl_1(20))
    with pytest.raises(Exception):
        O.solveProblem(objF, cnts, eqs, "BR")


def test_literal_and_vectorised_hessian_agree():
    prob, x, t = P.newton_step_inputs(30, 40, 4, 0)
    objF, cnts, eqs = P.to_oracle(prob)
    H1 = O.BarrierFunctions(objF, cnts, literal=True).hessian(t, x)
    H2 = O.BarrierFunctions(objF, cnts, literal=False).hessian(t, x)
    assert rel(H2, H1) < 1e-13
    assert np.array_equal(H1, H1.T) and np.array_equal(H2, H2.T)


def test_pd_bug_compat_does_not_converge():
    """solve_withEQs as written never leaves the initial iterate's neighbourhood (defects D1, D2;
    docs/Log.txt 2018-02-12)."""
    objF, cnts, eqs = P.to_oracle(P.slab_qp(32, 40, 4, 5))
    sol, _ = O.solveProblem(objF, cnts, eqs, "PD", bug_compat=True)
    assert sol.maxedOut and sol.dualityGap > 1e-3
    sol2, _ = O.solveProblem(objF, cnts, eqs, "PD", bug_compat=False)
    assert not sol2.maxedOut and sol2.dualityGap < 1e-8


def test_dual_route_recovers_known_minimisers():
    """MinimizationTests.scala:28-83: the KL problems solved via Duality.solveDual reach the analytic optimum."""
    pr = P.kl_1A(20)
    sol = O.solveDual(20, pr["G"][:2], pr["ub"][:2], None, None)
    assert np.max(np.abs(sol.x - pr["xopt"])) < 1e-8
    assert np.all(sol.lam >= 0)
    pr = P.kl_random(40, 30, 5, 1)
    sold = O.solveDual(40, pr["G"][:30], pr["ub"][:30], pr["A"][:5], pr["b"][:5])
    objF, c, e = P.to_oracle(pr)
    solp, _ = O.solveProblem(objF, c, e, "BR")
    assert rel(sold.x, solp.x) < 1e-7


def test_oracle_kkt_against_extended_precision():
    """Independent of LAPACK: the oracle's KKT solution against a 50-digit mpmath solve of the same system."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    s = P.kkt_planted_pd(12, 3, 7)
    x, w = O.kkt_solve(s["H"], s["A"], s["q"], s["b"], 1e-9)
    n, p = 12, 3
    M = mp.zeros(n + p, n + p)
    for i in range(n):
        for j in range(n):
            M[i, j] = mp.mpf(float(s["H"][i, j]))
    for i in range(p):
        for j in range(n):
            M[n + i, j] = M[j, n + i] = mp.mpf(float(s["A"][i, j]))
    rhs = mp.matrix([mp.mpf(float(-v)) for v in s["q"]] + [mp.mpf(float(v)) for v in s["b"]])
    sol = mp.lu_solve(M, rhs)
    ref = np.array([float(sol[i]) for i in range(n + p)])
    got = np.concatenate([x, w])
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10


def test_oracle_lp_against_highs():
    """Whole-solve answer against an independent solver: scipy's HiGHS on the slab LP with equalities (C1 family)."""
    from scipy.optimize import linprog
    prob = P.slab_lp(30, 40, 5, 3)
    objF, cnts, eqs = P.to_oracle(prob)
    sol, _ = O.solveProblem(objF, cnts, eqs, "BR")
    res = linprog(prob["a"], A_ub=prob["G"], b_ub=prob["ub"], A_eq=prob["A"], b_eq=prob["b"], bounds=[(None, None)] * 30,
                  method="highs")
    assert res.status == 0
    assert abs(objF.valueAt(sol.x) - res.fun) < 1e-6 * max(1.0, abs(res.fun))


def test_oracle_qp_kkt_conditions():
    """QP with slab constraints: at the barrier solution the KKT conditions of the ORIGINAL problem hold to the
    duality-gap accuracy with multipliers lambda_i = 1/(t d_i) (B&V 11.2.2)."""
    prob = P.slab_qp(24, 30, 4, 2)
    objF, cnts, eqs = P.to_oracle(prob)
    sol, _ = O.solveProblem(objF, cnts, eqs, "BR")
    x = sol.x
    t = 10.0 ** (sol.outer_stages - 1)
    d = prob["ub"] - prob["G"] @ x
    lam = 1.0 / (t * d)
    g = prob["a"] + prob["P"] @ x + prob["G"].T @ lam
    nu = np.linalg.lstsq(prob["A"].T, -g, rcond=None)[0]
    assert np.linalg.norm(g + prob["A"].T @ nu) < 1e-5 * max(1.0, np.linalg.norm(prob["a"]))   # lambda = 1/(t d), d ~ 1e-10: rounding in d
    assert np.all(d > 0) and float(lam @ d) < 1e-7


def test_oracle_phase_I_SOI():
    """Sum-of-infeasibilities phase I (ConstraintSet.scala:511-545): at a feasible set the optimum is sum s = 0
    (reached up to the duality gap 2p/t), at the infeasible KL set some s_j stays bounded away from 0."""
    for prob, feasible in [(P.slab_lp(12, 12, 2, seed=3, feasible_start=False), True), (P.kl_random(10, 5, 2, seed=1), True),
                           (P.lin_quad_set(8, 6, 2, 1, 5, "quadratic", False), True), (P.infeasible_kl_1(8), False)]:
        objF, cnts, eqs = P.to_oracle(prob)
        rep, sol = O.phase_I_Analysis_SOI(cnts, eqs, O.SolverParams())
        assert rep.s.shape == (cnts.numConstraints,) and np.all(rep.s > 0) and not rep.isStrictlyFeasible
        assert rep.isFeasible(1e-9) == feasible
        assert (len(rep.violatedConstraints(1e-9)) == 0) == feasible
        if feasible:
            assert rep.s.sum() < 2 * cnts.numConstraints * 1e-9
            assert rep.equalityConstraintError < 1e-9

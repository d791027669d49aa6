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


# ---- the remaining test designs of the reference's MatrixUtilsTests / KktTest, with seeds added ------------------

@pytest.mark.parametrize("n,seed", [(10, 0), (40, 1), (120, 2)])
def test_solve_underdetermined(n, seed):
    """MatrixUtilsTests.testSolveUnderdetermined (:206-233): m = n/2 equations, A, b ~ U(0,1), +1 on A's diagonal;
    the solution space (z0, F) satisfies A z0 = b, A F = 0; F has orthonormal columns and z0 is orthogonal to them
    (z0 = Q1 R'^-1 b, F = Q2 of the QR of A', MatrixUtils.scala:536-550)."""
    rng = np.random.default_rng(seed)
    m = n // 2
    A = rng.uniform(0, 1, (m, n))
    A[np.arange(m), np.arange(m)] += 1.0
    b = rng.uniform(0, 1, m)
    z0, F = O.solveUnderdetermined(A, b)
    assert F.shape == (n, n - m)
    assert np.linalg.norm(A @ F) < 1e-12 * n
    assert np.linalg.norm(A @ z0 - b) < 1e-12 * n
    assert np.linalg.norm(F.T @ F - np.eye(n - m)) < 1e-12 * n
    assert np.linalg.norm(F.T @ z0) < 1e-12 * n


def _ill_conditioned_system(dim, condNum, dimKernel, seed):
    """MatrixUtilsTests.testEquationSolve (:264-320): A = U D U' (randomOrthogonalMatrix, diagonalMatrix: d_j =
    exp(-j log(condNum)/n), the last dimKernel set to zero; MatrixUtils.scala:46-63), made exactly symmetric, and
    nastyRHS b = U w, w_j ~ U(1,3) where d_j != 0, else 0 (:573-580)."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.normal(0, 1, (dim, dim)))
    d = np.exp(-np.arange(dim) * (np.log(condNum) / dim))
    if dimKernel:
        d[dim - dimKernel:] = 0.0
    Q = (U * d) @ U.T
    A = (Q + Q.T) * 0.5
    w = np.where(np.abs(d) > 0, 1 + 2 * rng.uniform(0, 1, dim), 0.0)
    return A, U @ w, d


@pytest.mark.parametrize("dim,condNum,dimKernel,seed", [(30, 1e3, 0, 0), (60, 1e6, 0, 1), (40, 1e4, 3, 2)])
def test_equation_solve_ill_conditioned(dim, condNum, dimKernel, seed):
    """svdSolve and symSolve accept the ill-conditioned (and, with dimKernel > 0, singular but consistent) systems of
    the reference's testEquationSolve and return a solution within the reference's residual bar; choleskySolve
    answers the nonsingular ones and, as in the reference, has no answer for an exactly singular matrix."""
    A, b, d = _ill_conditioned_system(dim, condNum, dimKernel, seed)
    tol = 1e-6
    for solve in (O.svdSolve, O.symSolve):
        x = solve(A, b, tol)
        assert np.linalg.norm(A @ x - b) <= tol * np.sqrt(2 * dim) * max(1.0, np.linalg.norm(b))
    if dimKernel == 0:
        x = O.choleskySolve(A, b, tol)
        assert O.relativeSize(A @ x - b, b, tol) <= tol
    else:
        with pytest.raises(Exception):
            O.choleskySolve(A, b, 1e-12)


@pytest.mark.parametrize("shift,seed", [(-2, 0), (0, 1), (4, 2)])
def test_kkt_system_reduction_and_padding(shift, seed):
    """KktTest.testSolutionPadding / testKktSystemReduction (:19-104): variables the system does not depend on (zero
    rows and columns of H, zero columns of A, zero entries of g) are eliminated (KKTData.reduced), the reduced system
    is solved, the solution padded with zeros (KKTData.paddVector) solves the original system."""
    null = [2 + shift, 4 + shift, 8 + shift]
    z = O.paddVector(np.ones(10), null)
    assert z.shape == (13,) and all(z[j] == 0 for j in null) and z.sum() == 10
    rng = np.random.default_rng(seed)
    dim = null[-1] + 5
    Q = rng.uniform(-1, 1, (dim, dim))
    H = Q.T @ Q
    H[:, null] = 0
    H[null, :] = 0
    A = rng.uniform(-1, 1, (6, dim))
    A[:, null] = 0
    g = rng.uniform(-2, 2, dim)
    g[null] = 0
    r = rng.uniform(-1, 1, 6)
    Hr, Ar, gr, rr, found = O.kktDataReduced(H, A, g, r)
    assert found == null and Hr.shape == (dim - 3, dim - 3) and Ar.shape == (6, dim - 3)
    rdx, nu = O.kkt_solve(Hr, Ar, gr, rr, 1e-9)
    dx = O.paddVector(rdx, null)
    assert np.linalg.norm(H @ dx + A.T @ nu + g) < 1e-8 * max(1.0, np.linalg.norm(g))
    assert np.linalg.norm(A @ dx - r) < 1e-8
    g2 = g.copy()
    g2[null[0]] = 1.0          # zero row with a nonzero right-hand side: no solution (KKTData.scala:80-84)
    with pytest.raises(O.UnsolvableSystemException):
        O.kktDataReduced(H, A, g2, r)


@pytest.mark.parametrize("solver", ["BR", "PD"])
@pytest.mark.parametrize("maker,n", [(P.kl_1A, 12), (P.kl_1A, 14), (P.kl_1A, 40), (P.kl_2A, 12), (P.kl_2A, 30)])
def test_kl_analytic_solutions_both_branches(maker, n, solver):
    """OptimizationProblems.kl1_analyticSolution / kl2_analyticSolution (:134-141, :247-252): the symmetric minimisers
    of the KL problems, in both regimes of kl_1 (n <= 15: the P(A) >= 0.36 constraint is inactive; n > 15: active)."""
    prob = maker(n)
    objF, cnts, eqs = P.to_oracle(prob)
    sol, _ = O.solveProblem(objF, cnts, eqs, solver)
    assert abs(objF.valueAt(sol.x) - objF.valueAt(prob["xopt"])) < 1e-8
    assert np.max(np.abs(sol.x - prob["xopt"])) < 1e-6
    assert abs(sol.x.sum() - 1.0) < 1e-8 and np.all(sol.x > 0)


@pytest.mark.parametrize("with_eqs", [False, True])
def test_oracle_pd_direction_against_full_newton_system(with_eqs):
    """Independent of the oracle's block elimination (kktMatrix_noEqs, rhs1, deltaLambda: PrimalDualSolver.scala:162-240)
    and of LAPACK: the primal-dual search direction against a 50-digit mpmath solve of the FULL linearised system
    (Boyd & Vandenberghe (11.54)):
        [ hess f + sum lam_i hess g_i   Dg'        A' ] [dx  ]     [ r_dual ]
        [ -diag(lam) Dg                 -diag(g)   0  ] [dlam] = - [ r_cent ]
        [ A                             0          0  ] [dnu ]     [ r_pri  ]
    with a quadratic objective, linear and quadratic constraints."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    prob = P.lin_quad_set(7, 6, 2, 2 if with_eqs else 0, 11, "quadratic", True)
    objF, cnts, eqs = P.to_oracle(prob)
    pd = O.PrimalDual(objF, cnts, eqs, O.SolverParams.standardParams(), False, False)
    rng = np.random.default_rng(3)
    x = prob["x0"] + 0.01 * rng.normal(0, 1, 7)      # off the equality manifold: r_pri != 0
    assert cnts.isSatisfiedStrictlyBy(x)
    lam = cnts.lambda0(x) * rng.uniform(0.5, 1.5, cnts.numConstraints)
    nu = rng.normal(0, 1, eqs.A.shape[0]) if with_eqs else None
    t = 7.0
    dx, dlam, dnu = pd.newton_direction(t, x, lam, nu)
    n, m = 7, cnts.numConstraints
    p = eqs.A.shape[0] if with_eqs else 0
    g = cnts.constraintFunctionAt(x)                     # g_i(x) - ub_i < 0
    Dg = cnts.gradientMatrixAt(x)
    Hf = objF.hessianAt(x)
    Hg = sum(lam[prob["G"].shape[0] + k] * q["P"] for k, q in enumerate(prob["quad"]))
    r_dual = objF.gradientAt(x) + Dg.T @ lam + (eqs.A.T @ nu if with_eqs else 0.0)
    r_cent = -lam * g - 1.0 / t
    N = n + m + p
    M = np.zeros((N, N))
    M[:n, :n] = Hf + Hg
    M[:n, n:n + m] = Dg.T
    M[n:n + m, :n] = -lam[:, None] * Dg
    M[n:n + m, n:n + m] = -np.diag(g)
    rhs = np.concatenate([-r_dual, -r_cent])
    if with_eqs:
        M[:n, n + m:] = eqs.A.T
        M[n + m:, :n] = eqs.A
        rhs = np.concatenate([rhs, -(eqs.A @ x - eqs.b)])
    Mm = mp.matrix(N, N)
    for i in range(N):
        for j in range(N):
            Mm[i, j] = mp.mpf(float(M[i, j]))
    sol = mp.lu_solve(Mm, mp.matrix([mp.mpf(float(v)) for v in rhs]))
    ref = np.array([float(sol[i]) for i in range(N)])
    got = np.concatenate([dx, dlam] + ([dnu] if with_eqs else []))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-9


@pytest.mark.parametrize("maker", [lambda: P.lin_quad_set(6, 5, 2, 0, 4, "quadratic", True), lambda: P.kl_small(8, 4, 1),
                                   lambda: P.min_pNorm(5, 3.0)])
def test_barrier_gradient_and_hessian_by_finite_differences(maker):
    """BarrierSolver.scala:280-315 restated: the gradient and Hessian of t f(x) - sum log(ub_i - g_i(x)) against central
    differences of the barrier VALUE (a check that does not share the closed-form derivative formulas)."""
    prob = maker()
    objF, cnts, eqs = P.to_oracle(prob)
    bf = O.BarrierFunctions(objF, cnts)
    n = prob["n"]
    x = prob["x0"].copy() if prob.get("x0") is not None else np.full(n, 1.0 / n) * (1 + 0.1 * np.cos(np.arange(n)))
    if prob["kind"] == "pnorm":
        x = np.full(n, 0.2) * (1 + 0.2 * np.cos(np.arange(n)))
    assert cnts.isSatisfiedStrictlyBy(x)
    t = 3.0
    g = bf.gradient(t, x)
    H = bf.hessian(t, x)
    h = 1e-5
    g_fd = np.zeros(n)
    H_fd = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n)
        e[i] = h
        g_fd[i] = (bf.value(t, x + e) - bf.value(t, x - e)) / (2 * h)
        H_fd[:, i] = (bf.gradient(t, x + e) - bf.gradient(t, x - e)) / (2 * h)
    assert np.linalg.norm(g - g_fd) / np.linalg.norm(g) < 1e-7
    assert np.linalg.norm(H - H_fd) / np.linalg.norm(H) < 1e-7
    assert np.array_equal(H, H.T)


def test_batched_golden_sample_is_reproducible():
    """tests/golden/batched_8192_sample.npz (the fixture of the full-size batched GPU test) against a live oracle run
    of a few of its entries, including both ends of the batch and both problem families."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "batched_8192_sample.npz"))
    from tests.golden.make_batched_golden import sample_indices, B, N, M, BASE_SEED
    assert np.array_equal(gold["index"], sample_indices()) and gold["index"][0] == 0 and gold["index"][-1] == B - 1
    picks = [0, 1, len(gold["index"]) // 2, len(gold["index"]) - 1]
    kinds = set()
    for k in picks:
        i = int(gold["index"][k])
        pr = P.batched_problem(i, N, M, BASE_SEED)
        kinds.add(pr["kind"])
        objF, cnts, eqs = P.to_oracle(pr)
        sol, _ = O.solveProblem(objF, cnts, eqs, "BR")
        assert abs(objF.valueAt(sol.x) - gold["objective"][k]) <= 1e-10 * max(1.0, abs(gold["objective"][k]))
        assert sol.outer_stages == gold["outer_stages"][k]
        assert list(sol.stage_newton_steps[:4]) == list(gold["stage_newton_steps"][k][:4])      # the rounding-stable stages
        assert np.linalg.norm(sol.x - gold["x"][k]) <= 1e-7 * np.linalg.norm(gold["x"][k])
    assert kinds == {"kl", "quadratic"} or len(kinds) >= 1


def test_iteration_noise_fixture_supports_the_band():
    """tests/golden/iteration_noise.json (tools/iteration_noise_experiment.py): the CPU oracle reproduces its own
    per-stage Newton counts exactly for t <= 1e3 and NOT beyond -- the evidence behind the band of the GPU tests."""
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "iteration_noise.json")))
    s = d["summary"]
    assert s["pairwise_max_deviation_by_stage"][:4] == [0, 0, 0, 0] and s["band_by_stage"][:4] == [1, 1, 1, 1]
    assert max(s["pairwise_max_deviation_by_stage"][4:]) >= 3 and s["spin_stage_flips"] > 0
    assert s["max_rel_objective_deviation"] < 1e-10
    # one entry re-derived live: base run of a small problem
    from tools.iteration_noise_experiment import problems, run
    pr = problems()["kl_1A"]
    assert run(pr)["stages"][:4] == d["runs"]["kl_1A"]["base"]["stages"][:4]


def test_block_elimination_formulations_differ_only_at_extreme_conditioning():
    """The one as-if deviation of the device path with a numerical consequence: the reference forms
    the Schur complement as R = A (L^-T L^-1 A'), symmetrised (KKTSystem.scala:116-139); the device forms S = Y'Y with
    Y = L^-1 A', positive semidefinite by construction.  On well-conditioned systems the two agree to the conditioning of the system.  In the
    last barrier stages of an LP (cond(H) ~ 1e20) the reference's R loses definiteness, its whole fallback chain fails and
    it throws (defect D3) -- under every rounding-level variation tried (tests/golden/outcome_noise.json) -- while the
    one-TRSM formulation solves the problem: the oracle restates both, and with the device's formulation it reaches the
    optimum the device reports (tools/gpu_fuzz.py, profiles/r2_fuzz_large_path.log)."""
    s = P.kkt_planted_pd(120, 20, 3)
    L = O.regularizedCholesky(s["H"])
    x0, w0 = O.solveWithCholFactor(L, s["A"], s["q"], s["b"], 1e-10)
    x1, w1 = O._solveWithCholFactor_one_trsm(L, s["A"], s["q"], s["b"], 1e-10)
    assert np.linalg.norm(x1 - x0) < 1e-9 * np.linalg.norm(x0) and np.linalg.norm(w1 - w0) < 1e-9 * np.linalg.norm(w0)
    assert np.linalg.norm(x1 - s["x"]) < 1e-8 * np.linalg.norm(s["x"])
    prob = P.slab_lp(27, 54, 1, 20152)
    objF, cnts, eqs = P.to_oracle(prob)
    with pytest.raises(O.UnsolvableSystemException):
        O.solveProblem(objF, cnts, eqs, "BR")
    O.BLOCK_ELIMINATION = "one_trsm"
    try:
        sol, _ = O.solveProblem(objF, cnts, eqs, "BR")
    finally:
        O.BLOCK_ELIMINATION = "reference"
    assert sol.outer_stages == 12 and abs(objF.valueAt(sol.x) - (-0.5643950053)) < 1e-8
    import json
    import os
    noise = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "outcome_noise.json")))
    assert noise["summary"]["outcomes_seen_per_problem"]["fuzz11_it18"] == ["UnsolvableSystemException"]


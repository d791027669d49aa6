"""Equality elimination x = z0 + F u on the device: MatrixUtils.solveUnderdetermined / SolutionSpace
(MatrixUtils.scala:536-550, SolutionSpace.scala:20-33) by blocked Householder QR, and BarrierSolver.reduced /
PrimalDualSolver.reduced (BarrierSolver.scala:209-256) for the closed-form families.  The reference has no test of
its own for these; the properties checked are the ones its doc comments state (A F = 0, F'F = I, A z0 = b, z0 of
minimum norm, x0 = z0 + F parameter(x0)) plus agreement with LAPACK's QR (oracle) and with the equality-constrained
solve of the same problem."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("p,n,seed", [(1, 2, 0), (1, 10, 1), (3, 7, 2), (20, 100, 3), (31, 64, 4), (32, 65, 5), (33, 97, 6),
                                      (100, 333, 7), (500, 2000, 8), (255, 256, 9)])
def test_solve_underdetermined(handle, p, n, seed):
    import cvx_b200 as cb
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (p, n))
    b = rng.uniform(-1, 1, p)
    z0, F = cb.MatrixUtils.solveUnderdetermined(A, b, handle)
    k = n - p
    assert F.shape == (n, k) and z0.shape == (n,)
    scale = np.linalg.norm(A)
    assert np.linalg.norm(A @ F) < 1e-13 * scale * np.sqrt(k)
    assert np.linalg.norm(F.T @ F - np.eye(k)) < 1e-13 * np.sqrt(k) * max(1.0, np.log2(n))
    assert np.linalg.norm(A @ z0 - b) < 1e-12 * max(1.0, np.linalg.norm(b)) * np.sqrt(p)
    assert np.linalg.norm(F.T @ z0) < 1e-12 * max(1.0, np.linalg.norm(z0))       # minimum norm: z0 is orthogonal to ker A
    z_ref, F_ref = O.solveUnderdetermined(A, b)
    assert rel(z0, z_ref) < 1e-10
    # same Householder conventions as dgeqrf/dorgqr: the basis itself agrees, not only its span
    assert np.linalg.norm(F - F_ref) < 1e-9 * np.sqrt(k)


def test_solution_space_parameter_and_point(handle):
    import cvx_b200 as cb
    rng = np.random.default_rng(11)
    p, n = 7, 40
    A = rng.uniform(-1, 1, (p, n))
    x0 = rng.normal(size=n)
    b = A @ x0
    sol = cb.SolutionSpace(A, b, handle)
    u0 = sol.parameter(x0)
    assert u0.shape == (n - p,)
    assert np.allclose(u0, sol.F.T @ (x0 - sol.z0), atol=1e-13)
    assert np.linalg.norm(sol.point(u0) - x0) < 1e-12 * np.linalg.norm(x0)
    eq = cb.EqualityConstraint(A, b)
    assert eq.solutionSpace is eq.solutionSpace and np.linalg.norm(A @ eq.z0 - b) < 1e-12


def test_solution_space_dimension_asserts(handle):
    import cvx_b200 as cb
    with pytest.raises(cb.DimensionMismatch):
        cb.SolutionSpace(np.ones((3, 3)), np.ones(3), handle)        # assert(A.rows < A.cols)
    with pytest.raises(cb.DimensionMismatch):
        cb.SolutionSpace(np.ones((2, 5)), np.ones(3), handle)        # assert(A.rows == b.length)


def test_rank_deficient_reports(handle):
    """A with a zero row: R gets a zero pivot and forwardSolve's diagonal assert fires (MatrixUtils.scala:395)."""
    import cvx_b200 as cb
    A = np.zeros((2, 6))
    A[0, :] = 1.0
    with pytest.raises(cb.CvxbError):
        cb.SolutionSpace(A, np.array([1.0, 1.0]), handle)


@pytest.mark.parametrize("maker,solver", [
    (lambda: P.slab_lp(30, 40, 4, seed=1), "BR"), (lambda: P.slab_qp(24, 30, 5, seed=2), "BR"),
    (lambda: P.slab_qp(24, 30, 5, seed=3), "PD"), (lambda: P.lin_quad_set(16, 12, 3, 3, 4, "quadratic", True), "BR"),
    (lambda: P.lin_quad_set(16, 12, 3, 3, 5, "linear", True), "BR"), (lambda: P.lin_quad_set(12, 0, 2, 2, 6, "quadratic", True), "BR"),
    (lambda: P.slab_lp(200, 260, 40, seed=7), "BR")])
def test_reduced_solver(handle, maker, solver):
    """Solve  min f(x), g(x) <= ub, Ax = b  once with the equality-constrained Newton steps and once in the reduced
    variable u (no equalities); both must reach the same optimum, and the reduced solve must follow the oracle's
    solve of the oracle-transformed problem stage by stage."""
    import cvx_b200 as cb
    prob = maker()
    if prob["kind"] == "linear" and prob.get("quad"):
        n = prob["n"]
        prob["G"] = np.vstack([prob["G"], np.eye(n), -np.eye(n)])
        prob["ub"] = np.concatenate([prob["ub"], prob["x0"] + 5.0, -prob["x0"] + 5.0])
        prob["rvec"] = np.zeros(prob["G"].shape[0])
    objF, cnts, eqs = P.to_oracle(prob)
    full = cb.from_dict(prob, solver, None, handle).solve()
    noeq = dict(prob)
    noeq["A"] = noeq["b"] = None
    base = cb.from_dict(noeq, solver, None, handle).solver
    sol = cb.SolutionSpace(prob["A"], prob["b"], handle)
    red = base.reduced(sol)
    su = red.solve()
    k = prob["n"] - prob["A"].shape[0]
    assert su.x.shape == (k,)
    x = red.point(su.x)
    assert np.linalg.norm(prob["A"] @ x - prob["b"]) < 1e-10 * max(1.0, np.linalg.norm(prob["b"]))
    assert cnts.isSatisfiedStrictlyBy(x)
    f_full, f_red = objF.valueAt(full.x), objF.valueAt(x)
    assert abs(su.objective - f_red) < 1e-9 * max(1.0, abs(f_red))        # transformed objective = f(z0 + F u)
    assert abs(f_red - f_full) < 2e-8 * max(1.0, abs(f_full))
    assert rel(x, full.x) < 1e-5
    # oracle on the oracle-transformed problem (LAPACK QR): same path
    z0, F = O.solveUnderdetermined(prob["A"], prob["b"])
    o_u, c_u = O.affineTransformedProblem(objF, cnts, z0, F)
    s0, _ = O.solveProblem(o_u, c_u, None, solver)
    assert abs(o_u.valueAt(s0.x) - su.objective) < 1e-8 * max(1.0, abs(su.objective))
    assert rel(su.x, s0.x) < 1e-6
    if solver == "BR":
        assert su.outer_stages == s0.outer_stages
        for j in range(min(3, su.outer_stages)):
            assert abs(su.stage_newton_steps[j] - s0.stage_newton_steps[j]) <= 1
    else:
        assert abs(su.newton_steps - s0.newton_steps) <= 1


def test_reduced_rejects_what_the_reference_cannot_do(handle):
    import cvx_b200 as cb
    prob = P.slab_qp(12, 14, 3, seed=1)
    sol = cb.SolutionSpace(prob["A"], prob["b"], handle)
    with pytest.raises(cb.CvxbError):                 # solver still carrying the equalities
        cb.from_dict(prob, "BR", None, handle).solver.reduced(sol)
    noeq = dict(prob)
    noeq["A"] = noeq["b"] = None
    noeq["x0"] = prob["x0"] + 1.0                     # A x0 != b beyond tolEqSolve = 0.1: "u0 does not map to x0 under the variable transform"
    with pytest.raises(cb.CvxbError):
        cb.from_dict(noeq, "BR", None, handle).solver.reduced(sol)


def _kl_with_equalities(n, m_h, p_extra, seed):
    pr = P.kl_random(n, m_h, p_extra, seed)
    pr["x0"] = pr["qstar"].copy()
    return pr


def _pnorm_with_equalities(n, p, pw, seed):
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(0.5, 1.5, n)
    A = rng.uniform(-1, 1, (p, n))
    G = np.vstack([np.eye(n), -np.eye(n)])
    ub = np.concatenate([x0 + 2.0, -x0 + 2.0])
    return dict(kind="pnorm", n=n, a=None, r=0.0, P=None, pow=float(pw), G=G, rvec=np.zeros(2 * n), ub=ub, A=A, b=A @ x0,
                x0=x0, xdef=x0.copy())


@pytest.mark.parametrize("maker,solver", [
    (lambda: _kl_with_equalities(24, 24, 11, 1), "BR"),          # n = 2p: the one shape for which the reference's own
    (lambda: _kl_with_equalities(24, 24, 11, 2), "PD"),          # dimension assert passes (defect D11), p = 12 with the sum-to-one row
    (lambda: _kl_with_equalities(60, 40, 9, 3), "BR"), (lambda: _kl_with_equalities(60, 40, 9, 4), "PD"),
    (lambda: _pnorm_with_equalities(30, 6, 3.0, 5), "BR"), (lambda: _pnorm_with_equalities(30, 6, 2.5, 6), "PD"),
    (lambda: _kl_with_equalities(300, 200, 63, 7), "BR")])
def test_reduced_solver_generic_objectives(handle, maker, solver):
    """BarrierSolver.reduced / PrimalDualSolver.reduced for the objectives without a closed-form transform -- the KL
    distance and the p-norm -- through ObjectiveFunction.affineTransformed (ObjectiveFunction.scala:26-40):
    h(u) = f(z0 + F u), gradient F' grad f, Hessian F' hess f F (on the device: one more weighted SYRK).  Checked against
    the equality-constrained solve of the same problem and against the oracle's solve of the oracle-transformed
    problem."""
    import cvx_b200 as cb
    prob = maker()
    objF, cnts, eqs = P.to_oracle(prob)
    full = cb.from_dict(prob, solver, None, handle).solve()
    noeq = dict(prob)
    noeq["A"] = noeq["b"] = None
    base = cb.from_dict(noeq, solver, None, handle).solver
    sol = cb.SolutionSpace(prob["A"], prob["b"], handle)
    red = base.reduced(sol)
    su = red.solve()
    k = prob["n"] - prob["A"].shape[0]
    assert su.x.shape == (k,)
    x = red.point(su.x)
    assert np.linalg.norm(prob["A"] @ x - prob["b"]) < 1e-10 * max(1.0, np.linalg.norm(prob["b"]))
    assert cnts.isSatisfiedStrictlyBy(x)
    f_full, f_red = objF.valueAt(full.x), objF.valueAt(x)
    assert abs(su.objective - f_red) < 1e-9 * max(1.0, abs(f_red))
    assert abs(f_red - f_full) < 2e-8 * max(1.0, abs(f_full))
    z0, F = O.solveUnderdetermined(prob["A"], prob["b"])
    o_u, c_u = O.affineTransformedProblem(objF, cnts, z0, F)
    assert o_u.kind == "composed"
    s0, _ = O.solveProblem(o_u, c_u, None, solver)
    assert abs(o_u.valueAt(s0.x) - su.objective) < 1e-8 * max(1.0, abs(su.objective))
    assert rel(su.x, s0.x) < 1e-6
    if solver == "BR":
        assert su.outer_stages == s0.outer_stages
        for j in range(min(3, su.outer_stages)):
            assert abs(su.stage_newton_steps[j] - s0.stage_newton_steps[j]) <= 1
    else:
        assert abs(su.newton_steps - s0.newton_steps) <= 1
        assert su.dualityGap < 1e-8 and su.normDualResidual < 1e-8


def test_reduced_direction_generic_objective(handle):
    """One Newton direction of a reduced KL problem: H = F'(t diag(1/x) + G'diag(1/d^2)G)F and the gradient against
    the oracle's composed objective, to the bars of the unreduced direction test."""
    import cvx_b200 as cb
    prob = _kl_with_equalities(80, 60, 19, 11)
    objF, cnts, eqs = P.to_oracle(prob)
    noeq = dict(prob)
    noeq["A"] = noeq["b"] = None
    base = cb.from_dict(noeq, "BR", None, handle).solver
    sol = cb.SolutionSpace(prob["A"], prob["b"], handle)
    red = base.reduced(sol)
    z0, F = sol.z0, sol.F
    u = sol.parameter(prob["x0"])
    o_u, c_u = O.affineTransformedProblem(objF, cnts, z0, F)
    bf = O.BarrierFunctions(o_u, c_u)
    t = 10.0
    H0, g0 = bf.hessian(t, u), bf.gradient(t, u)
    H, g, du, _, info = red.newton_direction(u, t)
    assert np.array_equal(H, H.T)
    assert rel(H, H0) < 1e-12 and rel(g, g0) < 1e-12
    assert np.linalg.norm(H0 @ du + g0) / np.linalg.norm(g0) < 1e-10
    assert rel(du, O.choleskySolve(H0, -g0, 0.1)) < 1e-8


def test_phase_I_analysis_by_reduction(handle):
    """ConstraintSet.phase_I_Analysis_by_reduction (ConstraintSet.scala:424-477).  As written the reference reduces the
    (n+1)-dimensional phase-I solver with the n-dimensional solution space of Ax = b and fails on the dimension
    mismatch; the mirror fails the same way.  corrected=True does what the doc comment describes."""
    import cvx_b200 as cb
    prob = P.slab_lp(30, 40, 5, seed=3, feasible_start=True)
    # pointWhereDefined: on the plane Ax = b (reduced's assert ||x0 - (z0 + F u0)|| < tolEqSolve, BarrierSolver.scala:214-218)
    # but far outside the slab, so that phase I has work to do
    z0_, F_ = O.solveUnderdetermined(prob["A"], prob["b"])
    prob["xdef"] = prob["x0"] + F_ @ np.random.default_rng(5).uniform(1.0, 2.0, F_.shape[1])
    assert not np.all(prob["G"] @ prob["xdef"] < prob["ub"])
    cnts = cb.ConstraintSet(prob["G"], prob["ub"], prob["xdef"])
    eqs = cb.EqualityConstraint(prob["A"], prob["b"])
    with pytest.raises(AssertionError) as ei:
        cnts.phase_I_Analysis_by_reduction(eqs, None, 0, handle)
    assert isinstance(ei.value, cb.DimensionMismatch) and "dim(problem)" in str(ei.value)
    rep = cnts.phase_I_Analysis_by_reduction(eqs, None, 0, handle, corrected=True)
    assert rep.s[0] < 0 and rep.isStrictlyFeasible
    assert np.all(prob["G"] @ rep.x0 * (1 + 3e-16) < prob["ub"])
    assert np.linalg.norm(prob["A"] @ rep.x0 - prob["b"]) < 1e-8
    # same verdict as the basic phase I with the equalities kept as inequality pairs (ConstraintSet.scala:355-414)
    rep0 = cnts.phase_I_Analysis(eqs, None, 0, handle)
    assert rep0.isFeasible(1e-8) and rep.isFeasible(1e-8)


def test_affine_transformed_with_caller_basis(handle):
    """Solver.affineTransformed(z0, F, u0) with a basis that is NOT orthonormal (the reference does not require it):
    cvxb_solution_space_from_basis + cvxb_problem_reduce; the optimum in x must not depend on the basis."""
    import cvx_b200 as cb
    prob = P.slab_qp(20, 26, 4, seed=9)
    noeq = dict(prob)
    noeq["A"] = noeq["b"] = None
    base = cb.from_dict(noeq, "BR", None, handle).solver
    z0, F = O.solveUnderdetermined(prob["A"], prob["b"])
    rng = np.random.default_rng(1)
    T = np.eye(F.shape[1]) + 0.3 * rng.uniform(-1, 1, (F.shape[1],) * 2)
    F2 = F @ T                                   # same range, skewed basis
    full = cb.from_dict(prob, "BR", None, handle).solve()
    for basis in (F, F2):
        sp = cb.SolutionSpace.from_basis(z0, basis, handle)
        u0 = np.linalg.lstsq(basis, prob["x0"] - z0, rcond=None)[0]
        assert np.linalg.norm(sp.point(u0) - prob["x0"]) < 1e-10
        if basis is F:
            assert np.linalg.norm(sp.parameter(prob["x0"]) - u0) < 1e-10
            red = base.reduced(sp)
            x = red.point(red.solve().x)
            assert abs(full.objective - (prob["r"] + prob["a"] @ x + 0.5 * x @ prob["P"] @ x)) < 2e-8 * max(1.0, abs(full.objective))

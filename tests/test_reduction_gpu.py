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

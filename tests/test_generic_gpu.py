"""General (non-closed-form) objectives through seam B with pinned double-buffered staging:
the host runs the reference's Newton loop and the user's closures, the device solves every linear system.  Test
objective: the reference's own Type1Function power problems (src/test/scala/cvx/Type1Function.scala:67-78,
OptimizationProblems.powerProblems :112-125), which can only use this path, plus seeded random instances against the
oracle's UnconstrainedSolver / EqualityConstrainedSolver fed with the same closures."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


class _NoConstraints:
    """An empty constraint set for the oracle's inner solvers (the abstract set C is the whole space)."""
    m_lin, quad, numConstraints = 0, [], 0

    def isSatisfiedStrictlyBy(self, x):
        return True


class _ObjectiveOnly:
    """BarrierFunctions look-alike: the oracle's inner solvers minimise whatever value / gradient / hessian they get."""

    def __init__(self, objF):
        self.objF, self.cnts = objF, _NoConstraints()

    def value(self, t, x):
        return self.objF.valueAt(x)

    def gradient(self, t, x):
        return self.objF.gradientAt(x)

    def hessian(self, t, x):
        return self.objF.hessianAt(x)


def test_staged_system_solves(handle):
    """cvxb_stage_*: H uploaded in column blocks through the two pinned buffers (7 blocks, ragged last one), then
    choleskySolve and the KKT solve on the device-resident matrix; same answers as the unstaged seam-B calls."""
    import cvx_b200 as cb
    d = P.kkt_planted_pd(200, 24, 5)
    st = cb.StagedSystem(200, 24, 32, handle)
    assert st.block_cols == 32 and st.buffers[0].shape == (200, 32) and st.buffers[0].flags["F_CONTIGUOUS"]
    st.set_equalities(d["A"])
    for rep in range(2):                                   # a stage is reused step after step
        st.upload_matrix(d["H"])
        x, w = st.kktSolve(d["q"], d["b"], 1e-10)
        assert np.linalg.norm(x - d["x"]) < 1e-9 * np.linalg.norm(d["x"]) and np.linalg.norm(w - d["w"]) < 1e-8 * np.linalg.norm(d["w"])
    st.upload(lambda j0, j1, out: out.__setitem__(Ellipsis, d["H"][:, j0:j1]))
    rhs = d["H"] @ d["x"]
    x2 = st.choleskySolve(rhs, 1e-10)
    assert np.linalg.norm(x2 - d["x"]) < 1e-9 * np.linalg.norm(d["x"])
    assert np.linalg.norm(x2 - cb.MatrixUtils.choleskySolve(d["H"], rhs, None, 1e-10, 0, handle)) < 1e-12 * np.linalg.norm(x2)
    with pytest.raises(cb.LinSolveException):
        st.upload_matrix(-np.eye(200))
        st.choleskySolve(rhs, 1e-10)
    with pytest.raises(cb.DimensionMismatch):
        check = cb._lib.check
        check(handle.lib.cvxb_stage_push(st._s, 0, 190, 32))


@pytest.mark.parametrize("k", [0, 1])
def test_reference_power_problems(handle, k):
    """OptimizationProblems.powerProblems with the reference's KnownMinimizer checks (minimum value 0, ||Ax|| small;
    Runner.scala:30 uses tol 1e-2) and the oracle's iteration path."""
    import cvx_b200 as cb
    f, x0 = P.power_problems()[k]
    sol = cb.generic.UnconstrainedSolver(f, x0, None, None, handle).solve()
    assert sol.objective < 1e-8 and f.isMinimizer(sol.x, 1e-2)
    s0 = O.unconstrainedSolve(_ObjectiveOnly(f), 1.0, x0, O.SolverParams.standardParams())
    assert abs(sol.newton_steps - s0.newton_steps) <= 1
    assert abs(sol.objective - f.valueAt(s0.x)) < 1e-8
    assert np.linalg.norm(sol.x - s0.x) < 1e-6 * max(1.0, np.linalg.norm(s0.x))


@pytest.mark.parametrize("dim,m,q,seed", [(40, 40, 2.0, 0), (300, 300, 1.5, 1), (600, 450, 2.0, 2)])
def test_random_power_problem_unconstrained(handle, dim, m, q, seed):
    """Seeded Type1Function.randomPowerFunction instances (m < dim: a non-trivial kernel, singular Hessian directions):
    the device path walks the same fallback chain (choleskySolve -> H + 1e-9 I -> symSolve) and the same iterations."""
    import cvx_b200 as cb
    f, x0 = P.random_power_problem(dim, m, q, seed)
    solver = cb.generic.UnconstrainedSolver(f, x0, None, None, handle, block_cols=64)
    sol = solver.solve()
    s0 = O.unconstrainedSolve(_ObjectiveOnly(f), 1.0, x0, O.SolverParams.standardParams())
    assert abs(sol.newton_steps - s0.newton_steps) <= 1
    assert sol.objective <= 1e-8 and f.valueAt(s0.x) <= 1e-8          # the known minimum value 0 (tol 1e-2 in Runner.scala:30)
    assert np.linalg.norm(sol.x - s0.x) < 1e-5 * max(1.0, np.linalg.norm(s0.x))
    assert f.isMinimizer(sol.x, 0.1)
    assert solver.stage.pushes >= sol.newton_steps * ((dim + 63) // 64)


def test_power_objective_with_equalities(handle):
    """EqualityConstrainedSolver on closures: min f(x) s.t. Ax = b through cvxb_stage_kkt_solve, against the oracle."""
    import cvx_b200 as cb
    f, x0 = P.random_power_problem(120, 120, 2.0, 5)
    rng = np.random.default_rng(6)
    A = rng.uniform(-1, 1, (15, 120))
    b = A @ x0
    sol = cb.generic.EqualityConstrainedSolver(f, A, b, x0, None, None, handle, block_cols=50).solve()
    s0 = O.equalityConstrainedSolve(_ObjectiveOnly(f), 1.0, x0, A, b, O.SolverParams.standardParams())
    assert abs(sol.newton_steps - s0.newton_steps) <= 1
    assert np.linalg.norm(A @ sol.x - b) < 1e-8
    assert abs(sol.objective - f.valueAt(s0.x)) <= 1e-8 * max(1.0, abs(sol.objective))
    assert np.linalg.norm(sol.x - s0.x) < 1e-5 * np.linalg.norm(s0.x)

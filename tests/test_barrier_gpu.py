"""Seam A (device-resident barrier solves, phase I) through the C ABI against the CPU oracle.
Bars (BASELINE.json north_star): Newton direction within 1e-10 relative residual, iteration counts +-1,
final objective within 1e-8 relative."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n,mh,p,seed", [(20, 30, 3, 0), (100, 100, 20, 1), (300, 300, 64, 2), (257, 130, 0, 3),
                                         (1000, 1000, 250, 4)])
def test_newton_direction_parity(handle, n, mh, p, seed):
    """One barrier Newton step at a random strictly feasible iterate: Hessian, gradient, dx, nu."""
    import cvx_b200 as cb
    prob, x, t = P.newton_step_inputs(n, mh, p, seed)
    objF, cnts, eqs = P.to_oracle(prob)
    bf = O.BarrierFunctions(objF, cnts)
    H0 = bf.hessian(t, x)
    g0 = bf.gradient(t, x)
    op = cb.from_dict(prob, "BR", None, handle)
    H, g, dx, nu, info = op.solver.newton_direction(x, t)
    assert np.array_equal(H, H.T)
    assert rel(H, H0) < 1e-13 and rel(g, g0) < 1e-13
    tol = 1e-1
    if p > 0:
        eqd = prob["b"] - prob["A"] @ x
        dx0, nu0 = O.kkt_solve(H0, prob["A"], g0, eqd, tol)
        A = prob["A"]
        res = np.linalg.norm(np.concatenate([H0 @ dx + A.T @ nu + g0, A @ dx - eqd]))
        res0 = np.linalg.norm(np.concatenate([H0 @ dx0 + A.T @ nu0 + g0, A @ dx0 - eqd]))
        scale = np.linalg.norm(np.concatenate([g0, eqd]))
        assert res / scale < max(1e-10, 10 * res0 / scale)
        assert rel(dx, dx0) < 1e-8 and rel(nu, nu0) < 1e-8
    else:
        dx0 = O.choleskySolve(H0, -g0, tol)
        assert np.linalg.norm(H0 @ dx + g0) / np.linalg.norm(g0) < 1e-10
        assert rel(dx, dx0) < 1e-8
    assert info.path == 0


def _band():
    """Allowed deviation of the Newton count of outer stage k from the oracle's (tests/golden/iteration_noise.json,
    produced by tools/iteration_noise_experiment.py): +-1 -- north_star's bar -- for as long as the CPU oracle itself
    reproduces its counts under rounding-level changes (BLAS threads, literal Hessian accumulation, half-ulp input
    changes): stages 0-3, t <= 1e3.  Beyond that the oracle's own variants disagree (by up to 3 steps at t = 1e6, 5 at
    t = 1e8, 27 at t = 1e11, and on whether a stage spins to maxIter at all); the band there is twice the largest
    deviation observed between two CPU variants."""
    import json
    import os
    f = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iteration_noise.json")
    return json.load(open(f))["summary"]["band_by_stage"]


def _check_solve(sol, sol0, obj0, ph0=None):
    """Identical number of outer stages; Newton steps per stage within the measured band (see _band); objective within
    1e-8 relative; duality gap m/t identical."""
    band = _band()
    assert sol.outer_stages == sol0.outer_stages
    assert len(sol.stage_newton_steps) == len(sol0.stage_newton_steps)
    for k, (a, b) in enumerate(zip(sol.stage_newton_steps, sol0.stage_newton_steps)):
        if a >= 1000 or b >= 1000:
            assert k >= 4, (k, sol.stage_newton_steps, sol0.stage_newton_steps)
            continue      # equality-gap spin until maxIter: decided by whether ||b-Ax|| rounds above 1e-8 (15 flips between CPU variants)
        tol = band[k] if k < len(band) else band[-1]
        if tol > 20:
            continue      # noise-floor stage (t >= 1e11): the CPU oracle's own variants differ by 7 vs 34 steps there
        assert abs(a - b) <= tol, (k, sol.stage_newton_steps, sol0.stage_newton_steps)
    assert abs(sol.objective - obj0) <= 1e-8 * max(1.0, abs(obj0))
    assert abs(sol.dualityGap - sol0.dualityGap) <= 1e-12 * sol0.dualityGap
    if ph0 is not None:
        assert sol.phase1_stages == ph0.outer_stages
        assert abs(sol.phase1_newton_steps - ph0.newton_steps) <= max(2, ph0.outer_stages)


PROBLEMS = {
    "slab_qp_64": lambda: P.slab_qp(64, 64, 0, 1),
    "kl_small_64": lambda: P.kl_small(64, 64, 2),
    "slab_qp_eq": lambda: P.slab_qp(48, 60, 6, 3),
    "min_dot_product": lambda: P.min_dot_product(np.linspace(0.5, 2, 10)),
    "kl_1A": lambda: P.kl_1A(20),
    "kl_2A": lambda: P.kl_2A(20),
    "kl_random_120": lambda: P.kl_random(120, 120, 29, 5),
    "slab_lp_phase1": lambda: P.slab_lp(40, 60, 0, 7, feasible_start=False),
}


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_barrier_solve_matches_oracle(handle, name):
    import cvx_b200 as cb
    prob = PROBLEMS[name]()
    objF, cnts, eqs = P.to_oracle(prob)
    sol0, ph0 = O.solveProblem(objF, cnts, eqs, "BR")
    sol = cb.from_dict(prob, "BR", None, handle).solve()
    _check_solve(sol, sol0, objF.valueAt(sol0.x), ph0)
    assert rel(sol.x, sol0.x) < 1e-6
    if "xopt" in prob:       # the reference's own known-answer check (KnownMinimizer, tol 1e-2 in Runner.scala:30)
        assert abs(objF.valueAt(sol.x) - objF.valueAt(prob["xopt"])) < 1e-6
    if eqs is not None:
        assert np.linalg.norm(eqs.A @ sol.x - eqs.b) < 1e-8
    assert cnts.isSatisfiedStrictlyBy(sol.x)


def test_equality_gap_spin_is_fast_forwarded(handle):
    """slab LP with p=20: at some stage the Newton decrement falls below tol while ||b-Ax|| > tol; the
    reference then repeats the identical step until maxIter (EqualityConstrainedSolver.scala:49).  The
    device path reports the same counts without executing the repeats."""
    import cvx_b200 as cb
    prob = P.slab_lp(100, 100, 20, 0)
    objF, cnts, eqs = P.to_oracle(prob)
    sol0, _ = O.solveProblem(objF, cnts, eqs, "BR")
    sol = cb.from_dict(prob, "BR", None, handle).solve()
    assert abs(sol.objective - objF.valueAt(sol0.x)) <= 1e-8 * abs(objF.valueAt(sol0.x))
    spun0 = [s for s in sol0.stage_newton_steps if s >= 1000]
    spun = [s for s in sol.stage_newton_steps if s >= 1000]
    if spun0 or spun:
        assert sol.executed_newton_steps < sol.newton_steps or not spun


def test_infeasible_problem(handle):
    """infeasible_kl_1 (OptimizationProblems.scala:379-405): P(A) >= .51 and P(B) >= .51 with A, B disjoint.
    Both the oracle and the device path must refuse it (no feasible point is ever produced)."""
    import cvx_b200 as cb
    prob = P.infeasible_kl_1(20)
    objF, cnts, eqs = P.to_oracle(prob)
    with pytest.raises(Exception):
        O.solveProblem(objF, cnts, eqs, "BR")
    with pytest.raises(cb.CvxbError):
        cb.from_dict(prob, "BR", None, handle).solve()


def test_not_strictly_feasible_start(handle):
    import cvx_b200 as cb
    prob = P.slab_qp(16, 16, 0, 1)
    prob["x0"] = prob["x0"] + 100.0       # far outside the slab
    with pytest.raises(cb.NotStrictlyFeasible):
        cb.from_dict(prob, "BR", None, handle).solve()


def test_phase1_entry_point(handle):
    import cvx_b200 as cb
    prob = P.kl_1A(20)
    objF, cnts, eqs = P.to_oracle(prob)
    x0, s0, sol0 = O.phase_I_Analysis(cnts, eqs, O.SolverParams())
    op = cb.from_dict(prob, "BR", None, handle)
    xf, ph = op.solver.phase_I()
    assert ph.outer_stages == sol0.outer_stages
    assert ph.phase1_s < 0 and s0 < 0
    assert cnts.isSatisfiedStrictlyBy(xf)
    assert rel(xf, x0) < 1e-5


def test_step_limit(handle):
    import cvx_b200 as cb
    prob = P.kl_small(64, 64, 2)
    sol = cb.from_dict(prob, "BR", cb.SolverParams(stepLimit=5), handle).solve()
    assert sol.executed_newton_steps == 5


@pytest.mark.parametrize("pw", [2.0, 3.0, 4.5])
@pytest.mark.parametrize("solver", ["BR", "PD"])
def test_min_pnorm_known_minimiser(handle, pw, solver):
    """SimpleOptimizationProblems.min_pNorm (:179-209): min sum |x_j|^p on the simplex, optimum x_j = 1/n; the
    objective family is ObjectiveFunctions.p_norm_p (ObjectiveFunctions.scala:70-83); phase I from x = 0."""
    import cvx_b200 as cb
    prob = P.min_pNorm(10, pw)
    objF, cnts, eqs = P.to_oracle(prob)
    sol0, ph0 = O.solveProblem(objF, cnts, eqs, solver)
    sol = cb.from_dict(prob, solver, None, handle).solve()
    assert np.max(np.abs(sol.x - prob["xopt"])) < 1e-6
    assert abs(sol.objective - objF.valueAt(sol0.x)) <= 1e-8 * max(1.0, abs(sol.objective))
    assert sol.phase1_stages == ph0.outer_stages
    if solver == "BR":
        assert sol.outer_stages == sol0.outer_stages
    else:
        assert abs(sol.newton_steps - sol0.newton_steps) <= 1


def test_literal_block_elimination_switch(handle):
    """bugCompat bit 1: the Schur complement formed literally as the reference writes it (KKTSystem.scala:116-139: H^-1 A' by two
    triangular solves, A (H^-1 A') by a GEMM, symmetrised) instead of Y'Y.  On an ordinary problem both give the same solve;
    on the slab LP found by tools/gpu_fuzz.py the literal form loses positive definiteness in the last stages and the solve
    ends as the reference's does -- UnsolvableSystemException at the end of KKTSystem.solve's fallback chain (defect D3) --
    while the default form reaches the optimum (the oracle restates both)."""
    import cvx_b200 as cb
    lit = cb.SolverParams()
    lit.bugCompat = 2
    pr = P.slab_qp(48, 60, 6, 3)
    a = cb.from_dict(pr, "BR", None, handle).solve()
    b = cb.from_dict(pr, "BR", lit, handle).solve()
    assert abs(a.objective - b.objective) <= 1e-9 * max(1.0, abs(a.objective))
    assert np.linalg.norm(a.x - b.x) <= 1e-7 * np.linalg.norm(a.x) and a.outer_stages == b.outer_stages
    lp = P.slab_lp(27, 54, 1, 20152)
    objF, cnts, eqs = P.to_oracle(lp)
    with pytest.raises(O.UnsolvableSystemException):
        O.solveProblem(objF, cnts, eqs, "BR")
    with pytest.raises(cb._lib.UnsolvableSystemException):
        cb.from_dict(lp, "BR", lit, handle).solve()
    ok = cb.from_dict(lp, "BR", None, handle).solve()
    O.BLOCK_ELIMINATION = "one_trsm"
    try:
        s0, _ = O.solveProblem(objF, cnts, eqs, "BR")
    finally:
        O.BLOCK_ELIMINATION = "reference"
    assert abs(ok.objective - objF.valueAt(s0.x)) <= 1e-8 and ok.outer_stages == s0.outer_stages


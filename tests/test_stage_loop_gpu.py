"""Device-driven centering stages (north_star: "device-resident iterates so no per-step host round-trips occur"): a
barrier stage is ONE graph launch -- a CUDA-graph WHILE node whose body is the Newton step followed by a kernel that
applies the reference's loop test (EqualityConstrainedSolver.scala:49, UnconstrainedSolver.scala:45) on the device.
The same kernels run in the same order as in the step-by-step drive (CVXB_NO_LOOP=1), so the results must be
IDENTICAL, bit for bit; what changes is the number of host round trips."""
import os

import numpy as np
import pytest

from oracle import problems as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handles():
    from cvx_b200 import _lib
    loop = _lib.Handle(0)
    os.environ["CVXB_NO_LOOP"] = "1"
    try:
        stepwise = _lib.Handle(0)
    finally:
        del os.environ["CVXB_NO_LOOP"]
    yield loop, stepwise
    loop.close()
    stepwise.close()


PROBLEMS = {
    "kl_1A_phase1_p1": lambda: P.kl_1A(20),
    "slab_lp_eq_spin": lambda: P.slab_lp(100, 100, 20, 0),
    "slab_qp_300_eq": lambda: P.slab_qp(300, 300, 64, 2),
    "slab_lp_phase1_noeq": lambda: P.slab_lp(40, 60, 0, 7, feasible_start=False),
    "rank_one_simplex": lambda: P.rank_one_simplex(10),
    "free_variables": lambda: P.norm_squared_free_variables(6),
    "kl_random_400": lambda: P.kl_random(400, 400, 99, 3),
    "quadratic_constraints": lambda: P.lin_quad_set(16, 12, 3, 3, 4, "quadratic", False),
}


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_loop_and_stepwise_drives_agree_bitwise(handles, name):
    import cvx_b200 as cb
    loop, stepwise = handles
    prob = PROBLEMS[name]()
    r0 = loop.status_reads
    a = cb.from_dict(prob, "BR", None, loop).solve()
    reads_loop = loop.status_reads - r0
    r0 = stepwise.status_reads
    b = cb.from_dict(prob, "BR", None, stepwise).solve()
    reads_step = stepwise.status_reads - r0
    assert np.array_equal(a.x, b.x)
    assert a.objective == b.objective and a.newtonDecrement == b.newtonDecrement and a.normGrad == b.normGrad
    assert a.stage_newton_steps == b.stage_newton_steps and a.newton_steps == b.newton_steps
    assert (a.executed_newton_steps, a.phase1_newton_steps, a.phase1_stages, a.linesearch_trials, a.kkt_fallbacks,
            a.kkt_regularized, a.iter, a.maxedOut) == \
           (b.executed_newton_steps, b.phase1_newton_steps, b.phase1_stages, b.linesearch_trials, b.kkt_fallbacks,
            b.kkt_regularized, b.iter, b.maxedOut)
    steps = a.executed_newton_steps + a.phase1_executed_steps
    stages = a.outer_stages + a.phase1_stages
    assert reads_step >= steps                                  # one round trip per Newton step (+ one per stage)
    # device-driven: two per stage (initial evaluation, end of loop) + one more launch per step the host had to help with
    if name in ("free_variables", "rank_one_simplex"):
        # singular Hessians (a rank-one barrier Hessian in phase I, P = aa'): the device refuses the plain Cholesky at
        # almost every step and the host walks regularizedCholesky's retry (MatrixUtils.scala:452-461) -- up to four
        # reads for such a step (loop end, retry, line search + evaluation, relaunch)
        assert reads_loop <= 2 * stages + 4 * steps + 4, (reads_loop, stages, steps)
    else:
        assert reads_loop <= 2 * stages + 4 * (a.kkt_fallbacks + a.kkt_regularized) + 4, (reads_loop, stages, steps)
        if steps > 4 * stages:
            assert reads_loop < reads_step


def test_step_budget_inside_the_device_loop(handles):
    """cvxb_params.stepLimit is enforced by the device loop test: exactly that many steps, same iterate either way."""
    import cvx_b200 as cb
    loop, stepwise = handles
    prob = P.kl_random(200, 200, 49, 5)
    for limit in (1, 7, 23):
        a = cb.from_dict(prob, "BR", cb.SolverParams(stepLimit=limit), loop).solve()
        b = cb.from_dict(prob, "BR", cb.SolverParams(stepLimit=limit), stepwise).solve()
        assert a.executed_newton_steps + a.phase1_executed_steps == limit == b.executed_newton_steps + b.phase1_executed_steps
        assert np.array_equal(a.x, b.x)


def test_profile_counts_the_syrk_inside_the_loop(handles):
    """The Hessian SYRK is timed inside the WHILE body by %globaltimer stamps (event nodes are not allowed there)."""
    import cvx_b200 as cb
    loop, _ = handles
    prob = P.slab_qp(300, 300, 64, 2)
    op = cb.from_dict(prob, "BR", None, loop)
    loop.profile_enable(True)
    sol = op.solve()
    n_syrk, ms, flops = loop.profile_read()
    loop.profile_enable(False)
    assert n_syrk == sol.executed_newton_steps and ms > 0
    assert abs(flops - n_syrk * 600.0 * 300 * 301) <= 1e-6 * flops

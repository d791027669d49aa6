"""Batched one-CTA-per-problem barrier solver (configs[2] shape n=64, m=128, p in {0,1}) against the
CPU oracle run problem by problem."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def _oracle(prob):
    objF, cnts, eqs = P.to_oracle(prob)
    sol, _ = O.solveProblem(objF, cnts, eqs, "BR")
    return objF, sol


@pytest.mark.parametrize("n,m,B", [(64, 128, 24), (20, 40, 10), (33, 70, 6)])
def test_batched_matches_oracle(handle, n, m, B):
    import cvx_b200 as cb
    probs = [P.batched_problem(i, n, m, 1000) if n == 64 else
             (P.kl_small(n, m - n, 50 + i) if i % 2 == 0 else P.slab_qp(n, m // 2, 0, 50 + i)) for i in range(B)]
    sol = cb.BatchedBarrierSolver(cb.pack_problems(probs), None, handle).solve()
    assert np.all(sol.status == 0), sol.status
    for i, pr in enumerate(probs):
        objF, s0 = _oracle(pr)
        o0 = objF.valueAt(s0.x)
        assert abs(sol.objective[i] - o0) <= 1e-8 * max(1.0, abs(o0)), (i, sol.objective[i], o0)
        assert np.linalg.norm(sol.x[i] - s0.x) <= 1e-6 * np.linalg.norm(s0.x)
        assert sol.outer_stages[i] == s0.outer_stages
        # early stages agree exactly; stages with t >= 1e4 terminate at the rounding-noise floor (see
        # tests/test_barrier_gpu.py::_check_solve), so the total is held to a band only
        spun = max(s0.stage_newton_steps) >= 1000 or sol.newton_steps[i] >= 1000   # ||b-Ax|| > 1e-8 spin, noise-decided
        if not spun:
            assert abs(int(sol.newton_steps[i]) - s0.newton_steps) <= max(8, (35 * s0.newton_steps) // 100), (i, sol.newton_steps[i], s0.newton_steps)
        assert abs(sol.dualityGap[i] - s0.dualityGap) <= 1e-12 * s0.dualityGap


def test_batched_agrees_with_large_path(handle):
    """Same problems through the one-problem-at-a-time device path (solver.cu): identical semantics."""
    import cvx_b200 as cb
    probs = [P.batched_problem(i, 64, 128, 2000) for i in range(6)]
    sol = cb.BatchedBarrierSolver(cb.pack_problems(probs), None, handle).solve()
    for i, pr in enumerate(probs):
        s1 = cb.from_dict(pr, "BR", None, handle).solve()
        assert abs(sol.objective[i] - s1.objective) <= 1e-9 * max(1.0, abs(s1.objective))
        assert sol.outer_stages[i] == s1.outer_stages


def test_batched_infeasible_start_is_flagged(handle):
    import cvx_b200 as cb
    probs = [P.batched_problem(i, 64, 128, 3000) for i in range(4)]
    probs[2]["x0"] = probs[2]["x0"] + 50.0
    sol = cb.BatchedBarrierSolver(cb.pack_problems(probs), None, handle).solve()
    assert sol.status[2] == cb._lib.ENOTFEASIBLE
    assert np.all(sol.status[[0, 1, 3]] == 0)


def test_batch_dimension_limits(handle):
    import cvx_b200 as cb
    with pytest.raises(AssertionError):
        cb.BatchedBarrierSolver(cb.pack_problems([P.slab_qp(65, 40, 0, 1)]), None, handle)

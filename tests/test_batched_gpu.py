"""Batched one-CTA-per-problem barrier solver (configs[2] shape n=64, m=128, p in {0,1}) against the
CPU oracle run problem by problem."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def _check_stage_counts(got, want, tag):
    """Per-stage Newton counts against the oracle within the measured band (tests/test_barrier_gpu.py::_band): +-1 for
    stages 0-3, the CPU oracle's own rounding spread afterwards; spins (a stage running to maxIter) from stage 4 on are
    rounding-decided and skipped."""
    from tests.test_barrier_gpu import _band
    band = _band()
    for k, b in enumerate(list(want)[:16]):
        a = int(got[k])
        if a >= 1000 or b >= 1000:
            assert k >= 4, (tag, k, list(got), list(want))
            continue
        tol = band[k] if k < len(band) else band[-1]
        if tol > NOISE_FLOOR_BAND:
            continue      # noise-floor stage (t >= 1e11): checked on the distribution, see _check_noise_floor_stages
        assert abs(a - b) <= tol, (tag, k, list(got), list(want))


# A stage whose measured band exceeds this is at the rounding-noise floor: there the CPU oracle's own variants differ by
# 7 vs 34 steps and 7 vs 1000 (spin) on the same problem (tests/golden/iteration_noise.json), so a per-problem bound says
# nothing; the distributions over many problems must still agree.
NOISE_FLOOR_BAND = 20


def _check_noise_floor_stages(got, want):
    """got, want: (problems x stages) Newton counts.  For the noise-floor stages compare the distributions: medians
    within 2 steps and a similar share of long stalls (> 3x the median)."""
    from tests.test_barrier_gpu import _band
    band = _band()
    for k in range(min(got.shape[1], want.shape[1], len(band))):
        if band[k] <= NOISE_FLOOR_BAND:
            continue
        g, w = got[:, k].astype(float), want[:, k].astype(float)
        if not np.any(w > 0):
            continue
        assert abs(np.median(g) - np.median(w)) <= 2, (k, np.median(g), np.median(w))
        stall_g, stall_w = np.mean(g > 3 * np.median(w)), np.mean(w > 3 * np.median(w))
        assert abs(stall_g - stall_w) <= 0.15, (k, stall_g, stall_w)


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
        _check_stage_counts(sol.stage_newton_steps[i], s0.stage_newton_steps, i)
        assert abs(sol.dualityGap[i] - s0.dualityGap) <= 1e-12 * s0.dualityGap


def test_batched_8192_against_golden_sample(handle):
    """BASELINE.json configs[2] at full size: all B = 8192 problems on the device, the 256 sampled ones against the
    committed oracle results (tests/golden/batched_8192_sample.npz, made by tests/golden/make_batched_golden.py):
    objective 1e-8 relative, x 1e-6 relative, same outer stages, per-stage Newton counts within the measured band;
    every problem of the batch converges and ends strictly feasible with its equality satisfied."""
    import os
    import cvx_b200 as cb
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batched_8192_sample.npz"))
    B = 8192
    probs = [P.batched_problem(i, 64, 128, 1000) for i in range(B)]
    sol = cb.BatchedBarrierSolver(cb.pack_problems(probs), None, handle).solve()
    assert np.all(sol.status == 0), np.nonzero(sol.status)[0][:10]
    for k, i in enumerate(gold["index"]):
        o0 = gold["objective"][k]
        assert abs(sol.objective[i] - o0) <= 1e-8 * max(1.0, abs(o0)), (i, sol.objective[i], o0)
        assert np.linalg.norm(sol.x[i] - gold["x"][k]) <= 1e-6 * np.linalg.norm(gold["x"][k]), i
        assert sol.outer_stages[i] == gold["outer_stages"][k]
        assert abs(sol.dualityGap[i] - gold["dualityGap"][k]) <= 1e-12 * gold["dualityGap"][k]
        _check_stage_counts(sol.stage_newton_steps[i], gold["stage_newton_steps"][k][:gold["outer_stages"][k]], i)
    _check_noise_floor_stages(sol.stage_newton_steps[gold["index"]], gold["stage_newton_steps"])
    # size-independent properties on the whole batch
    for i in range(0, B, 97):
        pr = probs[i]
        assert np.all(pr["G"] @ sol.x[i] * (1 + 3e-16) < pr["ub"]), i
        if pr.get("A") is not None:
            assert abs(float(pr["A"][0] @ sol.x[i]) - pr["b"][0]) < 1e-8, i
    assert np.all(sol.outer_stages == 12) and np.all(sol.stage_newton_steps[:, :12].sum(1) == sol.newton_steps)


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

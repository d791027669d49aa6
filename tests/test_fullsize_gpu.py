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

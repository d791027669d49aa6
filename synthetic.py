"""Seeded synthetic problems of the BASELINE.json shapes and the reference's known-answer problems.

INPUT GENERATION ONLY (numpy, no CUDA, no solver code): bench.py, tools/ and the tests build their inputs here;
the conversion of a problem dict into CPU-oracle objects lives in oracle/problems.py.  The generators mirror the reference's own
random helpers, with seeds added (the reference never seeds):
  Constraints.randomLinearIneqConstraint / linearIneqConstraint   (Constraints.scala:158-176)
  Constraints.randomEqualityConstraint                             (Constraints.scala:205-214)
  ObjectiveFunctions.quadraticObjectiveFunction                    (ObjectiveFunctions.scala:21-34)
  Dist_KL.apply                                                    (Dist_KL.scala:270-315)
  KktTest.testSolutionWithCholFactor / testPositiveDefinite        (src/test/scala/cvx/KktTest.scala:146-272)
Everything returns plain numpy arrays (row-major); a "problem" is a dict:
  kind: 'linear'|'quadratic'|'kl'   objective family, with a (n), r, P (n x n) as applicable
  G (m x n), rvec (m), ub (m)       inequalities  rvec_i + G_i x <= ub_i
  A (p x n), b (p)                  equalities or None
  x0 (n)                            strictly feasible start or None (=> phase I from `xdef`)
  xdef (n)                          pointWhereDefined
"""
from __future__ import annotations

import numpy as np


def slab_lp(n, m_half, p, seed=0, feasible_start=True):
    """C1 / C5 family:  min -c'x  s.t.  -e <= R(x-x0) <= e  (G=[R;-R]),  A x = A x0."""
    rng = np.random.default_rng(seed)
    x0 = np.full(n, 1.0 / n)
    c = rng.uniform(-1, 1, n)
    R = rng.uniform(-1, 1, (m_half, n))
    e = rng.uniform(0.1, 0.2, m_half)
    G = np.vstack([R, -R])
    Rx0 = R @ x0
    ub = np.concatenate([Rx0 + e, -Rx0 + e])
    prob = dict(kind="linear", n=n, a=-c, r=0.0, P=None, G=G, rvec=np.zeros(2 * m_half), ub=ub,
                A=None, b=None, x0=x0 if feasible_start else None, xdef=x0.copy())
    if p > 0:
        A = rng.uniform(-1, 1, (p, n))
        prob["A"], prob["b"] = A, A @ x0
    if not feasible_start:
        # deliberately infeasible point where everything is defined (cf. minDotProduct x0 = 2a)
        prob["xdef"] = x0 + 2.0 * rng.uniform(0.1, 0.2, n)
    return prob


def min_dot_product(a):
    """SimpleOptimizationProblems.minDotProduct (src/test/.../SimpleOptimizationProblems.scala:142-169):
    min -a'x s.t. |x_j| <= |a_j|; optimum x = a; pointWhereDefined 2a (infeasible => phase I)."""
    a = np.asarray(a, float)
    n = a.shape[0]
    G = np.zeros((2 * n, n))
    ub = np.zeros(2 * n)
    for j in range(n):   # Constraints.absoluteValuesBoundedBy: +x_j <= ub_j, -x_j <= ub_j per j
        G[2 * j, j], G[2 * j + 1, j] = 1.0, -1.0
        ub[2 * j] = ub[2 * j + 1] = abs(a[j])
    return dict(kind="linear", n=n, a=-a, r=0.0, P=None, G=G, rvec=np.zeros(2 * n), ub=ub, A=None, b=None,
                x0=None, xdef=2.0 * a, xopt=a.copy())


def min_pNorm(n, p):
    """SimpleOptimizationProblems.min_pNorm (src/test/.../SimpleOptimizationProblems.scala:179-209):
    min sum |x_j|^p  s.t.  x_j >= 0, sum x = 1;  unique optimum x_j = 1/n;  pointWhereDefined = 0 => phase I."""
    return dict(kind="pnorm", n=n, a=None, r=0.0, P=None, pow=float(p), G=-np.eye(n), rvec=np.zeros(n), ub=np.zeros(n),
                A=np.ones((1, n)), b=np.array([1.0]), x0=None, xdef=np.zeros(n), xopt=np.full(n, 1.0 / n))


def rank_one_simplex(n):
    """SimpleOptimizationProblems.rankOneProblemSimplex (src/test/.../SimpleOptimizationProblems.scala:221-255):
    min (a'x)^2 = x'(aa')x  (QuadraticObjectiveFunction with P = aa', i.e. x'Px/2) s.t. x >= 0, sum x = 1,
    a = linspace(1, 2, n); rank-one Hessian; unique optimum e_1; pointWhereDefined 1/n => phase I."""
    a = np.linspace(1.0, 2.0, n)
    xopt = np.zeros(n)
    xopt[0] = 1.0
    return dict(kind="quadratic", n=n, a=np.zeros(n), r=0.0, P=np.outer(a, a), G=-np.eye(n), rvec=np.zeros(n),
                ub=np.zeros(n), A=np.ones((1, n)), b=np.array([1.0]), x0=None, xdef=np.full(n, 1.0 / n), xopt=xopt)


def norm_squared_free_variables(n):
    """SimpleOptimizationProblems.normSquaredWithFreeVariables (:308-340): min ||x||^2/2 s.t. x_1 <= -1 -- one
    constraint, n-1 variables it does not depend on (phase I sees a rank-one barrier Hessian); optimum (-1,0,..,0);
    pointWhereDefined = 1 (infeasible) => phase I."""
    G = np.zeros((1, n))
    G[0, 0] = 1.0
    xopt = np.zeros(n)
    xopt[0] = -1.0
    return dict(kind="quadratic", n=n, a=np.zeros(n), r=0.0, P=np.eye(n), G=G, rvec=np.zeros(1), ub=np.array([-1.0]),
                A=None, b=None, x0=None, xdef=np.ones(n), xopt=xopt)


def jopt_p1(n):
    """SimpleOptimizationProblems.joptP1 (:347-377): min sum(x) s.t. ||x||^2/2 <= 1/2 (Constraints.oneHalfNorm2BoundedBy,
    Constraints.scala:299-309: value x'x/2, gradient x, Hessian I -- the quadratic constraint r + a'x + x'Px/2 with
    P = I, a = 0, r = 0); optimum x_j = -1/sqrt(n); pointWhereDefined = 2 (infeasible) => phase I."""
    quad = [dict(P=np.eye(n), a=np.zeros(n), r=0.0, ub=0.5)]
    return dict(kind="linear", n=n, a=np.ones(n), r=0.0, P=None, G=np.zeros((0, n)), rvec=np.zeros(0), ub=np.zeros(0),
                quad=quad, A=None, b=None, x0=None, xdef=np.full(n, 2.0), xopt=np.full(n, -1.0 / np.sqrt(n)))


def jopt_p2():
    """SimpleOptimizationProblems.joptP2 (:384-414; docs/OptimizerExamples.pdf example 1.5): min x'Px/2, P = [[1,.4],[.4,1]],
    s.t. x >= 0, x_1 + x_2 = 1; optimum (1/2, 1/2); pointWhereDefined = (2, 2) (infeasible) => phase I."""
    P_ = np.array([[1.0, 0.4], [0.4, 1.0]])
    return dict(kind="quadratic", n=2, a=np.zeros(2), r=0.0, P=P_, G=-np.eye(2), rvec=np.zeros(2), ub=np.zeros(2),
                A=np.ones((1, 2)), b=np.array([1.0]), x0=None, xdef=np.full(2, 2.0), xopt=np.array([0.5, 0.5]))


def probability_simplex_problem(n):
    """SimpleOptimizationProblems.probabilitySimplexProblem (:425-453): min (sum(x) - 1)^2 / 2 =
    QuadraticObjectiveFunction(n, 0.5, -1, 11') over x >= 0; every point of the probability simplex is a minimiser
    (objective 0; the listed one is 1/n); pointWhereDefined = 2 (feasible, but phase I runs: withoutFeasiblePoint)."""
    return dict(kind="quadratic", n=n, a=-np.ones(n), r=0.5, P=np.ones((n, n)), G=-np.eye(n), rvec=np.zeros(n),
                ub=np.zeros(n), A=None, b=None, x0=None, xdef=np.full(n, 2.0), xopt=np.full(n, 1.0 / n))


def distance_from_origin(n, sliced=False):
    """SimpleOptimizationProblems.distanceFromOrigin0 / distanceFromOrigin1 (:462-552): in R^{n+1}, min ||x||^2/2 on
    the ball ||x - 2 e_{n+1}||^2/2 <= 1/2 (QuadraticConstraint(ub = 0, r = 1.5, a = -2 e_{n+1}, P = I)); `sliced` adds the
    2n linear constraints -(e_j + e_{n+1})'x <= -1 exactly as the reference builds them (its second constraint of each
    pair re-uses `-a` instead of `-b`, so every row appears twice); optimum e_{n+1}; pointWhereDefined = 0 => phase I."""
    d = n + 1
    e = np.zeros(d)
    e[n] = 1.0
    quad = [dict(P=np.eye(d), a=-2.0 * e, r=1.5, ub=0.0)]
    rows = []
    if sliced:
        for j in range(n):
            a = e.copy()
            a[j] += 1.0
            rows += [-a, -a]
    G = np.array(rows) if rows else np.zeros((0, d))
    m = G.shape[0]
    return dict(kind="quadratic", n=d, a=np.zeros(d), r=0.0, P=np.eye(d), G=G, rvec=np.zeros(m), ub=np.full(m, -1.0),
                quad=quad, A=None, b=None, x0=None, xdef=np.zeros(d), xopt=e)


def kl_random(n, m_h, p_extra, seed=0):
    """C2 family via Dist_KL.apply semantics: KL objective, m_h rows Hx<=u plus n positivity rows,
    p_extra rows A x = r plus the sum-to-one row (stacked last); start 1/n => phase I."""
    rng = np.random.default_rng(seed)
    z = rng.normal(0.0, 0.5, n)
    qs = np.exp(z - z.max())
    qs /= qs.sum()
    H = rng.uniform(-1, 1, (m_h, n))
    u = H @ qs + rng.uniform(0.05, 0.15, m_h)
    G = np.vstack([H, -np.eye(n)])
    ub = np.concatenate([u, np.zeros(n)])
    ones = np.ones((1, n))
    if p_extra > 0:
        A0 = rng.uniform(-1, 1, (p_extra, n))
        A = np.vstack([A0, ones])
        b = np.concatenate([A0 @ qs, [1.0]])
    else:
        A, b = ones, np.array([1.0])
    return dict(kind="kl", n=n, a=None, r=0.0, P=None, G=G, rvec=np.zeros(m_h + n), ub=ub, A=A, b=b,
                x0=None, xdef=np.full(n, 1.0 / n), qstar=qs)


def kl_1A(n):
    """OptimizationProblems.kl_1A (src/test/.../OptimizationProblems.scala:167-244): P(A)>=0.36, P(B)<=0.1."""
    assert n > 9 and n % 2 == 0
    I_A = (np.arange(n) < 3).astype(float)
    I_B = (np.arange(n) >= n // 2).astype(float)
    H = np.vstack([-I_A, I_B])
    u = np.array([-0.36, 0.1])
    G = np.vstack([H, -np.eye(n)])
    ub = np.concatenate([u, np.zeros(n)])
    if n <= 15:
        xopt = np.where(np.arange(n) < n // 2, 1.8 / n, 0.2 / n)
    else:
        j = np.arange(n)
        xopt = np.where(j < 3, 0.12, np.where(j >= n // 2, 0.2 / n, 1.08 / (n - 6)))
    return dict(kind="kl", n=n, a=None, r=0.0, P=None, G=G, rvec=np.zeros(n + 2), ub=ub,
                A=np.ones((1, n)), b=np.array([1.0]), x0=None, xdef=np.full(n, 1.0 / n), xopt=xopt)


def kl_2A(n):
    """OptimizationProblems.kl_2A (:331-369): P(A)=0.36, P(B)=0.1 as equalities."""
    assert n > 9 and n % 2 == 0
    j = np.arange(n)
    I_A = (j < 3).astype(float)
    I_B = (j >= n // 2).astype(float)
    A = np.vstack([I_A, I_B, np.ones(n)])
    b = np.array([0.36, 0.1, 1.0])
    xopt = np.where(j < 3, 0.12, np.where(j >= n // 2, 0.2 / n, 1.08 / (n - 6)))
    return dict(kind="kl", n=n, a=None, r=0.0, P=None, G=-np.eye(n), rvec=np.zeros(n), ub=np.zeros(n),
                A=A, b=b, x0=None, xdef=np.full(n, 1.0 / n), xopt=xopt)


def infeasible_kl_1(n):
    """OptimizationProblems.infeasible_kl_1 (:379-405): P(A)>=0.51 and P(B)>=0.51, disjoint A, B."""
    j = np.arange(n)
    I_A = (j < 3).astype(float)
    I_B = (j >= n // 2).astype(float)
    # ConstraintSets.probAB: positivity constraints first, then the two probability constraints
    G = np.vstack([-np.eye(n), -I_A, -I_B])
    ub = np.concatenate([np.zeros(n), [-0.51, -0.51]])
    return dict(kind="kl", n=n, a=None, r=0.0, P=None, G=G, rvec=np.zeros(n + 2), ub=ub,
                A=np.ones((1, n)), b=np.array([1.0]), x0=None, xdef=np.full(n, 1.0 / n))


def slab_qp(n, m_half, p, seed=0, scale=True):
    """C3(QP half) / C4 family:  min 0.5||R(x-xc)||^2  s.t. slab G=[R2;-R2] around x0,  A x = A x0."""
    rng = np.random.default_rng(seed)
    x0 = np.full(n, 1.0 / n)
    sc = 1.0 / np.sqrt(n) if scale else 1.0
    R = rng.uniform(-1, 1, (n, n)) * sc
    xc = x0 + rng.normal(0, 1, n)
    Rxc = R @ xc
    P = R.T @ R
    P = (P + P.T) * 0.5
    a = -(R.T @ Rxc)
    r = 0.5 * float(Rxc @ Rxc)
    R2 = rng.uniform(-1, 1, (m_half, n))
    e = rng.uniform(0.1, 0.2, m_half) * (np.sqrt(n) if scale else 1.0)
    G = np.vstack([R2, -R2])
    R2x0 = R2 @ x0
    ub = np.concatenate([R2x0 + e, -R2x0 + e])
    prob = dict(kind="quadratic", n=n, a=a, r=r, P=P, G=G, rvec=np.zeros(2 * m_half), ub=ub, A=None, b=None,
                x0=x0, xdef=x0.copy())
    if p > 0:
        A = rng.uniform(-1, 1, (p, n))
        prob["A"], prob["b"] = A, A @ x0
    return prob


def kl_small(n, m_h, seed=0):
    """C3(KL half): KL objective, m_h rows Hx<=u + n positivity rows, p=1 (sum to one), feasible
    start given (the softmax point q*, strictly inside by construction)."""
    pr = kl_random(n, m_h, 0, seed)
    pr["x0"] = pr["qstar"].copy()
    return pr


def batched_problem(i, n=64, m=128, base_seed=1000):
    """C3: problem i of the batch; even i -> KL (p=1), odd i -> QP (p=0)."""
    if i % 2 == 0:
        return kl_small(n, m - n, base_seed + i)
    return slab_qp(n, m // 2, 0, base_seed + i, scale=True)


def batched_problem_phase1(i, n=63, m=126, base_seed=7000):
    """Phase-I variant of C3: the same two families without a feasible start -- even i: KL problem of
    Dist_KL.apply form (start 1/n is defined but violates the H x <= u rows; one equality: the feasibility problem has n + 1
    variables and m + 2 rows), odd i: slab QP started outside its slab.  n = 63, m = 126 is the largest shape whose
    feasibility problem still fits the batched kernel's 64 x 128 layout."""
    if i % 2 == 0:
        return kl_random(n, m - n, 0, base_seed + i)          # x0 = None, xdef = 1/n
    pr = slab_qp(n, m // 2, 0, base_seed + i, scale=True)
    rng = np.random.default_rng(base_seed + 100000 + i)
    pr["xdef"] = pr["x0"] + rng.uniform(0.5, 1.0, n)
    pr["x0"] = None
    return pr


def lin_quad_set(n, m_lin, m_quad, p=0, seed=0, objective="quadratic", feasible_start=True):
    """FeasibilityTests / ConstraintSets.randomConstraintSet design (src/test/scala/cvx/FeasibilityTests.scala:105-117,
    ConstraintSets.scala:67-89; Constraints.randomLinearIneqConstraint / randomQuadraticConstraint,
    Constraints.scala:158-204): m_lin linear and m_quad convex quadratic constraints, all strictly satisfied
    at a known point x0, optional random equalities through x0; objective: convex quadratic or linear."""
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-1, 1, n)
    G = rng.uniform(-1, 1, (m_lin, n))
    ub = G @ x0 + rng.uniform(0.5, 1.5, m_lin)
    quad = []
    for k in range(m_quad):
        B = rng.uniform(-1, 1, (n, n)) / np.sqrt(n)
        Pk = B.T @ B
        Pk = (Pk + Pk.T) * 0.5
        ak = rng.uniform(-1, 1, n)
        rk = float(rng.uniform(-1, 1))
        val = rk + ak @ x0 + 0.5 * x0 @ Pk @ x0
        quad.append(dict(P=Pk, a=ak, r=rk, ub=float(val + rng.uniform(0.5, 1.5))))
    if objective == "quadratic":
        R = rng.uniform(-1, 1, (n, n)) / np.sqrt(n)
        P = R.T @ R + 0.1 * np.eye(n)
        P = (P + P.T) * 0.5
        a = rng.uniform(-1, 1, n)
        prob = dict(kind="quadratic", n=n, a=a, r=0.0, P=P)
    else:
        prob = dict(kind="linear", n=n, a=rng.uniform(-1, 1, n), r=0.0, P=None)
    prob.update(G=G, rvec=np.zeros(m_lin), ub=ub, quad=quad, A=None, b=None, x0=x0 if feasible_start else None,
                xdef=x0.copy() if feasible_start else x0 + rng.uniform(2.0, 3.0, n))
    if p > 0:
        A = rng.uniform(-1, 1, (p, n))
        prob["A"], prob["b"] = A, A @ x0
    return prob


# ---------------- planted KKT systems (KktTest.scala) ----------------


def kkt_planted_chol(n, p, seed=0):
    """KktTest.testSolutionWithCholFactor(n,p,...) :146-167:  L = tril(U(-5,5)) + sqrt(n) I,
    A = U(0,1)^{p x n} + I, x ~ U(-1,1), w ~ U(-2,2); q = -(Hx + A'w), b = A x."""
    rng = np.random.default_rng(seed)
    L = np.tril(rng.uniform(-5, 5, (n, n)))
    L[np.arange(n), np.arange(n)] += np.sqrt(n)
    A = rng.uniform(0, 1, (p, n))
    A[np.arange(p), np.arange(p)] += 1.0
    x = rng.uniform(-1, 1, n)
    w = rng.uniform(-2, 2, p)
    H = L @ L.T
    H = (H + H.T) * 0.5
    return dict(L=L, H=H, A=A, x=x, w=w, q=-(H @ x + A.T @ w), b=A @ x)


def kkt_planted_pd(n, p, seed=0):
    """KktTest.testPositiveDefinite(n,p,...) :246-262: H = sym(LL'), A = U(-5,5) + 20 I."""
    rng = np.random.default_rng(seed)
    L = np.tril(rng.uniform(-5, 5, (n, n)))
    L[np.arange(n), np.arange(n)] += np.sqrt(n)
    M = L @ L.T
    H = (M + M.T) * 0.5
    A = rng.uniform(-5, 5, (p, n))
    A[np.arange(p), np.arange(p)] += 20.0
    x = rng.uniform(-1, 1, n)
    w = rng.uniform(-2, 2, p)
    return dict(H=H, A=A, x=x, w=w, q=-(H @ x + A.T @ w), b=A @ x)


def newton_step_inputs(n, m_half, p, seed=0):
    """One barrier Newton step at a strictly feasible random iterate of a slab problem: returns
    the problem, an iterate x (not x0) and a barrier parameter t."""
    prob = slab_qp(n, m_half, p, seed)
    rng = np.random.default_rng(seed + 7919)
    d = rng.normal(0, 1, n)
    Gd = prob["G"] @ d
    slack = prob["ub"] - prob["G"] @ prob["x0"]
    smax = float(np.min(np.where(Gd > 0, slack / np.where(Gd > 0, Gd, 1.0), np.inf)))
    x = prob["x0"] + 0.5 * smax * d
    return prob, x, 10.0


# ---------------- objectives given as closures (seam B: Hessian assembled on the host) ----------------


class PowerFunction:
    """Type1Function.powerFunction(A, alpha, q) (src/test/scala/cvx/Type1Function.scala:25-78, docs/cvx_notes example 3.1):
    f(x) = sum_j alpha_j ((a_j . x)^2)^q, a_j = row_j(A), q > 1; global minimum 0 on ker(A).  Plain numpy closures --
    value, gradient, Hessian -- as an objective without a closed-form family looks to the solvers; `hessianColumns`
    produces a block of columns of  A' diag(alpha o phi''(Ax)) A  so that its upload can overlap the next block."""

    def __init__(self, A, alpha, q):
        self.A, self.alpha, self.q = np.asarray(A, float), np.asarray(alpha, float), float(q)
        assert self.q > 1, "q=%s is not > 1." % q
        assert self.A.shape[0] <= self.A.shape[1] and self.alpha.shape[0] == self.A.shape[0] and np.all(self.alpha > 0)
        self.dim = self.A.shape[1]
        self._w_at, self._w = None, None

    def _u(self, x):
        u = self.A @ x
        return np.where(np.abs(u) < 1e-14, 0.0, u)

    def valueAt(self, x):
        u = self._u(x)
        return float(self.alpha @ np.power(u * u, self.q))

    def gradientAt(self, x):
        u = self._u(x)
        d = np.where(u == 0.0, 0.0, (2 * self.q) * np.power(u * u, self.q - 0.5)) * np.sign(u)
        return self.A.T @ (self.alpha * d)

    def _weights(self, x):
        if self._w_at is None or not np.array_equal(self._w_at, x):
            u = self._u(x)
            with np.errstate(divide="ignore", invalid="ignore"):
                d2 = np.where(u == 0.0, 0.0, (2 * self.q) * (2 * self.q - 1) * np.power(u * u, self.q - 1))
            self._w_at, self._w = np.array(x, copy=True), self.alpha * d2
        return self._w

    def hessianAt(self, x):
        w = self._weights(x)
        H = self.A.T @ (self.A * w[:, None])
        return (H + H.T) / 2

    def hessianColumns(self, x, j0, j1, out):
        w = self._weights(x)
        np.matmul(self.A.T, self.A[:, j0:j1] * w[:, None], out=out)

    def isMinimizer(self, x, tol):
        return float(np.linalg.norm(self.A @ x)) < tol


def power_problem(A, alpha, q):
    """OptimizationProblems.powerProblem (src/test/scala/cvx/OptimizationProblems.scala:61-90): unconstrained, start
    x_j = -10 + j sqrt(n), known minimum value 0 on ker(A)."""
    f = PowerFunction(A, alpha, q)
    n = f.dim
    return f, np.array([-10.0 + j * np.sqrt(n) for j in range(n)])


def power_problems():
    """OptimizationProblems.powerProblems (:112-125): A = I_2 and A = [[1,0],[1,1]], alpha = (1,1), q = 2."""
    return [power_problem(np.eye(2), np.ones(2), 2.0), power_problem(np.array([[1.0, 0.0], [1.0, 1.0]]), np.ones(2), 2.0)]


def random_power_problem(dim, m, q, seed=0):
    """Type1Function.randomPowerFunction (Type1Function.scala:85-95) with a seed: A = U(0,1)^{m x dim} + I, alpha ~ U(0,1)."""
    rng = np.random.default_rng(seed)
    A = rng.uniform(0, 1, (m, dim))
    A[np.arange(m), np.arange(m)] += 1.0
    alpha = rng.uniform(0.05, 1.0, m)
    f = PowerFunction(A, alpha, q)
    return f, rng.uniform(-1.0, 1.0, dim)

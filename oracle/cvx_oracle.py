"""CPU restatement (oracle) of the interior-point Newton/KKT hot path of spyqqqdia/cvx.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cvx_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker / the thing timed as the CPU baseline.

Parity status: the reference (Scala 2.11 + Breeze 0.10 + netlib-java) cannot be built or run in
the build container (no JVM, no jars, no network) and its own tests hold NO golden vectors and no
seeds.  This restatement is therefore pinned by
  (i)  the reference's own test *designs*, with seeds added: planted-solution KKT systems
       (src/test/scala/cvx/KktTest.scala:117-184,197-272), triangular / Cholesky solves
       (src/test/scala/cvx/MatrixUtilsTests.scala:36-198), Ruiz with a zero row (:16-26), and
  (ii) the analytic optima of the reference's known-answer problems (minDotProduct, kl_1, kl_2,
       min_pNorm, infeasible_kl_1; src/test/scala/cvx/SimpleOptimizationProblems.scala:142-209,
       src/test/scala/cvx/OptimizationProblems.scala:136-141,249-251,379-405).
At the 1e-10 level the Newton direction itself is "parity unpinned" by reference outputs.

All file:line citations are relative to /root/reference/src/main/scala/cvx/ unless noted.
LAPACK routines are the same ones Breeze 0.10 reaches through netlib-java: dpotrf, dtrtrs,
dsyev*, dgesdd, dgemm (scipy/OpenBLAS here).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
import scipy.linalg as sla

# --------------------------------------------------------------------------------------
# exceptions (LinSolveException.scala, UnsolvableSystemException.scala,
# LineSearchFailedException.scala, InfeasibleProblemException.scala, Breeze exceptions)
# --------------------------------------------------------------------------------------


class LinSolveException(Exception):
    pass


class UnsolvableSystemException(Exception):
    pass


class LineSearchFailedException(Exception):
    pass


class InfeasibleProblemException(Exception):
    pass


class NotConvergedException(Exception):
    """Breeze NotConvergedException (dpotrf info>0, or line-search Breakdown)."""


class MatrixNotSymmetricException(Exception):
    """Breeze 0.10 requireSymmetricMatrix: exact A(i,j)==A(j,i) check inside cholesky/eigSym."""


class NotStrictlyFeasible(ValueError):
    """IllegalArgumentException thrown by the barrier function family, BarrierSolver.scala:284."""


# --------------------------------------------------------------------------------------
# SolverParams.scala:24-46
# --------------------------------------------------------------------------------------


@dataclass
class SolverParams:
    maxIter: int = 1000
    alpha: float = 0.04
    beta: float = 0.8
    tolSolver: float = 1e-8
    tolEqSolve: float = 1e-1
    tolFeas: float = 1e-7
    delta: float = 1e-6

    @staticmethod
    def standardParams() -> "SolverParams":
        return SolverParams()


# Solution.scala:32-43
@dataclass
class Solution:
    x: np.ndarray
    lam: Optional[np.ndarray] = None
    nu: Optional[np.ndarray] = None
    newtonDecrement: Optional[float] = None
    dualityGap: Optional[float] = None
    equalityGap: Optional[float] = None
    normGrad: Optional[float] = None
    normDualResidual: Optional[float] = None
    iter: int = 0
    maxedOut: bool = False
    # extras (not in the reference record): bookkeeping for parity tests
    newton_steps: int = 0
    stage_newton_steps: List[int] = field(default_factory=list)
    outer_stages: int = 0
    linesearch_trials: List[int] = field(default_factory=list)


# OptimizationState.scala:22-29
@dataclass
class OptimizationState:
    normGradient: Optional[float]
    newtonDecrement: Optional[float]
    dualityGap: Optional[float]
    equalityGap: Optional[float]
    objectiveFunctionValue: float
    normDualResidual: Optional[float] = None


def _ieee_div(a, b):
    """a / b with JVM (IEEE-754) semantics: a zero divisor gives +-Infinity or NaN instead of Python's exception
    (PrimalDualSolver.scala:576,612: `t = mu*numIneqs/dualityGap` with a surrogate gap of exactly 0.0)."""
    a, b = float(a), float(b)
    if b != 0.0:
        return a / b
    if a == 0.0 or a != a:
        return float("nan")
    return math.copysign(float("inf"), a) * math.copysign(1.0, b)


DOUBLE_MAX = float(np.finfo(np.float64).max)

# ======================================================================================
# MatrixUtils.scala
# ======================================================================================


def breeze_cholesky(X: np.ndarray) -> np.ndarray:
    """Breeze 0.10 `cholesky`: exact symmetry check, then LAPACK dpotrf('L') on the lower triangle;
    info>0 -> NotConvergedException.  Call sites MatrixUtils.scala:456,460, KKTSystem.scala:140."""
    if X.shape[0] != X.shape[1]:
        raise ValueError("not square")
    if not np.array_equal(X, X.T):
        raise MatrixNotSymmetricException()
    c, info = sla.lapack.dpotrf(X, lower=1, clean=1, overwrite_a=0)
    if info > 0:
        raise NotConvergedException("dpotrf info=%d" % info)
    assert info == 0
    return c


def ruizEquilibrate(H: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """MatrixUtils.scala:240-268.  Jacobi sweeps in the l2 norm; <= 20 sweeps, stop at rho <= 1e-6."""
    n = H.shape[0]
    assert H.shape[1] == n
    d = np.ones(n)
    rho = 1.0
    it = 0
    while it < 20 and rho > 1e-6:
        Q = np.outer(d, d) * H
        u = np.sqrt(np.sqrt(np.sum(Q * Q, axis=1)))  # sqrt(norm(row_i(Q)))
        v = np.where(u > 0, 1.0 / np.where(u > 0, u, 1.0), 1.0)
        d = d * v
        rho = float(np.max(np.abs(1.0 - u))) if n > 0 else 0.0
        it += 1
    ruizEquilibrate.last_sweeps = it
    return d, np.outer(d, d) * H


ruizEquilibrate.last_sweeps = 0


def triangularSolve(A: np.ndarray, Ltype: str, B: np.ndarray) -> np.ndarray:
    """MatrixUtils.scala:362-376 -> LAPACK dtrtrs (triangle copied first)."""
    assert Ltype in ("L", "U")
    Q = np.tril(A) if Ltype == "L" else np.triu(A)
    x, info = sla.lapack.dtrtrs(Q, B, lower=1 if Ltype == "L" else 0, trans=0, unitdiag=0)
    return x


def forwardSolve(L: np.ndarray, b: np.ndarray) -> np.ndarray:
    """MatrixUtils.scala:383-403 (scalar row-oriented loop; asserts a non-zero diagonal)."""
    n = L.shape[0]
    assert b.shape[0] == n
    assert np.all(np.abs(np.diag(L)) > 0), "Singular lower triangular matrix L: zero on the diagonal"
    if n <= 256:  # literal row-oriented dot-product order
        x = np.zeros(n)
        for i in range(n):
            x[i] = (b[i] - np.dot(L[i, :i], x[:i])) / L[i, i]
        return x
    return sla.solve_triangular(L, b, lower=True, check_finite=False)


def backSolve(U: np.ndarray, b: np.ndarray) -> np.ndarray:
    """MatrixUtils.scala:410-430."""
    n = U.shape[0]
    assert b.shape[0] == n
    assert np.all(np.abs(np.diag(U)) > 0), "Singular upper triangular matrix U: zero on the diagonal"
    if n <= 256:
        x = np.zeros(n)
        for i in range(n - 1, -1, -1):
            x[i] = (b[i] - np.dot(U[i, i + 1:], x[i + 1:])) / U[i, i]
        return x
    return sla.solve_triangular(U, b, lower=False, check_finite=False)


def relativeSize(a: np.ndarray, b: np.ndarray, tol: float) -> float:
    """MatrixUtils.scala:437-442."""
    nb = float(np.linalg.norm(b))
    f = tol if nb < tol else tol + nb
    return float(np.linalg.norm(a)) / f


def regularizedCholesky(Q: np.ndarray) -> np.ndarray:
    """MatrixUtils.scala:452-461."""
    Qd = Q + np.eye(Q.shape[0]) * 1e-10
    regularizedCholesky.last_regularized = False
    try:
        C = breeze_cholesky(Q)
    except Exception:
        regularizedCholesky.last_regularized = True
        C = breeze_cholesky(Qd)
    minD = float(np.min(np.diag(C)))
    if minD > 1e-7:
        return C
    regularizedCholesky.last_regularized = True
    return breeze_cholesky(Qd)


regularizedCholesky.last_regularized = False


def choleskySolve(H: np.ndarray, b: np.ndarray, tol: float) -> np.ndarray:
    """MatrixUtils.scala:468-516."""
    n = H.shape[0]
    assert H.shape[1] == n and b.shape[0] == n
    d, Q = ruizEquilibrate(H)
    L = regularizedCholesky(Q)
    w = forwardSolve(L, d * b)
    u = backSolve(L.T, w)
    x = d * u
    relErr = relativeSize(H @ x - b, b, tol)
    if relErr > tol:
        raise LinSolveException("choleskySolve: error exceeds tolerance: %g" % relErr)
    return x


def checkSymmetric(Q: np.ndarray, tol: float) -> bool:
    """MatrixUtils.scala:207-211."""
    diff = Q - Q.T
    return math.sqrt(float(np.sum(diff * diff))) < tol


def diagonalizationSolve(A, U, d, V, b, tol) -> np.ndarray:
    """MatrixUtils.scala:603-707 including defect D3 (the inner `val relError` shadows the outer
    one, so once the regularisation loop is entered it runs 18 times and then always throws)."""
    n = U.shape[0]
    nz = np.abs(d) > 0
    a = np.where(nz, U.T @ b, 0.0)
    b0 = U @ a
    relDist = relativeSize(b - b0, b, tol)
    if relDist > tol:
        raise UnsolvableSystemException("min_x||Ax-b||/||b|| = %g > tol" % relDist)
    z = np.where(nz, (U.T @ b) / np.where(nz, d, 1.0), 0.0)
    w = V @ z
    relError = relativeSize(A @ w - b, b, tol)
    if relError > tol:
        # D3: 18 futile tries, then the stale outer relError (> tol) throws.
        raise UnsolvableSystemException("system not solvable within tolerance (D3)")
    return w


def svdSolve(A, b, tol) -> np.ndarray:
    """MatrixUtils.scala:712-730 (Breeze svd returns V', :726)."""
    u, s, vt = sla.svd(A, lapack_driver="gesdd")
    return diagonalizationSolve(A, u, s, vt.T, b, tol)


def symSolve(A, b, tol) -> np.ndarray:
    """MatrixUtils.scala:734-751 (eigSym -> dsyev; requires exact symmetry in Breeze 0.10)."""
    if not np.array_equal(A, A.T):
        raise MatrixNotSymmetricException()
    lam, evs = sla.eigh(A)
    return diagonalizationSolve(A, evs, lam, evs, b, tol)


# ======================================================================================
# KKTSystem.scala
# ======================================================================================


@dataclass
class KKTInfo:
    path: int = 0          # 0 = solvePD(H), 1 = solvePD(H+A'A), 2 = kktSymSolve
    regularized: bool = False
    ruiz_sweeps: int = 0
    err1: float = 0.0
    err2: float = 0.0


# "reference": KKTSystem.solveWithCholFactor as written.  "one_trsm": the device's as-if variant of the block elimination
#  -- NOT the reference's arithmetic; it exists only so that tests and tools can show that a
# disagreement between device and oracle is exactly this deviation (tools/gpu_fuzz.py, tests/test_oracle_cpu.py).
BLOCK_ELIMINATION = "reference"


def _solveWithCholFactor_one_trsm(L, A, q, b, tol, info=None):
    """The device's formulation (kkt.cu): Y = L^-1 [A', q], S = Yp'Yp (positive semidefinite by construction), z = -(b + Yp'yq),
    x = -L^-T (yq + Yp w).  Same quantities as KKTSystem.scala:116-139 in exact arithmetic; the reference forms
    R = A (L^-T L^-1 A') and symmetrises it, which loses definiteness in floating point once cond(H) reaches ~1e20."""
    n = L.shape[1]
    p = A.shape[0]
    B = np.zeros((n, p + 1))
    B[:, :p] = A.T
    B[:, p] = q
    Y = triangularSolve(L, "L", B)
    Yp, yq = Y[:, :p], Y[:, p]
    S = Yp.T @ Yp
    S = (S + S.T) * 0.5
    K = breeze_cholesky(S)
    z = -(b + Yp.T @ yq)
    u = forwardSolve(K, z)
    w = backSolve(K.T, u)
    x = -triangularSolve(L.T, "U", (yq + Yp @ w)[:, None])[:, 0]
    Hx = L @ (L.T @ x)
    err1 = relativeSize(Hx + A.T @ w + q, -q, tol)
    err2 = relativeSize(A @ x - b, b, tol)
    if info is not None:
        info.err1, info.err2 = err1, err2
    if err1 > tol or err2 > tol:
        raise LinSolveException("Error in solution exceeds tolerance.")
    return x, w


def solveWithCholFactor(L, A, q, b, tol, info: Optional[KKTInfo] = None):
    """KKTSystem.scala:99-167."""
    if BLOCK_ELIMINATION == "one_trsm":
        return _solveWithCholFactor_one_trsm(L, A, q, b, tol, info)
    n = L.shape[1]
    assert L.shape[0] == n and A.shape[1] == n
    p = A.shape[0]
    B = np.zeros((n, p + 1))
    B[:, :p] = A.T
    B[:, p] = q
    Y = triangularSolve(L, "L", B)
    X = triangularSolve(L.T, "U", Y)
    Hinv_At = X[:, :p]
    Hinv_q = X[:, p]
    R = A @ Hinv_At
    S = (R + R.T) * 0.5
    K = breeze_cholesky(S)               # plain cholesky: throws if S is not PD
    z = -(b + A @ Hinv_q)
    u = forwardSolve(K, z)
    w = backSolve(K.T, u)
    x = -(Hinv_q + Hinv_At @ w)
    Ltx = L.T @ x
    Hx = L @ Ltx
    err1 = relativeSize(Hx + A.T @ w + q, -q, tol)
    err2 = relativeSize(A @ x - b, b, tol)
    if info is not None:
        info.err1, info.err2 = err1, err2
    if err1 > tol or err2 > tol:
        raise LinSolveException("Error in solution exceeds tolerance.")
    return x, w


def blockSolve(H, A, q, b, tol, info=None):
    """KKTSystem.scala:178-190."""
    L = regularizedCholesky(H)
    if info is not None:
        info.regularized = regularizedCholesky.last_regularized
    return solveWithCholFactor(L, A, q, b, tol, info)


def solvePD(H, A, q, b, tol, info=None):
    """KKTSystem.scala:200-246."""
    n = H.shape[1]
    assert H.shape[0] == n and A.shape[1] == n
    d, Q = ruizEquilibrate(H)
    if info is not None:
        info.ruiz_sweeps = ruizEquilibrate.last_sweeps
    B = A * d[None, :]
    Dq = d * q
    y, w = blockSolve(Q, B, Dq, b, tol, info)
    return d * y, w


def kktMatrix(H, A):
    """KKTSystem.scala:253-260."""
    p = A.shape[0]
    return np.block([[H, A.T], [A, np.zeros((p, p))]])


def kktSymSolve(H, A, g, r, tol):
    """KKTSystem.scala:283-310."""
    n = H.shape[0]
    q = np.concatenate([-g, r])
    M = kktMatrix(H, A)
    w = symSolve(M, q, tol)
    return w[:n], w[n:n + A.shape[0]]


def kkt_solve(H, A, q, b, tol, info: Optional[KKTInfo] = None):
    """KKTSystem.solve, KKTSystem.scala:43-66 (`delta` is never used, D5)."""
    n = H.shape[1]
    assert H.shape[0] == n, "Matrix M not square"
    assert A.shape[1] == n
    try:
        if info is not None:
            info.path = 0
        return solvePD(H, A, q, b, tol, info)
    except Exception:
        try:
            if info is not None:
                info.path = 1
            K = H + A.T @ A
            z = q - A.T @ b
            return solvePD(K, A, z, b, tol, info)
        except Exception:
            if info is not None:
                info.path = 2
            return kktSymSolve(H, A, q, b, tol)


def solveUnderdetermined(A, b):
    """MatrixUtils.solveUnderdetermined (MatrixUtils.scala:536-550): A' = QR (Breeze qr = LAPACK dgeqrf + dorgqr, full
    Q), F = Q(:, m..n-1), y = forwardSolve(R', b), z0 = Q(:, 0..m-1) y.  Returns (z0, F)."""
    import scipy.linalg as sla
    m, n = A.shape
    Q, R = sla.qr(A.T, mode="full")
    F = Q[:, m:n]
    y = forwardSolve(R[:m, :].T, b)
    return Q[:, :m] @ y, F


def affineTransformedProblem(objF, cnts, z0, F):
    """x = z0 + F u for the closed-form families: LinearConstraint.affineTransformed (LinearConstraint.scala:46-52),
    QuadraticConstraint.affineTransformed (QuadraticConstraint.scala:54-64), ObjectiveFunction.affineTransformed
    (ObjectiveFunction.scala:26-40) specialised to linear / quadratic objectives; starting points mapped with
    SolutionSpace.parameter (F'(x - z0)).  Returns (objF_u, cnts_u)."""
    par = lambda x: None if x is None else F.T @ (np.asarray(x, float) - z0)
    sym = lambda M: (M + M.T) / 2          # Breeze's cholesky / eigSym insist on exact symmetry
    quad = [QuadCnt(sym(F.T @ q.P @ F), F.T @ (q.a + q.P @ z0), q.r + float(q.a @ z0) + float(z0 @ (q.P @ z0)) / 2, q.ub)
            for q in cnts.quad]
    c = ConstraintSet(cnts.G @ F, cnts.r + cnts.G @ z0, cnts.ub, quad, par(cnts.pointWhereDefined))
    if cnts.feasiblePoint is not None:
        c.feasiblePoint = par(cnts.feasiblePoint)
    if objF.kind == "linear":
        o = LinearObjective(F.T @ objF.a, objF.r + float(objF.a @ z0))
    elif objF.kind == "quadratic":
        Pu = F.T @ objF.P @ F
        o = QuadraticObjective((Pu + Pu.T) / 2, F.T @ (objF.a + objF.P @ z0),
                               objF.r + float(objF.a @ z0) + float(z0 @ (objF.P @ z0)) / 2)
    else:
        o = ComposedObjective(objF, z0, F)
    return o, c


def kktDataReduced(H, A, g, r):
    """KKTData.reduced (KKTData.scala:68-91): returns (Hr, Ar, gr, r, nullIndices or None)."""
    n = H.shape[1]
    keep, null = [], []
    for j in range(n):
        t = np.linalg.norm(H[:, j]) + np.linalg.norm(A[:, j])
        if t > 0:
            keep.append(j)
        elif abs(g[j]) < 1e-15:
            null.append(j)
        else:
            raise UnsolvableSystemException("Unsolvable KKT system, row %d is zero with nonzero right hand side." % j)
    if not null:
        return H, A, g, r, None
    I = np.array(keep, dtype=int)
    return H[np.ix_(I, I)], A[:, I], g[I], r, null


def paddVector(x, nullIndices):
    """KKTData.paddVector (KKTData.scala:105-127)."""
    z = np.zeros(x.shape[0] + len(nullIndices))
    z[np.setdiff1d(np.arange(z.shape[0]), np.asarray(nullIndices, dtype=int))] = x
    return z


def symmetricLinearSystemSolve(H, r, tol):
    """SymmetricLinearSystem.scala:15-56 (double Ruiz equilibration, D6)."""
    d, Q = ruizEquilibrate(H)
    s = d * r
    if not checkSymmetric(Q, 1e-13):
        u = svdSolve(Q, s, tol)
    else:
        try:
            u = choleskySolve(Q, s, tol)
        except Exception:
            u = symSolve(Q, s, tol)
    return d * u


# ======================================================================================
# Model objects: closed-form objective / constraint families
# ======================================================================================


class Objective:
    """ObjectiveFunction.scala:12-14 (valueAt / gradientAt / hessianAt)."""
    dim: int

    def valueAt(self, x):
        raise NotImplementedError

    def gradientAt(self, x):
        raise NotImplementedError

    def hessianAt(self, x):
        raise NotImplementedError


class LinearObjective(Objective):
    """LinearObjectiveFunction.scala:5-22:  r + a'x ; Hessian zeros(n,n)."""
    kind = "linear"

    def __init__(self, a, r=0.0):
        self.a = np.asarray(a, dtype=np.float64)
        self.r = float(r)
        self.dim = self.a.shape[0]

    def valueAt(self, x):
        return self.r + float(self.a @ x)

    def gradientAt(self, x):
        return self.a.copy()

    def hessianAt(self, x):
        return np.zeros((self.dim, self.dim))


class QuadraticObjective(Objective):
    """QuadraticObjectiveFunction.scala:11-33:  r + a'x + x'Px/2."""
    kind = "quadratic"

    def __init__(self, P, a, r=0.0):
        self.P = np.asarray(P, dtype=np.float64)
        self.a = np.asarray(a, dtype=np.float64)
        self.r = float(r)
        self.dim = self.a.shape[0]

    def valueAt(self, x):
        return self.r + float(self.a @ x) + float(x @ (self.P @ x)) / 2

    def gradientAt(self, x):
        return self.a + self.P @ x

    def hessianAt(self, x):
        return self.P.copy()


class KLObjective(Objective):
    """Dist_KL.objectiveFunction, Dist_KL.scala:223-239:  sum x_j log(n x_j)."""
    kind = "kl"

    def __init__(self, n):
        self.dim = int(n)

    def valueAt(self, x):
        return float(x @ np.log(x * float(self.dim)))

    def gradientAt(self, x):
        return 1.0 + np.log(x) + math.log(self.dim)

    def hessianAt(self, x):
        return np.diag(1.0 / x)


class PNormObjective(Objective):
    """ObjectiveFunctions.p_norm_p, ObjectiveFunctions.scala:70-83:  sum |x_j|^p, p >= 2."""
    kind = "pnorm"

    def __init__(self, n, p):
        assert p >= 2
        self.dim = int(n)
        self.p = float(p)

    def valueAt(self, x):
        return float(np.sum(np.abs(x) ** self.p))

    def gradientAt(self, x):
        s = np.where(np.abs(x) < 1e-14, 0.0, np.sign(x))
        return s * self.p * (s * x) ** (self.p - 1)

    def hessianAt(self, x):
        return np.diag(self.p * (self.p - 1) * np.abs(x) ** (self.p - 2))


class DualKLObjective(Objective):
    """The convex dual objective -L_*(z) = w'z + R'exp(-B'z) of Dist_KL (Dist_KL.scala:143-163 through
    Duality.objF, Duality.scala:68-75).  B = [H; 1'; A] (Dist_KL.scala:131-137, 181-185), w = (u, 1, r) (:120-125),
    R = 1/(n e) with the reference's constant e = 2.7182811828459045 (:114-116, defect D8)."""
    kind = "kldual"
    E_REF = 2.7182811828459045

    def __init__(self, B, w, R):
        self.B = np.asarray(B, dtype=np.float64)
        self.w = np.asarray(w, dtype=np.float64)
        self.R = np.asarray(R, dtype=np.float64)
        self.dim = self.B.shape[0]

    def _y(self, z):
        return self.R * np.exp(-(self.B.T @ z))

    def valueAt(self, z):
        return float(self.w @ z) + float(np.sum(self._y(z)))

    def gradientAt(self, z):
        return self.w - self.B @ self._y(z)

    def hessianAt(self, z):
        y = self._y(z)
        return _sym_from_lower((self.B * y[None, :]) @ self.B.T)

    def primalOptimum(self, z):
        """Dist_KL.primalOptimum (:163)."""
        return self._y(z)


class ComposedObjective(Objective):
    """ObjectiveFunction.affineTransformed(z, F) (ObjectiveFunction.scala:26-40): h(u) = f(z + F u), gradient
    F' grad f, Hessian F' hess f F -- the generic transform every objective without a closed form of its own goes
    through (KL, p-norm).  (The reference gives the transformed function the dimension `dim - F.cols` instead of F.cols,
    ObjectiveFunction.scala:28; BarrierSolver's constructor assert objF.dim == startingPoint.length, BarrierSolver.scala:32,
    therefore only passes when n = 2p -- defect D11.  The restatement uses F.cols.)"""
    kind = "composed"

    def __init__(self, inner, z0, F):
        self.inner, self.z0, self.F = inner, np.asarray(z0, float), np.asarray(F, float)
        self.dim = self.F.shape[1]
        self.reference_dim = inner.dim - self.F.shape[1]      # what ObjectiveFunction.scala:28 computes

    def valueAt(self, u):
        return self.inner.valueAt(self.z0 + self.F @ u)

    def gradientAt(self, u):
        return self.F.T @ self.inner.gradientAt(self.z0 + self.F @ u)

    def hessianAt(self, u):
        H = self.F.T @ self.inner.hessianAt(self.z0 + self.F @ u) @ self.F
        return (H + H.T) / 2          # Breeze's cholesky insists on exact symmetry (see affineTransformedProblem)


@dataclass
class QuadCnt:
    """QuadraticConstraint.scala:7-40:  r + a'x + x'Px/2 <= ub."""
    P: np.ndarray
    a: np.ndarray
    r: float
    ub: float


class ConstraintSet:
    """ConstraintSet.scala with the closed-form families only: a block of linear constraints
    r_i + G_i x <= ub_i (LinearConstraint.scala:22-32) followed by optional quadratic constraints.
    `literal=True` evaluates per constraint as the reference does; otherwise with BLAS-2/3."""

    def __init__(self, G, r, ub, quad: Optional[List[QuadCnt]] = None, pointWhereDefined=None):
        self.G = np.ascontiguousarray(np.asarray(G, dtype=np.float64))
        self.m_lin, self.dim = self.G.shape
        self.r = np.zeros(self.m_lin) if r is None else np.asarray(r, dtype=np.float64)
        self.ub = np.asarray(ub, dtype=np.float64)
        self.quad = list(quad) if quad else []
        self.pointWhereDefined = None if pointWhereDefined is None else np.asarray(pointWhereDefined, float)
        self.feasiblePoint = None

    @property
    def numConstraints(self):
        return self.m_lin + len(self.quad)

    def ub_all(self):
        return np.concatenate([self.ub, np.array([q.ub for q in self.quad])]) if self.quad else self.ub

    # value g(x) (NOT g(x)-ub), Constraint.scala:16
    def valuesAt(self, x):
        v = self.r + self.G @ x
        if self.quad:
            vq = np.array([q.r + float(q.a @ x) + float(x @ (q.P @ x)) / 2 for q in self.quad])
            v = np.concatenate([v, vq])
        return v

    def constraintFunctionAt(self, x):
        """ConstraintSet.scala:90-94:  g(x) - ub."""
        return self.valuesAt(x) - self.ub_all()

    def gradientMatrixAt(self, x):
        """ConstraintSet.scala:100-110."""
        if not self.quad:
            return self.G
        return np.vstack([self.G] + [(q.a + q.P @ x)[None, :] for q in self.quad])

    def lambda0(self, x):
        """ConstraintSet.scala:116-120."""
        return -1.0 / self.constraintFunctionAt(x)

    def isSatisfiedStrictlyBy(self, x):
        """ConstraintSet.scala:28-29 with Constraint.isSatisfiedStrictly, Constraint.scala:23."""
        return bool(np.all(self.valuesAt(x) * (1 + 3e-16) < self.ub_all()))

    def addFeasiblePoint(self, x0):
        """ConstraintSet.scala:43-54."""
        assert x0.shape[0] == self.dim
        assert self.isSatisfiedStrictlyBy(x0)
        c = ConstraintSet(self.G, self.r, self.ub, self.quad, x0)
        c.feasiblePoint = np.array(x0, dtype=np.float64)
        return c

    def phase_I(self):
        """Constraint.phase_I (Constraint.scala:64-89) applied to every constraint, plus the
        feasible start of ConstraintSet.phase_I_Constraints_noEqs (ConstraintSet.scala:155-168)."""
        n = self.dim
        G1 = np.hstack([self.G, -np.ones((self.m_lin, 1))])
        quad1 = []
        for q in self.quad:
            P1 = np.zeros((n + 1, n + 1))
            P1[:n, :n] = q.P
            # gradient [a + P x ; -1]: linear part a1 = [a; -1]
            quad1.append(QuadCnt(P1, np.concatenate([q.a, [-1.0]]), q.r, q.ub))
        x0 = self.pointWhereDefined
        y0 = float(np.max(self.valuesAt(x0) - self.ub_all()))
        fp = np.concatenate([x0, [1 + y0]])
        c = ConstraintSet(G1, self.r, self.ub, quad1, fp)
        c.feasiblePoint = fp
        return c


@dataclass
class EqualityConstraint:
    """EqualityConstraint.scala:16-23 (the eager SolutionSpace QR is out of scope)."""
    A: np.ndarray
    b: np.ndarray

    def asInequalities(self, tol):
        """EqualityConstraint.scala:84-100: rows interleaved  a_i x <= b_i+tol ; -a_i x <= -b_i+tol."""
        p, n = self.A.shape
        G = np.empty((2 * p, n))
        ub = np.empty(2 * p)
        G[0::2] = self.A
        G[1::2] = -self.A
        ub[0::2] = self.b + tol
        ub[1::2] = -self.b + tol
        return G, ub


# ======================================================================================
# Barrier function family: BarrierSolver.scala:280-315
# ======================================================================================


def _sym_from_lower(M: np.ndarray) -> np.ndarray:
    """The reference's Hessian is a sum of exactly symmetric rank-1 terms (G*G.t, BarrierSolver.scala:313),
    so it passes Breeze's exact symmetry check; a dgemm result need not.  The vectorised mode mirrors the
    lower triangle so that it is exactly symmetric too."""
    Lw = np.tril(M)
    return Lw + np.tril(M, -1).T


class BarrierFunctions:
    def __init__(self, objF: Objective, cnts: ConstraintSet, literal: bool = False):
        self.objF, self.cnts, self.literal = objF, cnts, literal
        self.hessian_calls = 0

    def _slack(self, x, who):
        d = self.cnts.ub_all() - self.cnts.valuesAt(x)
        if np.any(d <= 0):
            raise NotStrictlyFeasible("%s: x not strictly feasible" % who)
        return d

    def value(self, t, x):
        """BarrierSolver.scala:280-289."""
        d = self._slack(x, "barrierFunction")
        if self.literal:
            s = t * self.objF.valueAt(x)
            for di in d:
                s = s - math.log(di)
            return s
        return t * self.objF.valueAt(x) - float(np.sum(np.log(d)))

    def gradient(self, t, x):
        """BarrierSolver.scala:291-301."""
        d = self._slack(x, "gradientBarrierFunction")
        c = self.cnts
        g = self.objF.gradientAt(x) * t
        if self.literal:
            D = c.gradientMatrixAt(x)
            for i in range(c.numConstraints):
                g = g + D[i] / d[i]
            return g
        g = g + c.G.T @ (1.0 / d[:c.m_lin])
        for k, q in enumerate(c.quad):
            g = g + (q.a + q.P @ x) / d[c.m_lin + k]
        return g

    def hessian(self, t, x):
        """BarrierSolver.scala:303-315:  t*hess f + sum_i [ grad_i grad_i' / d_i^2 + hess_i / d_i ]."""
        self.hessian_calls += 1
        d = self._slack(x, "hessianBarrierFunction")
        c = self.cnts
        H = self.objF.hessianAt(x) * t
        n = c.dim
        if self.literal:
            for i in range(c.m_lin):
                G = c.G[i]
                GGt = np.outer(G, G)
                H = H + GGt / (d[i] * d[i]) + np.zeros((n, n)) / d[i]
        else:
            w = 1.0 / (d[:c.m_lin] * d[:c.m_lin])
            H = H + _sym_from_lower(c.G.T @ (c.G * w[:, None]))
        for k, q in enumerate(c.quad):
            dk = d[c.m_lin + k]
            G = q.a + q.P @ x
            H = H + np.outer(G, G) / (dk * dk) + q.P / dk
        return H


# ======================================================================================
# Inner Newton solvers
# ======================================================================================


def equalityConstrainedSolve(bf: BarrierFunctions, t, x0, A, b, pars: SolverParams, stats=None) -> Solution:
    """EqualityConstrainedSolver.solve, EqualityConstrainedSolver.scala:37-107 (objF = barrier
    function at parameter t, C = strictly feasible set).  D7: shared `it` counter."""
    maxIter, alpha, beta = pars.maxIter, pars.alpha, pars.beta
    tol, tolEqSolve = pars.tolSolver, pars.tolEqSolve
    it_n = 0
    newtonDecrement = tol + 1
    x = np.array(x0, dtype=np.float64)
    y = bf.gradient(t, x)
    normGrad = float(np.linalg.norm(y))
    eqDiff = b - A @ x
    trials = []
    while it_n < maxIter and ((newtonDecrement > tol and normGrad > tol) or np.linalg.norm(eqDiff) > tol):
        f = bf.value(t, x)
        H = bf.hessian(t, x)
        info = KKTInfo()
        d = kkt_solve(H, A, y, eqDiff, tolEqSolve, info)[0]
        if stats is not None:
            stats.append(info)
        q = float(d @ y)
        newtonDecrement = -q / 2
        if newtonDecrement > tol:
            it = 0
            s = 1.0
            while (not bf.cnts.isSatisfiedStrictlyBy(x + d * s)) and it < 100:
                s *= beta
                it += 1
            if it == 100:
                raise NotConvergedException("Line search: backtracking into the set C failed.")
            while bf.value(t, x + d * s) > f + alpha * s * q and it < 100:
                s *= beta
                it += 1
            if it == 100:
                raise NotConvergedException("Line search: sufficient decrease not reached after 100 iterations.")
            trials.append(it)
            x = x + d * s
            y = bf.gradient(t, x)
            normGrad = float(np.linalg.norm(y))
            eqDiff = b - A @ x
        it_n += 1
    equalityGap = float(np.linalg.norm(eqDiff))
    return Solution(x, None, None, newtonDecrement, None, equalityGap, normGrad, None, it_n, it_n >= maxIter,
                    newton_steps=it_n, linesearch_trials=trials)


def unconstrainedSolve(bf: BarrierFunctions, t, x0, pars: SolverParams) -> Solution:
    """UnconstrainedSolver.solve, UnconstrainedSolver.scala:34-125, with D4 (rho = 1+1/4 == 1 in
    integer arithmetic, loop bounds 200 but failure test `it == 100`)."""
    maxIter, alpha, beta = pars.maxIter, pars.alpha, pars.beta
    tol, tolEqSolve = pars.tolSolver, pars.tolEqSolve
    it_n = 0
    newtonDecrement = tol + 1
    x = np.array(x0, dtype=np.float64)
    y = bf.gradient(t, x)
    normGrad = float(np.linalg.norm(y))
    trustRadius = float("nan")
    trials = []
    inC = bf.cnts.isSatisfiedStrictlyBy
    while it_n < maxIter and newtonDecrement > tol and normGrad > tol:
        f = bf.value(t, x)
        H = bf.hessian(t, x)
        try:
            d = choleskySolve(H, -y, tolEqSolve)
        except Exception:
            try:
                M = H + np.eye(H.shape[0]) * 1e-9
                d = choleskySolve(M, -y, tolEqSolve)
            except Exception:
                d = symSolve(H, -y, tolEqSolve)
        q = float(d @ y)
        newtonDecrement = -q / 2
        if newtonDecrement > tol:
            hNorm_d = math.sqrt(-q)
            if it_n == 0:
                trustRadius = hNorm_d
            s = d if (it_n == 0 or hNorm_d <= trustRadius) else d * (trustRadius / hNorm_d)
            it = 0
            tt = 1.0
            while (not inC(x + s * tt)) and it < 200:
                tt *= beta
                it += 1
            if it == 100:
                raise NotConvergedException("Line search: backtracking into the set C failed.")
            rho = 1  # D4: `1+1/4` in Scala integer arithmetic
            if not inC(x + s):
                trustRadius /= rho
            else:
                f_new = bf.value(t, x + s * tt)
                if f_new > f + alpha * tt * q:
                    trustRadius /= rho
                if f_new < f + ((1 + alpha) / 2) * tt * q and trustRadius <= hNorm_d:
                    trustRadius *= rho
            while bf.value(t, x + s * tt) > f + alpha * tt * q and it < 200:
                tt *= beta
                it += 1
            if it == 100:
                raise NotConvergedException("Line search: sufficient decrease not reached")
            trials.append(it)
            x = x + s * tt
            y = bf.gradient(t, x)
            normGrad = float(np.linalg.norm(y))
        it_n += 1
    return Solution(x, None, None, newtonDecrement, None, None, normGrad, None, it_n, it_n >= maxIter,
                    newton_steps=it_n, linesearch_trials=trials)


# ======================================================================================
# CvxUtils.scala:61-87 termination criteria
# ======================================================================================


def standardTerminationCriterion(pars: SolverParams) -> Callable[[OptimizationState], bool]:
    def crit(os: OptimizationState) -> bool:
        return (os.dualityGap < pars.tolSolver) and (os.equalityGap is None or os.equalityGap < pars.tolSolver)
    return crit


def phase_I_TerminationCriterion(os: OptimizationState) -> bool:
    return (os.objectiveFunctionValue < 0) and (os.equalityGap is None or os.equalityGap < 1e-6)


# ======================================================================================
# BarrierSolver.scala:70-188
# ======================================================================================


def barrierSolve(objF: Objective, cnts: ConstraintSet, eqs: Optional[EqualityConstraint], pars: SolverParams,
                 terminationCriterion=None, literal=False, x0=None, kkt_stats=None) -> Solution:
    """BarrierSolver.solveWithEQs (:124-177) / solveWithoutEQs (:70-117); starting point
    cnts.feasiblePoint (BarrierSolver.apply :269-278)."""
    if terminationCriterion is None:
        terminationCriterion = standardTerminationCriterion(pars)
    bf = BarrierFunctions(objF, cnts, literal)
    x = np.array(cnts.feasiblePoint if x0 is None else x0, dtype=np.float64)
    assert cnts.isSatisfiedStrictlyBy(x), "Starting point x not in set C"
    mu = 10.0
    t = 1.0
    dualityGap = DOUBLE_MAX
    state = OptimizationState(None, None, dualityGap, DOUBLE_MAX if eqs is not None else None, DOUBLE_MAX)
    sol = None
    maxIter = 1000 / mu
    it = 0
    m = cnts.numConstraints
    stage_steps = []
    trials = []
    equalityGap = None
    while (not terminationCriterion(state)) and it < maxIter:
        if eqs is not None:
            sol = equalityConstrainedSolve(bf, t, x, eqs.A, eqs.b, pars, kkt_stats)
        else:
            sol = unconstrainedSolve(bf, t, x, pars)
        x = sol.x
        stage_steps.append(sol.newton_steps)
        trials.extend(sol.linesearch_trials)
        objValue = objF.valueAt(x)
        dualityGap = m / t
        if eqs is not None:
            equalityGap = sol.equalityGap
            state = OptimizationState(None, None, dualityGap, equalityGap, objValue)
        else:
            state = OptimizationState(None, None, dualityGap, 0.0, objValue)
        t = mu * t
        it += 1
    return Solution(sol.x, None, None, sol.newtonDecrement, dualityGap, equalityGap, sol.normGrad,
                    sol.normDualResidual, sol.iter, sol.maxedOut,
                    newton_steps=int(sum(stage_steps)), stage_newton_steps=stage_steps, outer_stages=it,
                    linesearch_trials=trials)


# ======================================================================================
# Phase I: ConstraintSet.scala:131-168,310-395,556-575
# ======================================================================================


def phase_I_Analysis(cnts: ConstraintSet, eqs: Optional[EqualityConstraint], pars: SolverParams, literal=False):
    """ConstraintSet.phase_I_Analysis (:404-414).  With equalities the rows of
    eqs.asInequalities(1e-6) are appended (:326-347); then the no-equality basic phase I (:355-395).
    Returns (x_feas, s_feas, barrier Solution)."""
    if eqs is not None:
        Ge, ube = eqs.asInequalities(1e-6)
        assert not cnts.quad or True
        # theConstraints = constraints ::: ineqs2  -> linear block then eq rows, quadratics keep order
        # only when there are none; mixed sets put the equality rows after the quadratics in the
        # reference.  The closed-form device path carries [linear ; eq-as-ineq] + quadratics.
        G = np.vstack([cnts.G, Ge])
        r = np.concatenate([cnts.r, np.zeros(Ge.shape[0])])
        ub = np.concatenate([cnts.ub, ube])
        work = ConstraintSet(G, r, ub, cnts.quad, cnts.pointWhereDefined)
    else:
        work = cnts
    n = work.dim
    feasCnts = work.phase_I()
    e = np.zeros(n + 1)
    e[n] = 1.0
    feasObjF = LinearObjective(e, 0.0)      # phase_I_ObjectiveFunction, ConstraintSet.scala:131-144
    sol = barrierSolve(feasObjF, feasCnts, None, pars, phase_I_TerminationCriterion, literal)
    w = sol.x
    return w[:n], float(w[n]), sol


@dataclass
class FeasibilityReport:
    """FeasibilityReport.scala:12-48."""
    x0: np.ndarray
    s: np.ndarray
    isStrictlyFeasible: bool
    constraintSet: "ConstraintSet"
    equalityConstraintError: Optional[float]

    def isFeasible(self, tol):
        e = 0.0 if self.equalityConstraintError is None else self.equalityConstraintError
        return bool(np.max(self.s) < tol and e < tol)

    def violatedConstraints(self, tol):
        """indices i with g_i(x0) > ub_i + tol (Constraint.isSatisfiedWithTolerance)."""
        c = self.constraintSet
        return [int(i) for i in np.nonzero(~(c.valuesAt(self.x0) <= c.ub_all() + tol))[0]]


def phase_I_SOI_problem(cnts: ConstraintSet, eqs: Optional[EqualityConstraint]):
    """Sum-of-infeasibilities phase I, [boyd] 11.4.1 p580: variable u = (x, s_1..s_p), one s_j per constraint,
    g_j(x) - s_j <= ub_j and -s_j <= 0, objective sum_j s_j.
    ConstraintSet.phase_I_SOI_ObjectiveFunction / phase_I_SOI_Constraints (ConstraintSet.scala:233-282),
    Constraint.phase_I_SOI / phase_I_SOI_Constraints (Constraint.scala:101-159),
    EqualityConstraint.phase_I_SOI_EqualityConstraint (EqualityConstraint.scala:50-55).
    Row order here: [linear g_j - s_j ; -s_j <= 0 ; quadratic g_j - s_j] (the reference lists the quadratic rows
    before the positivity rows; only the summation order differs)."""
    n, p, ml = cnts.dim, cnts.numConstraints, cnts.m_lin
    N = n + p
    G = np.zeros((ml + p, N))
    G[:ml, :n] = cnts.G
    G[np.arange(ml), n + np.arange(ml)] = -1.0
    G[ml + np.arange(p), n + np.arange(p)] = -1.0
    r = np.concatenate([cnts.r, np.zeros(p)])
    ub = np.concatenate([cnts.ub, np.zeros(p)])
    quad = []
    for k, q in enumerate(cnts.quad):
        P1 = np.zeros((N, N))
        P1[:n, :n] = q.P
        a1 = np.zeros(N)
        a1[:n] = q.a
        a1[n + ml + k] = -1.0
        quad.append(QuadCnt(P1, a1, q.r, q.ub))
    x = cnts.pointWhereDefined
    viol = cnts.valuesAt(x) - cnts.ub_all()
    fp = np.concatenate([x, np.maximum(0.5, 1.0 + viol)])           # ConstraintSet.scala:270-272
    soi = ConstraintSet(G, r, ub, quad, fp)
    soi.feasiblePoint = fp
    a = np.zeros(N)
    a[n:] = 1.0
    objF = LinearObjective(a, 0.0)
    eqs_soi = None if eqs is None else EqualityConstraint(np.hstack([eqs.A, np.zeros((eqs.A.shape[0], p))]), eqs.b)
    return objF, soi, eqs_soi


def phase_I_Analysis_SOI(cnts: ConstraintSet, eqs: Optional[EqualityConstraint], pars: SolverParams, literal=False):
    """ConstraintSet.phase_I_Analysis_SOI (ConstraintSet.scala:511-545): a full barrier solve (standard termination)
    of the SOI problem.  isStrictlySatisfied reproduces the reference's test `(0 until n).forall(s_feas(j) < 0)`
    (index range n instead of p -- defect D9; as every s_j > 0 inside the barrier's domain the first index already
    fails, so no out-of-range access can happen and the flag is always false)."""
    n, p = cnts.dim, cnts.numConstraints
    objF, soi, eqs_soi = phase_I_SOI_problem(cnts, eqs)
    sol = barrierSolve(objF, soi, eqs_soi, pars, None, literal)
    w = sol.x
    x_feas, s_feas = w[:n], w[n:n + p]
    eqError = None if eqs is None else float(np.linalg.norm(eqs.A @ x_feas - eqs.b))
    strict = True
    for j in range(n):
        if not (s_feas[j] < 0):          # j >= p would raise IndexError exactly where the JVM throws
            strict = False
            break
    strict = strict and ((0.0 if eqError is None else eqError) < pars.tolSolver)
    return FeasibilityReport(x_feas, s_feas, strict, cnts, eqError), sol


def withFeasiblePoint(cnts: ConstraintSet, eqs: Optional[EqualityConstraint], pars: SolverParams, literal=False):
    """ConstraintSet.withFeasiblePoint (:556-575) + FeasibilityReport.isFeasible
    (FeasibilityReport.scala:36-37)."""
    if cnts.feasiblePoint is not None:
        return cnts, None
    tol = pars.tolSolver
    x0, s, sol = phase_I_Analysis(cnts, eqs, pars, literal)
    if not (s < tol):
        raise InfeasibleProblemException("Problem not feasible within tolerance %g (s=%g)" % (tol, s))
    return cnts.addFeasiblePoint(x0), sol


# ======================================================================================
# PrimalDualSolver.scala
# ======================================================================================


class PrimalDual:
    def __init__(self, objF: Objective, cnts: ConstraintSet, eqs: Optional[EqualityConstraint],
                 pars: SolverParams, literal=False, bug_compat=False):
        self.objF, self.cnts, self.eqs, self.pars = objF, cnts, eqs, pars
        self.literal, self.bug_compat = literal, bug_compat
        self.dim = cnts.dim
        self.numIneqs = cnts.numConstraints

    # residuals :63-144
    def dualResidual(self, x, lam, nu=None):
        r = self.objF.gradientAt(x) + self.cnts.gradientMatrixAt(x).T @ lam
        if nu is not None:
            r = r + self.eqs.A.T @ nu
        return r

    def centralResidual(self, t, x, lam):
        g = self.cnts.constraintFunctionAt(x)
        return -lam * g - 1.0 / t

    def primalResidual(self, x):
        return self.eqs.A @ x - self.eqs.b

    def residual(self, t, x, lam, nu=None):
        if nu is None:
            return np.concatenate([self.dualResidual(x, lam), self.centralResidual(t, x, lam)])
        return np.concatenate([self.dualResidual(x, lam, nu), self.centralResidual(t, x, lam),
                               self.primalResidual(x)])

    def rhs1(self, t, x):
        """:162-176   -grad f + sum_i grad g_i / (t f_i)."""
        c = self.cnts
        fx = c.constraintFunctionAt(x)
        res = -self.objF.gradientAt(x)
        if self.literal:
            D = c.gradientMatrixAt(x)
            for i in range(self.numIneqs):
                res = res + D[i] / (t * fx[i])
            return res
        return res + c.gradientMatrixAt(x).T @ (1.0 / (t * fx))

    def deltaLambda(self, t, x, dx, lam):
        """:184-209."""
        r_cent = self.centralResidual(t, x, lam)
        gx = self.cnts.constraintFunctionAt(x)
        assert np.all(gx < 0), "gx not < 0, line search did not pull back into strictly feasible region!"
        w = self.cnts.gradientMatrixAt(x) @ dx
        return (-lam * w + r_cent) / gx

    def kktMatrix_noEqs(self, x, lam):
        """:216-240   hess f + sum_i [ lam_i hess g_i - (lam_i/f_i) grad g_i grad g_i' ]."""
        c = self.cnts
        fx = c.constraintFunctionAt(x)
        assert np.all(fx < 0), "fi not < 0"
        H = self.objF.hessianAt(x)
        if self.literal:
            n = self.dim
            for i in range(c.m_lin):
                g = c.G[i]
                H = H + (np.zeros((n, n)) * lam[i] - np.outer(g, g) * (lam[i] / fx[i]))
        else:
            w = -(lam[:c.m_lin] / fx[:c.m_lin])
            H = H + _sym_from_lower(c.G.T @ (c.G * w[:, None]))
        for k, q in enumerate(c.quad):
            i = c.m_lin + k
            g = q.a + q.P @ x
            H = H + (q.P * lam[i] - np.outer(g, g) * (lam[i] / fx[i]))
        return H

    def surrogateDualityGap(self, x, lam):
        """:289-297."""
        return float(-(self.cnts.constraintFunctionAt(x) @ lam))

    def newton_direction(self, t, x, lam, nu=None, info=None):
        """One search direction (dx, dlam, dnu) given the iterate; :254-285, :416-421, :589-593.
        With equalities and bug_compat the reference's sign defect D2 is reproduced
        (q = v + A'nu handed to KKTSystem, which solves H dx + A'w = -q)."""
        H = self.kktMatrix_noEqs(x, lam)
        v = self.rhs1(t, x)
        if nu is None:
            dx = symmetricLinearSystemSolve(H, v, self.pars.tolEqSolve)
            return dx, self.deltaLambda(t, x, dx, lam), None
        A = self.eqs.A
        r_pri = self.primalResidual(x)
        if self.bug_compat:
            q = v + A.T @ nu                      # D2 (":280-284  FIX ME: check this!")
        else:
            q = -v + A.T @ nu                     # B&V (11.55): H dx + A' dnu = v - A' nu
        dx, dnu = kkt_solve(H, A, q, -r_pri, self.pars.tolEqSolve, info)
        return dx, self.deltaLambda(t, x, dx, lam), dnu

    def lineSearch(self, t, x, lam, nu, dx, dlam, dnu):
        """lineSearch_noEQs :311-374 / lineSearch_withEQs :478-543.  Returns (x_s, lam_s, nu_s, trials)."""
        assert np.all(lam > 0.0), "lambda not positive"
        pars = self.pars
        s0 = 1.0
        neg = dlam < 0
        if np.any(neg):
            s0 = min(s0, float(np.min(-lam[neg] / dlam[neg])))
        s = 0.99 * s0
        alpha, beta = pars.alpha, pars.beta
        norm_rt = float(np.linalg.norm(self.residual(t, x, lam, nu)))

        def trial(s):
            x_s = x + dx * s
            lam_s = lam + dlam * s
            assert np.all(lam_s > 0)
            nu_s = None if nu is None else nu + dnu * s
            feas = self.cnts.isSatisfiedStrictlyBy(x_s)
            if self.objF.kind == "kl" and not np.all(x_s > 0):
                nrm = float("nan")
            else:
                with np.errstate(all="ignore"):
                    nrm = float(np.linalg.norm(self.residual(t, x_s, lam_s, nu_s)))
            ok = feas and (nrm < (1 - alpha * s) * norm_rt)
            return ok, x_s, lam_s, nu_s

        ok, x_s, lam_s, nu_s = trial(s)
        it = 0
        maxIter = -30 / math.log(beta)
        while (not ok) and it <= maxIter:
            s *= beta
            ok, x_s, lam_s, nu_s = trial(s)
            it += 1
        if it >= maxIter:
            raise LineSearchFailedException("Line search unsuccessful.")
        return x_s, lam_s, nu_s, it

    def solve(self, terminationCriterion=None, x0=None, max_steps=None) -> Solution:
        """solve_noEQs :381-460 / solve_withEQs :550-621; solve :628-634.
        bug_compat reproduces D1 (with equalities the line search always restarts from the initial
        iterate because `u` is never reassigned, :569,594-598)."""
        pars = self.pars
        if terminationCriterion is None:
            def terminationCriterion(os):
                return os.dualityGap < pars.tolSolver and os.normDualResidual < pars.tolSolver
        mu = 10.0
        x = np.array(self.cnts.feasiblePoint if x0 is None else x0, dtype=np.float64)
        lam = self.cnts.lambda0(x)
        withEqs = self.eqs is not None
        nu = np.zeros(self.eqs.A.shape[0]) if withEqs else None
        u0 = (x.copy(), lam.copy(), None if nu is None else nu.copy())
        dualityGap = self.surrogateDualityGap(x, lam)
        equalityGap = DOUBLE_MAX if withEqs else 0.0
        normDualResidual = DOUBLE_MAX
        t = _ieee_div(mu * self.numIneqs, dualityGap)
        state = OptimizationState(None, None, dualityGap, equalityGap, DOUBLE_MAX, normDualResidual)
        maxIter = (1500 if withEqs else 2000) / mu
        if max_steps is not None:
            maxIter = min(maxIter, max_steps)
        it = 0
        trials = []
        while (not terminationCriterion(state)) and it < maxIter:
            dx, dlam, dnu = self.newton_direction(t, x, lam, nu)
            if withEqs and self.bug_compat:
                x, lam, nu, k = self.lineSearch(t, u0[0], u0[1], u0[2], dx, dlam, dnu)     # D1
            else:
                x, lam, nu, k = self.lineSearch(t, x, lam, nu, dx, dlam, dnu)
            trials.append(k)
            objValue = self.objF.valueAt(x)
            dualityGap = self.surrogateDualityGap(x, lam)
            if withEqs:
                equalityGap = float(np.linalg.norm(self.eqs.A @ x - self.eqs.b))
                normDualResidual = float(np.linalg.norm(self.residual(t, x, lam, nu)))
            else:
                normDualResidual = float(np.linalg.norm(self.dualResidual(x, lam)))
            state = OptimizationState(None, None, dualityGap, equalityGap, objValue, normDualResidual)
            t = _ieee_div(mu * self.numIneqs, dualityGap)
            it += 1
        return Solution(x, lam, nu, None, dualityGap, equalityGap if withEqs else None, None, normDualResidual,
                        it - 1, it == maxIter, newton_steps=it, linesearch_trials=trials)


# ======================================================================================
# Problem-level helpers: OptimizationProblem.scala:133-196, Dist_KL.scala:270-315
# ======================================================================================


def solveProblem(objF, cnts, eqs, solverType="BR", pars=None, literal=False, bug_compat=False):
    """OptimizationProblem.withoutFeasiblePoint(...).solve: phase I when the constraint set has no
    feasible point, then the BR / PD solver."""
    pars = pars or SolverParams.standardParams()
    assert solverType in ("BR", "PD")
    c, phase1 = withFeasiblePoint(cnts, eqs, pars, literal)
    if solverType == "BR":
        sol = barrierSolve(objF, c, eqs, pars, None, literal)
    else:
        sol = PrimalDual(objF, c, eqs, pars, literal, bug_compat).solve()
    return sol, phase1


def dist_KL_dual_problem(n, H=None, u=None, A=None, r=None):
    """Duality.dualProblem for Dist_KL (Duality.scala:77-112): min -L_*(z) s.t. lambda = z[:numInequalities] >= 0
    (Constraints.firstCoordinatesPositive), no equalities, feasible start z = 0.001."""
    ones = np.ones((1, n))
    Aext = ones if A is None else np.vstack([ones, A])                      # A_with_probEQ: the sum-to-one row FIRST
    rext = np.array([1.0]) if r is None else np.concatenate([[1.0], r])
    B = Aext if H is None else np.vstack([H, Aext])
    w = rext if H is None else np.concatenate([u, rext])
    R = np.full(n, 1.0 / (n * DualKLObjective.E_REF))
    mI = 0 if H is None else H.shape[0]
    D = B.shape[0]
    G = np.zeros((mI, D))
    G[np.arange(mI), np.arange(mI)] = -1.0
    cnts = ConstraintSet(G, np.zeros(mI), np.zeros(mI), None, np.zeros(D))
    cnts = cnts.addFeasiblePoint(np.full(D, 0.001)) if mI > 0 else cnts
    if mI == 0:
        cnts.feasiblePoint = np.full(D, 0.001)
    return DualKLObjective(B, w, R), cnts, mI


def solveDual(n, H=None, u=None, A=None, r=None, pars=None):
    """Duality.solveDual (Duality.scala:119-133) with the barrier solver: returns the primal Solution."""
    pars = pars or SolverParams.standardParams()
    objF, cnts, mI = dist_KL_dual_problem(n, H, u, A, r)
    sol = barrierSolve(objF, cnts, None, pars)
    z = sol.x
    sol.lam = z[:mI]
    sol.nu = z[mI:] if z.shape[0] > mI else None
    sol.z = z
    sol.x = objF.primalOptimum(z)
    return sol


def dist_KL_problem(n, H=None, u=None, A=None, r=None):
    """Dist_KL.apply (Dist_KL.scala:270-315): inequality rows Hx<=u then positivity rows -x_j<=0;
    equalities [A; 1'] x = [r; 1] (addEqualities stacks the sum-to-one row LAST, :249-251 and
    EqualityConstraint.scala:31-37); pointWhereDefined = 1/n."""
    Gpos = -np.eye(n)
    if H is not None:
        G = np.vstack([H, Gpos])
        ub = np.concatenate([u, np.zeros(n)])
    else:
        G, ub = Gpos, np.zeros(n)
    cnts = ConstraintSet(G, np.zeros(G.shape[0]), ub, None, np.full(n, 1.0 / n))
    ones = np.ones((1, n))
    if A is not None:
        eqs = EqualityConstraint(np.vstack([A, ones]), np.concatenate([r, [1.0]]))
    else:
        eqs = EqualityConstraint(ones, np.array([1.0]))
    return KLObjective(n), cnts, eqs

"""Host-side mirror of the reference's problem / solver API (seam A), the closed-form families only:

  SolverParams                       SolverParams.scala:24-46
  Solution                           Solution.scala:32-43
  LinearObjectiveFunction(dim,r,a)   LinearObjectiveFunction.scala:5-22
  QuadraticObjectiveFunction(dim,r,a,P)   QuadraticObjectiveFunction.scala:11-33
  Dist_KL objective                  Dist_KL.scala:223-239
  ConstraintSet(H, u)                ConstraintSet.apply(H,u,C), ConstraintSet.scala:621-638 (rows are LinearConstraints)
  EqualityConstraint(A, b)           EqualityConstraint.scala:16-23
  BarrierSolver(objF, cnts, eqs, pars).solve()          BarrierSolver.scala:184-188, 269-278
  PrimalDualSolver(objF, cnts, eqs, pars).solve()       PrimalDualSolver.scala:628-641, 718-728
  OptimizationProblem(id, objF, ineqs, eqs, solverType, pars).solve()   OptimizationProblem.scala:19,133-196
  Dist_KL(n, H, u, A, r, solverType, pars)              Dist_KL.scala:270-315

The objects only describe the problem; `solve` uploads the descriptor once (cvxb_problem_create) and
the whole solve -- phase I included -- runs device-resident in libcvxb (cvxb_barrier_solve /
cvxb_pd_solve).  numpy arrays stand in for Breeze DenseMatrix / DenseVector.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import (KktInfo, Params, ProblemDesc, SolutionC, check, dptr, fmat, fvec, ptr,
                   OBJ_KL, OBJ_KLDUAL, OBJ_LINEAR, OBJ_PNORM, OBJ_QUADRATIC)


@dataclass
class SolverParams:
    maxIter: int = 1000
    alpha: float = 0.04
    beta: float = 0.8
    tolSolver: float = 1e-8
    tolEqSolve: float = 1e-1
    tolFeas: float = 1e-7
    delta: float = 1e-6
    # not in the reference record: PrimalDualSolver.solve_withEQs defects D1/D2 reproduced when True
    bugCompat: int = 0         # bit 0: PD defects D1/D2; bit 1: the reference's literal block elimination (cvxb.h)
    # benchmark aid: stop after this many Newton steps in total (0 = run to termination)
    stepLimit: int = 0

    @staticmethod
    def standardParams(dim: int = 0) -> "SolverParams":
        return SolverParams()

    def to_c(self, handle) -> Params:
        p = handle.default_params()
        for k in ("maxIter", "alpha", "beta", "tolSolver", "tolEqSolve", "tolFeas", "delta"):
            setattr(p, k, getattr(self, k))
        p.bugCompat = int(self.bugCompat)
        p.stepLimit = int(self.stepLimit)
        return p


@dataclass
class Solution:
    """Solution.scala:32-43 (Option fields are None when the reference has None) + bookkeeping."""
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
    objective: float = float("nan")
    outer_stages: int = 0
    newton_steps: int = 0
    executed_newton_steps: int = 0
    stage_newton_steps: List[int] = field(default_factory=list)
    phase1_newton_steps: int = 0
    phase1_executed_steps: int = 0
    phase1_stages: int = 0
    phase1_s: float = float("nan")
    linesearch_trials: int = 0
    kkt_fallbacks: int = 0
    kkt_regularized: int = 0
    solve_ms: float = 0.0


class ObjectiveFunction:
    dim: int
    kind: int


class LinearObjectiveFunction(ObjectiveFunction):
    kind = OBJ_LINEAR

    def __init__(self, dim, r, a):
        self.dim, self.r, self.a = int(dim), float(r), fvec(a)
        assert self.a.shape[0] == self.dim


class QuadraticObjectiveFunction(ObjectiveFunction):
    kind = OBJ_QUADRATIC

    def __init__(self, dim, r, a, P):
        self.dim, self.r, self.a, self.P = int(dim), float(r), fvec(a), fmat(P)
        assert self.a.shape[0] == self.dim and self.P.shape == (self.dim, self.dim)


class KLObjectiveFunction(ObjectiveFunction):
    """Dist_KL.objectiveFunction: sum_j x_j log(n x_j)."""
    kind = OBJ_KL

    def __init__(self, dim):
        self.dim, self.r, self.a = int(dim), 0.0, None


class PNormObjectiveFunction(ObjectiveFunction):
    """ObjectiveFunctions.p_norm_p(dim, p): sum |x_j|^p, p >= 2 (ObjectiveFunctions.scala:70-83)."""
    kind = OBJ_PNORM

    def __init__(self, dim, p):
        assert p >= 2, "p-norm needs p>=2 but p=%s" % p
        self.dim, self.p, self.r, self.a = int(dim), float(p), 0.0, None


class DualKLObjectiveFunction(ObjectiveFunction):
    """-L_*(z) = w'z + R'exp(-B'z): the objective of Duality.dualProblem for Dist_KL (Dist_KL.scala:143-163)."""
    kind = OBJ_KLDUAL

    def __init__(self, B, w, R):
        self.B, self.a, self.R, self.r = fmat(B), fvec(w), fvec(R), 0.0
        self.dim = self.B.shape[0]
        assert self.a.shape[0] == self.dim and self.R.shape[0] == self.B.shape[1]


class QuadraticConstraint:
    """QuadraticConstraint(id, dim, ub, r, a, P):  r + a'x + x'Px/2 <= ub, P symmetric (QuadraticConstraint.scala:7-40)."""

    def __init__(self, id, dim, ub, r, a, P):
        self.id, self.dim, self.ub, self.r, self.a, self.P = id, int(dim), float(ub), float(r), fvec(a), fmat(P)
        if self.a.shape[0] != self.dim:
            raise ValueError("Vector a must be of dimension %d but length(a) %d" % (self.dim, self.a.shape[0]))
        if self.P.shape != (self.dim, self.dim):
            raise ValueError("Matrix P must be square of dimension %d" % self.dim)
        assert np.linalg.norm(self.P - self.P.T) < 1e-13, "P not symmetric"       # checkSymmetric(P, 1e-13)


class ConstraintSet:
    """Linear inequality block  r + Hx <= u  (one LinearConstraint per row), followed by optional
    QuadraticConstraints (the closed-form families of ConstraintSet.scala)."""

    def __init__(self, H, u, pointWhereDefined=None, r=None, quadratic=None):
        self.H = fmat(H)
        self.u = fvec(u)
        self.r = None if r is None else fvec(r)
        self.dim = self.H.shape[1]
        assert self.u.shape[0] == self.H.shape[0]
        self.quadratic = list(quadratic) if quadratic else []
        assert all(q.dim == self.dim for q in self.quadratic)
        self.pointWhereDefined = None if pointWhereDefined is None else fvec(pointWhereDefined)
        self.feasiblePoint = None

    @property
    def numConstraints(self):
        return self.H.shape[0] + len(self.quadratic)

    def addFeasiblePoint(self, x0):
        """ConstraintSet.addFeasiblePoint (ConstraintSet.scala:43-54)."""
        c = ConstraintSet(self.H, self.u, self.pointWhereDefined if self.pointWhereDefined is not None else x0, self.r,
                          self.quadratic)
        c.feasiblePoint = fvec(x0)
        return c

    # ---- evaluation on the device (the set is uploaded once, with a zero objective)
    def _device(self, handle=None):
        dev = getattr(self, "_dev", None)
        if dev is None or (handle is not None and dev.handle is not handle):
            dev = self._dev = _DeviceProblem(LinearObjectiveFunction(self.dim, 0.0, np.zeros(self.dim)), self, None, handle)
        return dev

    def valuesAt(self, x, handle=None):
        """g_i(x) for every constraint (linear rows first, then the quadratic ones)."""
        return self._values(x, handle)[0]

    def isSatisfiedStrictlyBy(self, x, handle=None) -> bool:
        """ConstraintSet.scala:28-29."""
        return self._values(x, handle)[1]

    def _values(self, x, handle=None):
        dev = self._device(handle)
        g = np.empty(max(self.numConstraints, 1))
        ok = C.c_int(0)
        check(dev.handle.lib.cvxb_constraint_values(dev.handle._h, dev._p, ptr(fvec(x)), ptr(g), C.byref(ok)))
        return g[:self.numConstraints], bool(ok.value)

    def ub_all(self):
        return np.concatenate([self.u, np.array([q.ub for q in self.quadratic])]) if self.quadratic else self.u

    # ---- phase I, basic method (ConstraintSet.scala:355-414)
    def phase_I_Analysis(self, eqs, pars=None, debugLevel=0, handle=None) -> "FeasibilityReport":
        """Basic phase I ([boyd] 11.4.1 p579): one extra variable s, minimise s until s < 0; with equalities the rows
        of eqs.asInequalities(1e-6) are appended (ConstraintSet.scala:310-347).  Device-resident (cvxb_phase1)."""
        pars = pars if pars is not None else SolverParams.standardParams()
        solver = BarrierSolver(LinearObjectiveFunction(self.dim, 0.0, np.zeros(self.dim)), self, eqs, pars, None, handle)
        x_feas, ph = solver.phase_I()
        s_feas = float(ph.x[self.dim])
        eqError = 0.0 if eqs is None else eqs.errorAt(x_feas)
        strict = s_feas < 0.0 and (eqs is None or eqError < pars.tolSolver)
        rep = FeasibilityReport(x_feas, np.array([s_feas]), strict, self, eqError)
        rep.solution = ph
        return rep

    # ---- phase I with the equalities eliminated (ConstraintSet.scala:424-477)
    def phase_I_Constraints_noEqs(self) -> "ConstraintSet":
        """g_j(x) - s <= ub_j in dimension n + 1 with the feasible point (x0, 1 + max_j(g_j(x0) - ub_j)),
        x0 = pointWhereDefined (ConstraintSet.scala:153-168; Constraint.phase_I, Constraint.scala:64-89)."""
        n, ml = self.dim, self.H.shape[0]
        G1 = np.empty((ml, n + 1), order="F")
        G1[:, :n] = self.H
        G1[:, n] = -1.0
        quad = []
        for q in self.quadratic:
            P1 = np.zeros((n + 1, n + 1))
            P1[:n, :n] = q.P
            quad.append(QuadraticConstraint(str(q.id) + "_phase_I", n + 1, q.ub, q.r, np.concatenate([q.a, [-1.0]]), P1))
        x0 = self.pointWhereDefined
        assert x0 is not None, "phase I needs pointWhereDefined"
        s0 = 1.0 + float(np.max(self.valuesAt(x0) - self.ub_all()))
        fp = np.concatenate([x0, [s0]])
        return ConstraintSet(G1, self.u, fp, self.r, quad).addFeasiblePoint(fp)

    def phase_I_Analysis_by_reduction(self, eqs, pars=None, debugLevel=0, handle=None, corrected=False) -> "FeasibilityReport":
        """ConstraintSet.phase_I_Analysis_by_reduction (ConstraintSet.scala:424-477): phase I with Ax = b parameterised
        as x = z0 + F u.  AS WRITTEN THE REFERENCE CANNOT COMPLETE THIS CALL: it reduces the (n+1)-dimensional phase-I
        solver with the n-dimensional solution space of Ax = b (`solverNoEqs.reduced(solEqs)`, :447-449), and
        BarrierSolver.reduced -> SolutionSpace.parameter subtracts z0 (length n) from the starting point (length
        n + 1) -- a Breeze dimension error.  Mirrored: the device call refuses with the dimension error
        (CVXB_EDIM -> DimensionMismatch, an AssertionError).  corrected=True does what the method's doc comment
        describes: the map (x, s) = (z0, 0) + blockdiag(F, 1)(u, s), the barrier solve of min s in (u, s), and the
        report in x = z0 + F u with the reference's strictness test (s < 0 and ||Ax - b|| < tolSolver, :461-462)."""
        from .linalg import SolutionSpace
        pars = pars if pars is not None else SolverParams.standardParams()
        n = self.dim
        a = np.zeros(n + 1)
        a[n] = 1.0
        solverNoEqs = BarrierSolver(LinearObjectiveFunction(n + 1, 0.0, a), self.phase_I_Constraints_noEqs(), None, pars, None, handle)
        solEqs = eqs.solutionSpace if handle is None or eqs.solutionSpace.handle is solverNoEqs.handle else SolutionSpace(eqs.A, eqs.b, solverNoEqs.handle)
        if not corrected:
            solver = solverNoEqs.reduced(solEqs)          # raises DimensionMismatch: F.rows = n, dim(problem) = n + 1
            raise AssertionError("unreachable: the reference's call cannot succeed")      # pragma: no cover
        F, z0 = solEqs.F, solEqs.z0
        k = F.shape[1]
        F1 = np.zeros((n + 1, k + 1), order="F")
        F1[:n, :k] = F
        F1[n, k] = 1.0
        ext = SolutionSpace.from_basis(np.concatenate([z0, [0.0]]), F1, solverNoEqs.handle)
        solver = solverNoEqs.reduced(ext)
        sol = solver.solve(debugLevel)
        u_feas, s_feas = sol.x[:k], float(sol.x[k])
        x_feas = z0 + F @ u_feas
        eqError = eqs.errorAt(x_feas)
        rep = FeasibilityReport(x_feas, np.array([s_feas]), s_feas < 0.0 and eqError < pars.tolSolver, self, eqError)
        rep.solution = sol
        return rep

    # ---- phase I, sum of infeasibilities (ConstraintSet.scala:233-282, 488-545; Constraint.scala:101-159)
    def phase_I_SOI_ObjectiveFunction(self):
        """f(x, s) = s_1 + ... + s_p in dimension n + p (ConstraintSet.scala:233-249)."""
        n, p = self.dim, self.numConstraints
        a = np.zeros(n + p)
        a[n:] = 1.0
        return LinearObjectiveFunction(n + p, 0.0, a)

    def phase_I_SOI_Constraints(self, handle=None) -> "ConstraintSet":
        """g_j(x) - s_j <= ub_j and -s_j <= 0 in dimension n + p, with the feasible point
        (x, max(0.5, 1 + g_j(x) - ub_j)) at x = pointWhereDefined (ConstraintSet.scala:259-282).  Rows: linear
        g_j - s_j, then -s_j <= 0, then the quadratic constraints (the reference lists the positivity rows last)."""
        n, p, ml = self.dim, self.numConstraints, self.H.shape[0]
        N = n + p
        G = np.zeros((ml + p, N), order="F")
        G[:ml, :n] = self.H
        G[np.arange(ml), n + np.arange(ml)] = -1.0
        G[ml + np.arange(p), n + np.arange(p)] = -1.0
        r = np.concatenate([self.r if self.r is not None else np.zeros(ml), np.zeros(p)])
        u = np.concatenate([self.u, np.zeros(p)])
        quad = []
        for k, q in enumerate(self.quadratic):
            P1 = np.zeros((N, N))
            P1[:n, :n] = q.P
            a1 = np.zeros(N)
            a1[:n] = q.a
            a1[n + ml + k] = -1.0
            quad.append(QuadraticConstraint(str(q.id) + "_phase_I", N, q.ub, q.r, a1, P1))
        x = self.pointWhereDefined
        assert x is not None, "phase_I_SOI_Constraints needs pointWhereDefined"
        viol = self.valuesAt(x, handle) - self.ub_all()
        fp = np.concatenate([x, np.maximum(0.5, 1.0 + viol)])
        return ConstraintSet(G, u, fp, r, quad).addFeasiblePoint(fp)

    def phase_I_Analysis_SOI(self, eqs, pars=None, debugLevel=0, handle=None) -> "FeasibilityReport":
        """ConstraintSet.phase_I_Analysis_SOI (ConstraintSet.scala:511-545): full barrier solve of the SOI problem.
        isStrictlyFeasible follows the reference's test over `0 until n` (defect D9, see oracle)."""
        pars = pars if pars is not None else SolverParams.standardParams()
        n, p = self.dim, self.numConstraints
        eqs_soi = None if eqs is None else eqs.phase_I_SOI_EqualityConstraint(p)
        solver = BarrierSolver(self.phase_I_SOI_ObjectiveFunction(), self.phase_I_SOI_Constraints(handle), eqs_soi, pars,
                               None, handle)
        sol = solver.solve(debugLevel)
        x_feas, s_feas = sol.x[:n].copy(), sol.x[n:n + p].copy()
        eqError = None if eqs is None else eqs.errorAt(x_feas)
        strict = True
        for j in range(n):
            if not (s_feas[j] < 0):            # IndexError for j >= p, as the JVM would throw
                strict = False
                break
        strict = strict and ((0.0 if eqError is None else eqError) < pars.tolSolver)
        rep = FeasibilityReport(x_feas, s_feas, strict, self, eqError)
        rep.solution = sol
        return rep

    def withFeasiblePoint(self, eqs, pars=None, debugLevel=0, handle=None) -> "ConstraintSet":
        """ConstraintSet.withFeasiblePoint (ConstraintSet.scala:556-575)."""
        if self.feasiblePoint is not None:
            return self
        pars = pars if pars is not None else SolverParams.standardParams()
        rep = self.phase_I_Analysis(eqs, pars, debugLevel, handle)
        if not rep.isFeasible(pars.tolSolver):
            raise _lib.InfeasibleProblemException(rep.reasonWhyInfeasible(pars.tolSolver))
        return self.addFeasiblePoint(rep.x0)


@dataclass
class FeasibilityReport:
    """FeasibilityReport.scala:12-48."""
    x0: np.ndarray
    s: np.ndarray
    isStrictlyFeasible: bool
    constraintSet: ConstraintSet
    equalityConstraintError: Optional[float]
    solution: Optional["Solution"] = None

    def violatedConstraints(self, tol: float) -> List[int]:
        """indices of the constraints with g_i(x0) > ub_i + tol (Constraint.isSatisfiedWithTolerance)."""
        c = self.constraintSet
        return [int(i) for i in np.nonzero(~(c.valuesAt(self.x0) <= c.ub_all() + tol))[0]]

    def isFeasible(self, tol: float) -> bool:
        e = 0.0 if self.equalityConstraintError is None else self.equalityConstraintError
        return bool(np.max(self.s) < tol and e < tol)

    def reasonWhyInfeasible(self, tol: float) -> str:
        if self.isFeasible(tol):
            return "\nCannot determine if problem is feasible.\n"
        return ("\nProblem not feasible within tolerance %g\nFound point x0:\n%s\nviolates constraints:\n%s"
                "\nEqualityContraints, error: %s\n" % (tol, self.x0, self.violatedConstraints(tol),
                                                       self.equalityConstraintError or 0))


class EqualityConstraint:
    def __init__(self, A, b):
        self.A, self.b = fmat(A), fvec(b)
        assert self.A.shape[0] == self.b.shape[0]

    def errorAt(self, x) -> float:
        """||Ax - b|| (EqualityConstraint.scala:26)."""
        return float(np.linalg.norm(self.A @ fvec(x) - self.b))

    @property
    def solutionSpace(self):
        """EqualityConstraint.solutionSpace (EqualityConstraint.scala:21-23).  The reference computes this QR eagerly in
        the constructor although the main solver paths never use it; here it is lazy and runs on the device."""
        if getattr(self, "_space", None) is None:
            from .linalg import SolutionSpace
            self._space = SolutionSpace(self.A, self.b)
        return self._space

    @property
    def F(self):
        return self.solutionSpace.F

    @property
    def z0(self):
        return self.solutionSpace.z0

    def phase_I_SOI_EqualityConstraint(self, p: int) -> "EqualityConstraint":
        """[A, 0] u = b in dimension n + p (EqualityConstraint.scala:50-55)."""
        return EqualityConstraint(np.hstack([self.A, np.zeros((self.A.shape[0], int(p)))]), self.b)


class _DeviceProblem:
    """cvxb_problem: the uploaded descriptor.  Keeps the host arrays alive during the upload."""

    def __init__(self, objF, cnts: ConstraintSet, eqs: Optional[EqualityConstraint], handle):
        self.handle = handle if handle is not None else _lib.default_handle()
        n, m = cnts.dim, cnts.H.shape[0]
        assert objF.dim == n, "objective / constraint dimension mismatch"
        d = ProblemDesc()
        d.n, d.m, d.p = n, m, 0 if eqs is None else eqs.A.shape[0]
        d.objective = objF.kind
        d.obj_a = dptr(objF.a) if objF.a is not None else None
        d.obj_r = objF.r
        if objF.kind == OBJ_QUADRATIC:
            d.obj_P, d.obj_ldP = dptr(objF.P), n
        if objF.kind == OBJ_PNORM:
            d.obj_pow = objF.p
        if objF.kind == OBJ_KLDUAL:
            d.obj_P, d.obj_ldP, d.obj_k, d.obj_R = dptr(objF.B), n, objF.B.shape[1], dptr(objF.R)
        d.G, d.ldg = dptr(cnts.H), m
        d.g_r = dptr(cnts.r) if cnts.r is not None else None
        d.ub = dptr(cnts.u)
        if eqs is not None:
            assert eqs.A.shape[1] == n
            d.A, d.lda, d.b = dptr(eqs.A), eqs.A.shape[0], dptr(eqs.b)
        d.x_feasible = dptr(cnts.feasiblePoint) if cnts.feasiblePoint is not None else None
        d.x_defined = dptr(cnts.pointWhereDefined) if cnts.pointWhereDefined is not None else None
        mq = len(cnts.quadratic)
        d.mq = mq
        if mq:
            self._qP = np.ascontiguousarray(np.stack([np.asarray(q.P).T for q in cnts.quadratic]))      # packed column-major blocks
            self._qa = np.ascontiguousarray(np.stack([q.a for q in cnts.quadratic]))                      # mq x n row-major == n x mq column-major
            self._qr = np.array([q.r for q in cnts.quadratic], dtype=np.float64)
            self._qub = np.array([q.ub for q in cnts.quadratic], dtype=np.float64)
            d.q_P, d.q_a, d.q_r, d.q_ub = dptr(self._qP), dptr(self._qa), dptr(self._qr), dptr(self._qub)
        self.n, self.m, self.p = d.n, d.m + mq, d.p
        self._keep = (objF, cnts, eqs)
        self._p = C.c_void_p()
        check(self.handle.lib.cvxb_problem_create(self.handle._h, C.byref(d), C.byref(self._p)))

    def close(self):
        if getattr(self, "_p", None):
            self.handle.lib.cvxb_problem_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _solution_from_c(s: SolutionC, x, lam, nu) -> Solution:
    opt = lambda has, v: float(v) if has else None
    stages = int(s.outer_stages)
    return Solution(
        x=x, lam=lam if s.has_lambda else None, nu=nu if s.has_nu else None,
        newtonDecrement=opt(s.has_newtonDecrement, s.newtonDecrement), dualityGap=opt(s.has_dualityGap, s.dualityGap),
        equalityGap=opt(s.has_equalityGap, s.equalityGap), normGrad=opt(s.has_normGrad, s.normGrad),
        normDualResidual=opt(s.has_normDualResidual, s.normDualResidual), iter=int(s.iter), maxedOut=bool(s.maxedOut),
        objective=float(s.objective), outer_stages=stages, newton_steps=int(s.newton_steps),
        executed_newton_steps=int(s.executed_newton_steps),
        stage_newton_steps=[int(s.stage_newton_steps[i]) for i in range(min(stages, 128))],
        phase1_newton_steps=int(s.phase1_newton_steps), phase1_executed_steps=int(s.phase1_executed_steps),
        phase1_stages=int(s.phase1_stages), phase1_s=float(s.phase1_s), linesearch_trials=int(s.linesearch_trials),
        kkt_fallbacks=int(s.kkt_fallbacks), kkt_regularized=int(s.kkt_regularized), solve_ms=float(s.solve_ms))


class Solver:
    """Solver trait (Solver.scala:29-33)."""

    solverType = "BR"

    def __init__(self, objF, cnts, eqs=None, pars=None, logger=None, handle=None):
        self.objF, self.cnts, self.eqs = objF, cnts, eqs
        self.pars = pars if pars is not None else SolverParams.standardParams()
        self.problem = _DeviceProblem(objF, cnts, eqs, handle)
        self.handle = self.problem.handle

    def _run(self, fn):
        pr = self.problem
        x = np.zeros(pr.n)
        lam = np.zeros(max(pr.m, 1))
        nu = np.zeros(max(pr.p, 1))
        s = SolutionC()
        s.x, s.lam, s.nu = dptr(x), dptr(lam), dptr(nu)
        cp = self.pars.to_c(self.handle)
        check(fn(self.handle._h, pr._p, C.byref(cp), C.byref(s)))
        return _solution_from_c(s, x, lam[:pr.m], nu[:pr.p])

    def solve(self, debugLevel: int = 0) -> Solution:
        raise NotImplementedError


class _ReducedMixin:
    """Solver.reduced(sol) (BarrierSolver.scala:249-256, PrimalDualSolver.scala:699-712): the same problem in the variable
    u of x = z0 + F u.  As in the reference the returned solver works in u (Solution.x is u); `point(u)` maps back."""

    def reduced(self, sol):
        red = object.__new__(type(self))
        red.objF, red.cnts, red.eqs, red.pars = self.objF, self.cnts, None, self.pars
        red.handle = self.handle
        red.space = sol
        red.problem = _ReducedProblem(self.problem, sol, self.pars)
        return red

    def point(self, u):
        return self.space.point(u)


class _ReducedProblem:
    def __init__(self, base: "_DeviceProblem", sol, pars):
        self.handle = base.handle
        if sol.handle is not base.handle:
            raise ValueError("solution space and problem live on different handles")
        self._p = C.c_void_p()
        cp = pars.to_c(self.handle)
        check(self.handle.lib.cvxb_problem_reduce(self.handle._h, base._p, sol._s, C.byref(cp), C.byref(self._p)))
        self.n, self.m, self.p = sol.n - sol.p, base.m, 0
        self._keep = (base, sol)

    close = _DeviceProblem.close
    __del__ = _DeviceProblem.__del__


class BarrierSolver(_ReducedMixin, Solver):
    """BarrierSolver(objF, cnts, eqs, pars, logger).solve(debugLevel); phase I runs first when the
    constraint set has no feasible point (what OptimizationProblem.withoutFeasiblePoint does)."""
    solverType = "BR"

    def solve(self, debugLevel: int = 0) -> Solution:
        return self._run(self.handle.lib.cvxb_barrier_solve)

    def phase_I(self):
        """ConstraintSet.withFeasiblePoint: returns (x_feasible, phase-I Solution whose x is (x, s))."""
        pr = self.problem
        xf = np.zeros(pr.n)
        w = np.zeros(pr.n + 1)
        s = SolutionC()
        s.x = dptr(w)
        cp = self.pars.to_c(self.handle)
        check(self.handle.lib.cvxb_phase1(self.handle._h, pr._p, C.byref(cp), ptr(xf), C.byref(s)))
        return xf, _solution_from_c(s, w, None, None)

    def newton_direction(self, x, t):
        """One barrier Newton direction at (x, t): (H, grad, dx, nu, info)  -- per-step parity hook."""
        pr = self.problem
        x = fvec(x)
        H = np.empty((pr.n, pr.n), order="F")
        g = np.empty(pr.n)
        dx = np.empty(pr.n)
        nu = np.empty(max(pr.p, 1))
        info = KktInfo()
        cp = self.pars.to_c(self.handle)
        check(self.handle.lib.cvxb_barrier_newton_direction(self.handle._h, pr._p, C.byref(cp), ptr(x), float(t), ptr(H),
                                                            ptr(g), ptr(dx), ptr(nu), C.byref(info)))
        return H, g, dx, (nu[:pr.p] if pr.p else None), info


class PrimalDualSolver(_ReducedMixin, Solver):
    solverType = "PD"

    def solve(self, debugLevel: int = 0) -> Solution:
        return self._run(self.handle.lib.cvxb_pd_solve)

    def newton_direction(self, x, lam, nu, t):
        pr = self.problem
        x, lam = fvec(x), fvec(lam)
        nu = fvec(nu) if nu is not None else None
        H = np.empty((pr.n, pr.n), order="F")
        dx = np.empty(pr.n)
        dlam = np.empty(pr.m)
        dnu = np.empty(max(pr.p, 1))
        info = KktInfo()
        cp = self.pars.to_c(self.handle)
        check(self.handle.lib.cvxb_pd_newton_direction(self.handle._h, pr._p, C.byref(cp), ptr(x), ptr(lam), ptr(nu),
                                                       float(t), ptr(H), ptr(dx), ptr(dlam), ptr(dnu), C.byref(info)))
        return H, dx, dlam, (dnu[:pr.p] if pr.p else None), info


class OptimizationProblem:
    """OptimizationProblem.apply / withoutFeasiblePoint (OptimizationProblem.scala:133-196)."""

    def __init__(self, id, objF, ineqs: ConstraintSet, eqs: Optional[EqualityConstraint] = None, solverType="BR",
                 pars: Optional[SolverParams] = None, logger=None, debugLevel=0, handle=None):
        assert solverType in ("BR", "PD"), "solverType must be 'BR' or 'PD'"
        self.id, self.objectiveFunction = id, objF
        cls = BarrierSolver if solverType == "BR" else PrimalDualSolver
        self.solver = cls(objF, ineqs, eqs, pars, logger, handle)

    withoutFeasiblePoint = classmethod(lambda cls, *a, **k: cls(*a, **k))

    def solve(self, debugLevel: int = 0) -> Solution:
        return self.solver.solve(debugLevel)


class Dist_KL(OptimizationProblem):
    """Dist_KL.apply (Dist_KL.scala:270-315): min d_KL(x, uniform) s.t. Hx <= u, x >= 0 (positivity rows after the H
    rows), A x = r and sum x = 1 (stacked last); pointWhereDefined = 1/n, phase I first.  With Duality
    (Duality.scala:99-133): `solveDual` solves max L_*(z), lambda >= 0 in dimension rows(H) + rows(A) + 1 and maps
    the dual optimum back with primalOptimum -- the route the reference prefers for KL problems."""

    E_REF = 2.7182811828459045          # Dist_KL.scala:114 (sic; defect D8), used in vec_R

    def __init__(self, n, H=None, u=None, A=None, r=None, solverType="BR", pars=None, logger=None, debugLevel=0,
                 handle=None):
        assert H is not None or A is not None, "Must have some inequality or equality constraints"
        self.n = int(n)
        self.H = None if H is None else np.asarray(H, float)
        self.u = None if u is None else np.asarray(u, float)
        self.A = None if A is None else np.asarray(A, float)
        self.r = None if r is None else np.asarray(r, float)
        self._pars, self._handle = pars, handle
        Gpos = -np.eye(n)
        if H is not None:
            G, ub = np.vstack([self.H, Gpos]), np.concatenate([self.u, np.zeros(n)])
        else:
            G, ub = Gpos, np.zeros(n)
        ones = np.ones((1, n))
        if A is not None:
            Aeq, beq = np.vstack([self.A, ones]), np.concatenate([self.r, [1.0]])
        else:
            Aeq, beq = ones, np.array([1.0])
        cnts = ConstraintSet(G, ub, np.full(n, 1.0 / n))
        super().__init__("Dist_KL", KLObjectiveFunction(n), cnts, EqualityConstraint(Aeq, beq), solverType, pars, logger,
                         debugLevel, handle)

    # ---- Duality members (Dist_KL.scala:107-163)
    @property
    def numInequalities(self):
        return 0 if self.H is None else self.H.shape[0]

    @property
    def mat_B(self):
        ones = np.ones((1, self.n))
        Aext = ones if self.A is None else np.vstack([ones, self.A])          # A_with_probEQ: sum-to-one row first
        return Aext if self.H is None else np.vstack([self.H, Aext])

    @property
    def vec_w(self):
        rext = np.array([1.0]) if self.r is None else np.concatenate([[1.0], self.r])
        return rext if self.H is None else np.concatenate([self.u, rext])

    @property
    def vec_R(self):
        return np.full(self.n, 1.0 / (self.n * self.E_REF))

    def dualProblem(self, solverType="BR", pars=None, logger=None, debugLevel=0):
        """Duality.dualProblem (Duality.scala:99-112): min -L_*(z), lambda >= 0, start z = 0.001."""
        B = self.mat_B
        D, mI = B.shape[0], self.numInequalities
        G = np.zeros((mI, D))
        G[np.arange(mI), np.arange(mI)] = -1.0                                # Constraints.firstCoordinatesPositive
        cnts = ConstraintSet(G, np.zeros(mI), np.zeros(D)).addFeasiblePoint(np.full(D, 0.001))
        objF = DualKLObjectiveFunction(B, self.vec_w, self.vec_R)
        return OptimizationProblem("Dist_KL dual problem", objF, cnts, None, solverType,
                                   pars if pars is not None else self._pars, logger, debugLevel, self._handle)

    def solveDual(self, solverType="BR", pars=None, logger=None, debugLevel=0) -> Solution:
        """Duality.solveDual (Duality.scala:119-133): the returned Solution carries the PRIMAL optimum in x and the
        dual variables in lam / nu."""
        dP = self.dualProblem(solverType, pars, logger, debugLevel)
        solD = dP.solve(debugLevel)
        z = solD.x
        pr = dP.solver.problem
        xp = np.empty(self.n)
        check(pr.handle.lib.cvxb_kldual_primal_optimum(pr.handle._h, pr._p, ptr(xp)))
        mI = self.numInequalities
        solD.z = z
        solD.x, solD.lam, solD.nu = xp, z[:mI].copy(), (z[mI:].copy() if z.shape[0] > mI else None)
        return solD


def from_dict(prob: dict, solverType="BR", pars=None, handle=None) -> OptimizationProblem:
    """Problem dictionaries of oracle/problems.py (tests and bench) -> OptimizationProblem."""
    n = prob["n"]
    if prob["kind"] == "linear":
        objF = LinearObjectiveFunction(n, prob["r"], prob["a"])
    elif prob["kind"] == "quadratic":
        objF = QuadraticObjectiveFunction(n, prob["r"], prob["a"], prob["P"])
    elif prob["kind"] == "pnorm":
        objF = PNormObjectiveFunction(n, prob["pow"])
    else:
        objF = KLObjectiveFunction(n)
    quad = [QuadraticConstraint("q%d" % k, n, q["ub"], q["r"], q["a"], q["P"]) for k, q in enumerate(prob.get("quad") or [])]
    cnts = ConstraintSet(prob["G"], prob["ub"], prob["xdef"], prob.get("rvec"), quad)
    if prob.get("x0") is not None:
        cnts = cnts.addFeasiblePoint(prob["x0"])
    eqs = EqualityConstraint(prob["A"], prob["b"]) if prob.get("A") is not None else None
    return OptimizationProblem(prob.get("id", "problem"), objF, cnts, eqs, solverType, pars, None, 0, handle)

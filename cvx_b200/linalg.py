"""Host-side mirror of the reference's dense linear-algebra classes (seam B), same names and
argument meaning, every call executed by libcvxb on the GPU:

  KKTSystem(H, A, q, b).solve(delta, logger, tol, debugLevel) -> (x, w)     KKTSystem.scala:43-66
  KKTSystem.solveWithCholFactor(L, A, q, b, logger, tol, debugLevel)        KKTSystem.scala:99-167
  SymmetricLinearSystem(H, r, logger).solve(tol, debugLevel)                SymmetricLinearSystem.scala:15-56
  MatrixUtils.choleskySolve / ruizEquilibrate / regularizedCholesky / triangularSolve /
  forwardSolve / backSolve                                                  MatrixUtils.scala:240-516

numpy arrays stand in for Breeze DenseMatrix / DenseVector (converted to column-major FP64 at the
boundary, which is Breeze's layout).  Errors surface as the reference's exception types (_lib.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KktInfo, check, fmat, fvec, ptr


def _h(handle):
    return handle if handle is not None else _lib.default_handle()


class KKTSystem:
    """System  Hx + A'w = -q,  Ax = b  (KKTSystem.scala:23-33)."""

    def __init__(self, H, A, q, b, handle=None):
        self.H, self.A, self.q, self.b = fmat(H), fmat(A), fvec(q), fvec(b)
        n = self.H.shape[1]
        # constructor asserts, KKTSystem.scala:30-33
        if self.H.shape[0] != n:
            raise _lib.DimensionMismatch("Matrix M not square: n=M.cols=%d, M.rows=%d" % (n, self.H.shape[0]))
        if self.A.shape[1] != n:
            raise _lib.DimensionMismatch("Dimension mismatch A.cols=%d not equal to n=M.cols=%d" % (self.A.shape[1], n))
        if self.q.shape[0] != n or self.b.shape[0] != self.A.shape[0]:
            raise _lib.DimensionMismatch("Dimension mismatch in q or b")
        self.handle = _h(handle)
        self.info = KktInfo()

    def solve(self, delta=1e-6, logger=None, tol=1e-1, debugLevel=0):
        """`delta` is accepted and ignored exactly as in the reference (defect D5)."""
        n, p = self.H.shape[1], self.A.shape[0]
        x = np.empty(n)
        w = np.empty(p)
        hd = self.handle
        check(hd.lib.cvxb_kkt_solve(hd._h, n, p, ptr(self.H), self.H.shape[0], ptr(self.A), max(p, 1), ptr(self.q),
                                    ptr(self.b), float(tol), ptr(x), ptr(w), C.byref(self.info)))
        return x, w

    @staticmethod
    def solveWithCholFactor(L, A, q, b, logger=None, tol=1e-1, debugLevel=0, handle=None, info=None):
        L, A, q, b = fmat(L), fmat(A), fvec(q), fvec(b)
        n, p = L.shape[1], A.shape[0]
        if L.shape[0] != n or A.shape[1] != n:
            raise _lib.DimensionMismatch("solveWithCholFactor: dimension mismatch")
        x = np.empty(n)
        w = np.empty(p)
        hd = _h(handle)
        info = info if info is not None else KktInfo()
        check(hd.lib.cvxb_kkt_solve_with_chol_factor(hd._h, n, p, ptr(L), n, ptr(A), p, ptr(q), ptr(b), float(tol),
                                                     ptr(x), ptr(w), C.byref(info)))
        return x, w


class KKTData:
    """KKTData(H, A, g, r, nullIndices) (KKTData.scala:31-91):  Hx + A'nu = -g, Ax = r, with the elimination of the
    variables the system does not depend on.  `solveReduced` = reduced -> KKTSystem.solve -> paddVector in one
    device call (cvxb_kkt_solve_reduced); it returns (x padded with zeros, nu) and sets `nullIndices`."""

    def __init__(self, H, A, g, r, nullIndices=None, handle=None):
        self.H, self.A, self.g, self.r = fmat(H), fmat(A), fvec(g), fvec(r)
        n = self.H.shape[1]
        if self.H.shape[0] != n:
            raise _lib.DimensionMismatch("Matrix H not square, n=H.cols=%d, H.rows=%d" % (n, self.H.shape[0]))
        if self.A.shape[1] != n:
            raise _lib.DimensionMismatch("A.cols=%d not equal to n=M.rows=%d" % (self.A.shape[1], n))
        self.nullIndices = None if nullIndices is None else list(nullIndices)
        self.handle = _h(handle)
        self.info = KktInfo()

    def solveReduced(self, delta=1e-6, logger=None, tol=1e-1, debugLevel=0):
        n, p = self.H.shape[1], self.A.shape[0]
        x = np.empty(n)
        w = np.empty(p)
        idx = (C.c_int * n)()
        cnt = C.c_int(0)
        hd = self.handle
        check(hd.lib.cvxb_kkt_solve_reduced(hd._h, n, p, ptr(self.H), n, ptr(self.A), p, ptr(self.g), ptr(self.r), float(tol),
                                            ptr(x), ptr(w), idx, C.byref(cnt), C.byref(self.info)))
        self.nullIndices = [int(idx[k]) for k in range(cnt.value)] if cnt.value else None
        return x, w

    @staticmethod
    def paddVector(x, nullIndices):
        """KKTData.paddVector (KKTData.scala:105-127)."""
        x = fvec(x)
        z = np.zeros(x.shape[0] + len(nullIndices))
        z[np.setdiff1d(np.arange(z.shape[0]), np.asarray(nullIndices, dtype=int))] = x
        return z


class SymmetricLinearSystem:
    def __init__(self, H, r, logger=None, handle=None):
        self.H, self.r = fmat(H), fvec(r)
        if self.H.shape[0] != self.H.shape[1] or self.r.shape[0] != self.H.shape[0]:
            raise _lib.DimensionMismatch("SymmetricLinearSystem: dimension mismatch")
        self.handle = _h(handle)
        self.info = KktInfo()

    def solve(self, tol=1e-1, debugLevel=0):
        n = self.H.shape[0]
        x = np.empty(n)
        hd = self.handle
        check(hd.lib.cvxb_symmetric_solve(hd._h, n, ptr(self.H), n, ptr(self.r), float(tol), ptr(x),
                                          C.byref(self.info)))
        return x


class SolutionSpace:
    """SolutionSpace(A, b) (SolutionSpace.scala:20-33): all solutions of the underdetermined full-rank system Ax = b as
    x = z0 + F u, z0 the minimum-norm solution, the columns of F an orthonormal basis of ker(A).  The QR of A' runs
    on the device (blocked Householder, cvxb_solution_space_create) and (z0, F) stay device resident for
    BarrierSolver.reduced; `.z0` / `.F` download them on first use."""

    def __init__(self, A, b, handle=None):
        self.A, self.b = fmat(A), fvec(b)
        p, n = self.A.shape
        if p != self.b.shape[0]:
            raise _lib.DimensionMismatch("SolutionSpace: A.rows != b.length")
        if not p < n:
            raise _lib.DimensionMismatch("SolutionSpace: need A.rows < A.cols")
        self.handle = _h(handle)
        self._s = C.c_void_p()
        check(self.handle.lib.cvxb_solution_space_create(self.handle._h, p, n, ptr(self.A), p, ptr(self.b), C.byref(self._s)))
        self.n, self.p = n, p
        self._z0 = self._F = None

    @classmethod
    def from_basis(cls, z0, F, handle=None) -> "SolutionSpace":
        """The affine map x = z0 + F u handed in explicitly (Solver.affineTransformed(z0, F, u0), Solver.scala:33-46):
        cvxb_solution_space_from_basis.  F is n x k; `parameter` computes F'(x - z0)."""
        self = object.__new__(cls)
        F, z0 = fmat(F), fvec(z0)
        n, k = F.shape
        if z0.shape[0] != n:
            raise _lib.DimensionMismatch("affineTransformed: z0.length != F.rows")
        self.A = self.b = None
        self.handle = _h(handle)
        self._s = C.c_void_p()
        check(self.handle.lib.cvxb_solution_space_from_basis(self.handle._h, n, k, ptr(z0), ptr(F), n, C.byref(self._s)))
        self.n, self.p = n, n - k
        self._z0, self._F = z0, F
        return self

    def _fetch(self):
        if self._z0 is None:
            z0 = np.empty(self.n)
            F = np.empty((self.n, self.n - self.p), order="F")
            check(self.handle.lib.cvxb_solution_space_get(self.handle._h, self._s, ptr(z0), ptr(F), self.n))
            self._z0, self._F = z0, F

    @property
    def z0(self):
        self._fetch()
        return self._z0

    @property
    def F(self):
        self._fetch()
        return self._F

    def parameter(self, x0):
        """u0 with x0 = z0 + F u0 when A x0 = b:  F'(x0 - z0)  (SolutionSpace.scala:32)."""
        u = np.empty(self.n - self.p)
        check(self.handle.lib.cvxb_solution_space_parameter(self.handle._h, self._s, ptr(fvec(x0)), ptr(u)))
        return u

    def point(self, u):
        """x = z0 + F u."""
        x = np.empty(self.n)
        check(self.handle.lib.cvxb_solution_space_map(self.handle._h, self._s, ptr(fvec(u)), ptr(x)))
        return x

    def close(self):
        if getattr(self, "_s", None):
            self.handle.lib.cvxb_solution_space_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MatrixUtils:
    """Static methods of the reference's MatrixUtils object that lie on the hot path."""

    @staticmethod
    def solveUnderdetermined(A, b, handle=None):
        """(z0, F) of MatrixUtils.solveUnderdetermined (MatrixUtils.scala:536-550)."""
        A, b = fmat(A), fvec(b)
        p, n = A.shape
        if not (p == b.shape[0] and p < n):
            raise _lib.DimensionMismatch("solveUnderdetermined: need A.rows == b.length and A.rows < A.cols")
        z0 = np.empty(n)
        F = np.empty((n, n - p), order="F")
        hd = _h(handle)
        check(hd.lib.cvxb_solve_underdetermined(hd._h, p, n, ptr(A), p, ptr(b), ptr(z0), ptr(F), n))
        return z0, F

    @staticmethod
    def choleskySolve(H, b, logger=None, tol=1e-1, debugLevel=0, handle=None, info=None):
        H, b = fmat(H), fvec(b)
        n = H.shape[0]
        if H.shape[1] != n or b.shape[0] != n:
            raise _lib.DimensionMismatch("choleskySolve: dimension mismatch")
        x = np.empty(n)
        hd = _h(handle)
        info = info if info is not None else KktInfo()
        check(hd.lib.cvxb_cholesky_solve(hd._h, n, ptr(H), n, ptr(b), float(tol), ptr(x), C.byref(info)))
        return x

    @staticmethod
    def ruizEquilibrate(H, handle=None, return_sweeps=False):
        H = fmat(H)
        n = H.shape[0]
        if H.shape[1] != n:
            raise _lib.DimensionMismatch("ruizEquilibrate: not square")
        d = np.empty(n)
        Q = np.empty((n, n), order="F")
        sweeps = C.c_int()
        hd = _h(handle)
        check(hd.lib.cvxb_ruiz_equilibrate(hd._h, n, ptr(H), n, ptr(d), ptr(Q), n, C.byref(sweeps)))
        return (d, Q, sweeps.value) if return_sweeps else (d, Q)

    @staticmethod
    def regularizedCholesky(Q, handle=None, info=None):
        Q = fmat(Q)
        n = Q.shape[0]
        L = np.empty((n, n), order="F")
        hd = _h(handle)
        info = info if info is not None else KktInfo()
        check(hd.lib.cvxb_regularized_cholesky(hd._h, n, ptr(Q), n, ptr(L), n, C.byref(info)))
        return L

    @staticmethod
    def triangularSolve(A, Ltype, B, handle=None):
        assert Ltype in ("L", "U"), "Triangular matrix type must be 'L' or 'U'"
        A = fmat(A)
        B = np.asarray(B, dtype=np.float64)
        vec = B.ndim == 1
        X = np.array(B.reshape(-1, 1) if vec else B, dtype=np.float64, order="F")
        n = A.shape[0]
        if A.shape[1] != n or X.shape[0] != n:
            raise _lib.DimensionMismatch("triangularSolve: dimension mismatch")
        hd = _h(handle)
        check(hd.lib.cvxb_triangular_solve(hd._h, Ltype.encode(), n, X.shape[1], ptr(A), n, ptr(X), n))
        return X[:, 0].copy() if vec else X

    @staticmethod
    def forwardSolve(L, b, handle=None):
        """MatrixUtils.forwardSolve (MatrixUtils.scala:383-403): L lower triangular, single RHS."""
        return MatrixUtils.triangularSolve(L, "L", np.asarray(b, dtype=np.float64), handle)

    @staticmethod
    def backSolve(U, b, handle=None):
        """MatrixUtils.backSolve (MatrixUtils.scala:410-430): U upper triangular, single RHS."""
        return MatrixUtils.triangularSolve(U, "U", np.asarray(b, dtype=np.float64), handle)


def dgemm(a_kc, b_kc, M, N, K, alpha, A, B, beta, Cm, tri=0, handle=None):
    """Test hook for the DMMA GEMM kernel (cvxb_test_dgemm).  A, B are the raw column-major
    storage arrays: A is (K x M) if a_kc else (M x K); B is (K x N) if b_kc else (N x K)."""
    A, B = fmat(A), fmat(B)
    Cm = np.array(Cm, dtype=np.float64, order="F")
    hd = _h(handle)
    check(hd.lib.cvxb_test_dgemm(hd._h, int(a_kc), int(b_kc), M, N, K, float(alpha), ptr(A), A.shape[0], ptr(B),
                                 B.shape[0], float(beta), ptr(Cm), Cm.shape[0], int(tri)))
    return Cm

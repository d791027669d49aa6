"""ctypes binding of libcvxb.so (include/cvxb.h).  No compute happens in Python: every call goes to
the CUDA library, and loading fails loudly when the library or a GPU is missing (there is no CPU path).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcvxb.so")

# cvxb_status
OK, ELINSOLVE, EUNSOLVABLE, ELINESEARCH, ENOTFEASIBLE, EINFEASIBLE, EDIM, ENOTSYMMETRIC, ECUDA, EINVAL, ENOTIMPL = range(11)
FLAG_DEVICE_PTRS = 1
OBJ_LINEAR, OBJ_QUADRATIC, OBJ_KL, OBJ_KLDUAL, OBJ_PNORM = 0, 1, 2, 3, 4


# ---- exceptions: the reference's exception types (SURVEY.md 8b "Error conventions") ----------------
class CvxbError(RuntimeError):
    status = -1


class LinSolveException(CvxbError):
    """cvx.LinSolveException (LinSolveException.scala:11-17)."""
    status = ELINSOLVE


class UnsolvableSystemException(CvxbError):
    status = EUNSOLVABLE


class LineSearchFailedException(CvxbError):
    """cvx.LineSearchFailedException (PD) / breeze NotConvergedException(Breakdown) (barrier)."""
    status = ELINESEARCH


class NotStrictlyFeasible(CvxbError, ValueError):
    """IllegalArgumentException of the barrier function family (BarrierSolver.scala:284)."""
    status = ENOTFEASIBLE


class InfeasibleProblemException(CvxbError):
    status = EINFEASIBLE


class DimensionMismatch(CvxbError, AssertionError):
    status = EDIM


class MatrixNotSymmetricException(CvxbError):
    status = ENOTSYMMETRIC


class CudaError(CvxbError):
    status = ECUDA


class NotImplementedOnDevice(CvxbError, NotImplementedError):
    status = ENOTIMPL


_EXC = {ELINSOLVE: LinSolveException, EUNSOLVABLE: UnsolvableSystemException, ELINESEARCH: LineSearchFailedException,
        ENOTFEASIBLE: NotStrictlyFeasible, EINFEASIBLE: InfeasibleProblemException, EDIM: DimensionMismatch,
        ENOTSYMMETRIC: MatrixNotSymmetricException, ECUDA: CudaError, EINVAL: CvxbError, ENOTIMPL: NotImplementedOnDevice}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Params(C.Structure):
    """cvxb_params: SolverParams.scala:24-46 + the constants hard-coded in the solvers."""
    _fields_ = [("maxIter", C.c_int), ("alpha", C.c_double), ("beta", C.c_double), ("tolSolver", C.c_double),
                ("tolEqSolve", C.c_double), ("tolFeas", C.c_double), ("delta", C.c_double), ("mu", C.c_double),
                ("t0", C.c_double), ("ruizMaxSweeps", C.c_int), ("ruizTol", C.c_double), ("cholRegDelta", C.c_double),
                ("cholMinDiag", C.c_double), ("newtonRegDelta", C.c_double), ("phase1EqTol", C.c_double),
                ("pdStepFraction", C.c_double), ("bugCompat", C.c_int), ("stepLimit", C.c_longlong)]


class KktInfo(C.Structure):
    _fields_ = [("path", C.c_int), ("regularized", C.c_int), ("ruiz_sweeps", C.c_int), ("chol_info", C.c_int),
                ("min_diag", C.c_double), ("err1", C.c_double), ("err2", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ProblemDesc(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("p", C.c_int), ("objective", C.c_int), ("obj_a", _dp),
                ("obj_r", C.c_double), ("obj_P", _dp), ("obj_ldP", C.c_int), ("G", _dp), ("ldg", C.c_int),
                ("g_r", _dp), ("ub", _dp), ("A", _dp), ("lda", C.c_int), ("b", _dp), ("x_feasible", _dp),
                ("x_defined", _dp), ("mq", C.c_int), ("q_P", _dp), ("q_a", _dp), ("q_r", _dp), ("q_ub", _dp),
                ("obj_k", C.c_int), ("obj_R", _dp), ("obj_pow", C.c_double)]


class SolutionC(C.Structure):
    _fields_ = [("x", _dp), ("lam", _dp), ("nu", _dp), ("has_lambda", C.c_int), ("has_nu", C.c_int),
                ("newtonDecrement", C.c_double), ("has_newtonDecrement", C.c_int),
                ("dualityGap", C.c_double), ("has_dualityGap", C.c_int),
                ("equalityGap", C.c_double), ("has_equalityGap", C.c_int),
                ("normGrad", C.c_double), ("has_normGrad", C.c_int),
                ("normDualResidual", C.c_double), ("has_normDualResidual", C.c_int),
                ("iter", C.c_int), ("maxedOut", C.c_int), ("objective", C.c_double), ("outer_stages", C.c_int),
                ("newton_steps", C.c_longlong), ("executed_newton_steps", C.c_longlong),
                ("phase1_newton_steps", C.c_longlong), ("phase1_executed_steps", C.c_longlong), ("phase1_stages", C.c_int),
                ("phase1_s", C.c_double), ("linesearch_trials", C.c_longlong), ("kkt_fallbacks", C.c_int),
                ("kkt_regularized", C.c_int), ("stage_newton_steps", C.c_int * 128), ("solve_ms", C.c_double)]


class BatchDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("n", C.c_int), ("m", C.c_int), ("p", C.c_int), ("objective", _ip), ("pcount", _ip),
                ("obj_a", _dp), ("obj_r", _dp), ("obj_P", _dp), ("G", _dp), ("ub", _dp), ("A", _dp), ("b", _dp), ("x0", _dp),
                ("phase1", _ip)]


class BatchResult(C.Structure):
    _fields_ = [("x", _dp), ("status", _ip), ("newton_steps", _ip), ("outer_stages", _ip), ("objective", _dp),
                ("duality_gap", _dp), ("equality_gap", _dp), ("solve_ms", C.c_double), ("stage_newton_steps", _ip),
                ("cycles", C.POINTER(C.c_longlong)), ("phase1_newton_steps", _ip), ("phase1_stages", _ip), ("phase1_s", _dp)]


# every symbol include/cvxb.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "cvxb_create": (C.c_int, [C.c_int, _vp, C.c_uint, C.POINTER(_vp)]),
    "cvxb_destroy": (C.c_int, [_vp]),
    "cvxb_synchronize": (C.c_int, [_vp]),
    "cvxb_last_error": (C.c_char_p, []),
    "cvxb_version": (C.c_char_p, []),
    "cvxb_launch_count": (C.c_longlong, [_vp]),
    "cvxb_status_read_count": (C.c_longlong, [_vp]),
    "cvxb_default_params": (C.c_int, [C.POINTER(Params)]),
    "cvxb_profile_enable": (C.c_int, [_vp, C.c_int]),
    "cvxb_profile_read": (C.c_int, [_vp, C.POINTER(C.c_longlong), _dp, _dp]),
    "cvxb_profile_read_range": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_longlong), _dp, _dp]),
    "cvxb_batch_device_records": (C.c_int, [_vp, C.POINTER(_vp), _ip]),
    "cvxb_kkt_solve": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_double, _vp, _vp,
                                 C.POINTER(KktInfo)]),
    "cvxb_kkt_solve_with_chol_factor": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp,
                                                  C.c_double, _vp, _vp, C.POINTER(KktInfo)]),
    "cvxb_cholesky_solve": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, _vp, C.POINTER(KktInfo)]),
    "cvxb_symmetric_solve": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, _vp, C.POINTER(KktInfo)]),
    "cvxb_ruiz_equilibrate": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _ip]),
    "cvxb_regularized_cholesky": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.POINTER(KktInfo)]),
    "cvxb_triangular_solve": (C.c_int, [_vp, C.c_char, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int]),
    "cvxb_problem_create": (C.c_int, [_vp, C.POINTER(ProblemDesc), C.POINTER(_vp)]),
    "cvxb_problem_destroy": (C.c_int, [_vp]),
    "cvxb_kldual_primal_optimum": (C.c_int, [_vp, _vp, _vp]),
    "cvxb_kkt_solve_reduced": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_double, _vp, _vp,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(KktInfo)]),
    "cvxb_solution_space_create": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.POINTER(C.c_void_p)]),
    "cvxb_stage_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "cvxb_stage_destroy": (C.c_int, [_vp]),
    "cvxb_stage_buffer": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), _ip]),
    "cvxb_stage_push": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "cvxb_stage_wait": (C.c_int, [_vp, C.c_int]),
    "cvxb_stage_set_equalities": (C.c_int, [_vp, _vp, C.c_int]),
    "cvxb_stage_cholesky_solve": (C.c_int, [_vp, _vp, C.c_double, _vp, C.POINTER(KktInfo)]),
    "cvxb_stage_kkt_solve": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp, _vp, C.POINTER(KktInfo)]),
    "cvxb_solution_space_from_basis": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.POINTER(C.c_void_p)]),
    "cvxb_solution_space_destroy": (C.c_int, [_vp]),
    "cvxb_solution_space_get": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "cvxb_solution_space_parameter": (C.c_int, [_vp, _vp, _vp, _vp]),
    "cvxb_solution_space_map": (C.c_int, [_vp, _vp, _vp, _vp]),
    "cvxb_solve_underdetermined": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, C.c_int]),
    "cvxb_problem_reduce": (C.c_int, [_vp, _vp, _vp, C.POINTER(Params), C.POINTER(C.c_void_p)]),
    "cvxb_constraint_values": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(C.c_int)]),
    "cvxb_phase1": (C.c_int, [_vp, _vp, C.POINTER(Params), _vp, C.POINTER(SolutionC)]),
    "cvxb_barrier_solve": (C.c_int, [_vp, _vp, C.POINTER(Params), C.POINTER(SolutionC)]),
    "cvxb_pd_solve": (C.c_int, [_vp, _vp, C.POINTER(Params), C.POINTER(SolutionC)]),
    "cvxb_barrier_newton_direction": (C.c_int, [_vp, _vp, C.POINTER(Params), _vp, C.c_double, _vp, _vp, _vp, _vp,
                                                C.POINTER(KktInfo)]),
    "cvxb_pd_newton_direction": (C.c_int, [_vp, _vp, C.POINTER(Params), _vp, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp,
                                           C.POINTER(KktInfo)]),
    "cvxb_batch_create": (C.c_int, [_vp, C.POINTER(BatchDesc), C.POINTER(_vp)]),
    "cvxb_batch_destroy": (C.c_int, [_vp]),
    "cvxb_batch_barrier_solve": (C.c_int, [_vp, _vp, C.POINTER(Params), C.POINTER(BatchResult)]),
    "cvxb_test_dgemm": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _vp, C.c_int, _vp,
                                  C.c_int, C.c_double, _vp, C.c_int, C.c_int]),
    "cvxb_bench_kernel": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "cvxb_debug_set_schedule": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "cvxb_debug_dag_blocks": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
}

_lib = None


def load():
    """Loads libcvxb.so (RTLD_GLOBAL not needed) and types every symbol.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libcvxb.so is not built (%s); run `python -m cvx_b200.build` or "
                          "__graft_entry__.build(). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status == OK:
        return
    msg = load().cvxb_last_error().decode("utf-8", "replace")
    raise _EXC.get(status, CvxbError)("cvxb status %d: %s" % (status, msg))


def fmat(a) -> np.ndarray:
    """float64 column-major (Breeze DenseMatrix layout) copy/view of a 2-D array."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def fvec(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Handle:
    """cvxb_handle: one CUDA device + stream.  Not thread-safe (one handle per thread)."""

    def __init__(self, device: int = 0, stream=None, device_ptrs: bool = False):
        lib = load()
        h = C.c_void_p()
        check(lib.cvxb_create(int(device), stream, FLAG_DEVICE_PTRS if device_ptrs else 0, C.byref(h)))
        self._h = h
        self.device = device
        self.lib = lib

    def default_params(self) -> Params:
        p = Params()
        check(self.lib.cvxb_default_params(C.byref(p)))
        return p

    @property
    def launches(self) -> int:
        return int(self.lib.cvxb_launch_count(self._h))

    @property
    def status_reads(self) -> int:
        """Host round trips so far (status block read + stream synchronisation)."""
        return int(self.lib.cvxb_status_read_count(self._h))

    def synchronize(self):
        check(self.lib.cvxb_synchronize(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cvxb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def profile_enable(self, on: bool = True):
        check(self.lib.cvxb_profile_enable(self._h, int(on)))

    def profile_read(self):
        """(launches, total ms, total algorithmic flops) of the timed Hessian-assembly SYRK launches."""
        n, ms, fl = C.c_longlong(), C.c_double(), C.c_double()
        check(self.lib.cvxb_profile_read(self._h, C.byref(n), C.byref(ms), C.byref(fl)))
        return n.value, ms.value, fl.value

    PROF_RANGES = {"hessian_syrk": 0, "chol_trailing_update": 1, "factor_h_with_trsm": 2, "schur_syrk": 3, "ruiz": 4,
                   "gemv_g": 5, "chol_lookahead_phases": 6, "chol_trsm_right": 7}

    def profile_read_range(self, which):
        """(count, total ms, total algorithmic work) of one timed range of the step (cvxb_profile_read_range)."""
        rid = self.PROF_RANGES[which] if isinstance(which, str) else int(which)
        n, ms, wk = C.c_longlong(), C.c_double(), C.c_double()
        check(self.lib.cvxb_profile_read_range(self._h, rid, C.byref(n), C.byref(ms), C.byref(wk)))
        return n.value, ms.value, wk.value

    def set_schedule(self, dag_block: int = -1, dag_min_n: int = -1, dag_reserve: int = -1):
        """Tile-DAG schedule of the big factorisations (cvxb_debug_set_schedule); negative = keep, dag_block 0 = off."""
        check(self.lib.cvxb_debug_set_schedule(self._h, dag_block, dag_min_n, dag_reserve))

    # measurement helper (bench.py)
    def bench_kernel(self, which: int, n: int, k: int, reps: int):
        ms = C.c_double()
        work = C.c_double()
        check(self.lib.cvxb_bench_kernel(self._h, which, n, k, reps, C.byref(ms), C.byref(work)))
        return ms.value, work.value


_default = {}


def default_handle(device: int = 0) -> Handle:
    if device not in _default:
        _default[device] = Handle(device)
    return _default[device]

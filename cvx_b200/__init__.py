"""cvx_b200 -- B200-native (sm_100a) implementation of the interior-point Newton/KKT hot path of
spyqqqdia/cvx, behind the reference's own solver API.  All numerics run in libcvxb.so (hand-written
CUDA, FP64 DMMA tensor cores); this package is the thin host-side mirror of the reference's classes.
There is no CPU path: importing the solver classes without the built library or a GPU fails loudly."""
from . import _lib  # noqa: F401
from ._lib import (CvxbError, LinSolveException, UnsolvableSystemException, LineSearchFailedException,  # noqa: F401
                   NotStrictlyFeasible, InfeasibleProblemException, DimensionMismatch, Handle, default_handle)
from .linalg import KKTSystem, KKTData, SolutionSpace, SymmetricLinearSystem, MatrixUtils  # noqa: F401

__version__ = "0.1.0"
from .solvers import (SolverParams, Solution, LinearObjectiveFunction, QuadraticObjectiveFunction,  # noqa: F401,E402
                      KLObjectiveFunction, DualKLObjectiveFunction, PNormObjectiveFunction, ConstraintSet, QuadraticConstraint, EqualityConstraint, BarrierSolver, PrimalDualSolver,
                      OptimizationProblem, Dist_KL, FeasibilityReport, from_dict)
from .batched import BatchedBarrierSolver, BatchSolution, pack_problems, shard_range, gather_solutions  # noqa: F401,E402
from . import generic  # noqa: F401,E402
from .generic import StagedSystem  # noqa: F401,E402

import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""One resident kernel of the Newton step timed alone (cvxb_bench_kernel), for `ncu --set full` captures:
  python tools/gpu_kernels.py <which> <n> <k> [reps]
which: 1 Hessian SYRK (TN), 2 Cholesky trailing update (NT), 3 blocked Cholesky, 5 forward+backward single-RHS
solves, 6 Ruiz equilibration (20 sweeps enqueued), 7 gemv_n over a k x n matrix, 8 gemv_t."""
from cvx_b200 import _lib
h = _lib.default_handle()
which, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ms, work = h.bench_kernel(which, n, k, reps)
unit = "GB/s" if which in (4, 5, 6, 7, 8) else "TFLOP/s"
rate = work / ms / (1e6 if unit == "GB/s" else 1e9)
print("kernel %d n=%d k=%d: %.4f ms per launch, %.2f %s" % (which, n, k, ms, rate, unit))

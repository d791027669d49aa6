import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# blocked Cholesky timings (bench_kernel 3); CVXB_RL_MAX_N picks the largest n of the look-ahead schedule
from cvx_b200 import _lib
h = _lib.default_handle()
for n in [int(a) for a in sys.argv[1:]] or [2000, 4096, 8192]:
    ms, fl = h.bench_kernel(3, n, 0, 3)
    print("potrf n=%d (RL_MAX_N=%s): %.3f ms, %.2f TFLOP/s" % (n, os.environ.get("CVXB_RL_MAX_N", "default"), ms, fl / ms / 1e9), flush=True)

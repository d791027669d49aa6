import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Time line of one tile-DAG factorisation (CVXB_DAG_TRACE=1 python tools/gpu_dag_trace.py n r block reserve)"""
os.environ["CVXB_DAG_TRACE"] = "1"
from cvx_b200 import _lib
h = _lib.default_handle()
n, r, blk, res = [int(a) for a in sys.argv[1:5]]
h.set_schedule(blk, blk + 1, res)
ms, fl = h.bench_kernel(9 if r > 0 else 3, n, r, 1)
print("n=%d r=%d dag_block=%d reserve=%d: %.3f ms, %.2f TFLOP/s" % (n, r, blk, res, ms, fl / ms / 1e9), flush=True)

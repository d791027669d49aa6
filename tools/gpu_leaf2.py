import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from cvx_b200 import _lib
h = _lib.default_handle()
lib = h.lib
lib.cvxb_debug_leaf_clocks.argtypes = [C.POINTER(C.c_longlong), C.c_int]
buf = (C.c_longlong * 8)()
h.bench_kernel(3, 128, 0, 2)
lib.cvxb_debug_leaf_clocks(buf, 1)
ms, _ = h.bench_kernel(3, 128, 0, 10)
lib.cvxb_debug_leaf_clocks(buf, 1)
names = ["load", "zero-fill + diag 0", "panels", "windows (trailing || next diag || inverse pieces)", "last diag-inverse", "remaining inverse levels", "store"]
tot = sum(buf[:7])
print("potrf_128 %.1f us per call; leaf clocks per call (11 calls):" % (ms * 1e3))
for n, v in zip(names, buf[:7]):
    print("  %-56s %8.0f cycles  %5.1f%%" % (n, v / 11, 100 * v / tot))
print("  total %.0f cycles = %.1f us @1.965GHz" % (tot / 11, tot / 11 / 1965))

import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the dominant kernel at the C2 shape (Hessian-assembly SYRK n=2000, k=m=4000) and at the C4 shape
import sys
from cvx_b200 import _lib
h = _lib.default_handle()
n, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 4000)
ms, fl = h.bench_kernel(1, n, k, 3)
print("syrk_tn n=%d k=%d: %.4f ms/launch, %.2f TFLOP/s" % (n, k, ms, fl / ms / 1e9))

import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvx_b200 import _lib
h = _lib.default_handle()
print(h.bench_kernel(3, 128, 0, 3))

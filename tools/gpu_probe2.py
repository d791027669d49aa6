import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json, sys
from cvx_b200 import _lib
h = _lib.default_handle()
out = {}
for n in [128, 256, 1024, 2000, 4096, 8192]:
    ms, fl = h.bench_kernel(3, n, 0, 5)
    out["potrf_%d" % n] = dict(ms=round(ms,4), tflops=round(fl / ms / 1e9,3))
print(json.dumps(out))

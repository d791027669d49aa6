import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
from cvx_b200 import _lib
h = _lib.default_handle()
out = {}
for n in [2000, 8192]:
    ms, by = h.bench_kernel(5, n, 0, 5); out["trsv_fwd_bwd_%d" % n] = dict(ms=round(ms, 4), gbs=round(by / ms / 1e6, 1))
    ms, by = h.bench_kernel(6, n, 0, 5); out["ruiz20_%d" % n] = dict(ms=round(ms, 4))
    ms, fl = h.bench_kernel(3, n, 0, 5); out["potrf_%d" % n] = dict(ms=round(ms, 4), tflops=round(fl / ms / 1e9, 2))
print(json.dumps(out))

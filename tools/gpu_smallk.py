import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# rank-128 / rank-256 Cholesky trailing updates (NT, lower) with the 128x128 tile against the 64x64 tile (CVXB_SMALLK)
from cvx_b200 import _lib
h = _lib.default_handle()
pk_ms, pk_fl = h.bench_kernel(0, 20000, 0, 1)
pk = pk_fl / pk_ms / 1e9
print("dmma peak %.2f TFLOP/s, CVXB_SMALLK=%s" % (pk, os.environ.get("CVXB_SMALLK")))
for n, k in [(8064, 128), (4096, 128), (1920, 128), (8064, 256), (4096, 256), (8192, 16384)]:
    which = 2 if k <= 256 else 1
    ms, fl = h.bench_kernel(which, n, k, 5)
    print("which=%d n=%d k=%d: %.4f ms/launch, %.2f TFLOP/s (%.1f%% of peak)" % (which, n, k, ms, fl / ms / 1e9, 100 * fl / ms / 1e9 / pk))
for n in (2000, 4096):
    ms, fl = h.bench_kernel(3, n, 0, 5)
    print("potrf n=%d: %.3f ms" % (n, ms))

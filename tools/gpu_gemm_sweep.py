import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# SYRK (Hessian TN, which=1) and Cholesky trailing update (NT lower, which=2) rates at the BASELINE shapes
from cvx_b200 import _lib
h = _lib.default_handle()
pk_ms, pk_fl = h.bench_kernel(0, 20000, 0, 1)
print("dmma peak %.2f TFLOP/s" % (pk_fl / pk_ms / 1e9))
for which, n, k in [(1, 2000, 4000), (1, 2001, 5000), (1, 8192, 16384), (2, 2048, 128), (2, 4096, 2048), (2, 4096, 4096), (2, 8192, 128)]:
    ms, fl = h.bench_kernel(which, n, k, 5)
    print("which=%d n=%d k=%d: %.4f ms/launch, %.2f TFLOP/s (%.1f%% of peak)" % (which, n, k, ms, fl / ms / 1e9, 100 * fl / ms / 1e9 / (pk_fl / pk_ms / 1e9)))
for n in (2000, 8192):
    ms, by = h.bench_kernel(6, n, 0, 5)
    print("ruiz n=%d: %.4f ms per equilibration" % (n, ms))
for which, name in ((7, "gemv_n"), (8, "gemv_t")):
    for n, k in ((2000, 4000), (8192, 16384)):
        ms, by = h.bench_kernel(which, n, k, 10)
        print("%s G %dx%d: %.4f ms, %.0f GB/s" % (name, k, n, ms, by / ms / 1e6))

import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""First-contact GPU probe: kernel-level timings (DMMA peak, SYRK, Cholesky, HBM copy)."""
import json, sys
from cvx_b200 import _lib
h = _lib.default_handle()
out = {}
ms, fl = h.bench_kernel(0, 40000, 0, 1)
out["dmma_peak_tflops"] = fl / ms / 1e9
for n, k in [(2000, 4000), (8192, 16384)]:
    ms, fl = h.bench_kernel(1, n, k, 3)
    out["syrk_tn_%d_%d" % (n, k)] = dict(ms=ms, tflops=fl / ms / 1e9)
for n, k in [(8192, 512), (8192, 128), (2048, 1024)]:
    ms, fl = h.bench_kernel(2, n, k, 5)
    out["syrk_nt_%d_%d" % (n, k)] = dict(ms=ms, tflops=fl / ms / 1e9)
for n in [2000, 8192]:
    ms, fl = h.bench_kernel(3, n, 0, 3)
    out["potrf_%d" % n] = dict(ms=ms, tflops=fl / ms / 1e9)
ms, by = h.bench_kernel(4, 16384, 16384, 5)
out["copy_gbs"] = by / ms / 1e6
try:
    import torch, time
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record(); 
    for _ in range(3): c = a @ b
    e1.record(); torch.cuda.synchronize()
    out["cublas_dgemm_8192_tflops"] = 3 * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e9)
    e0.record(); L = torch.linalg.cholesky(a @ a.T + 8192 * torch.eye(8192, dtype=torch.float64, device="cuda")); e1.record(); torch.cuda.synchronize()
except Exception as ex:
    out["torch_err"] = repr(ex)
print(json.dumps(out, indent=1))

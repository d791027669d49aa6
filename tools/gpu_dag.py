import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Blocked Cholesky with / without the right-hand-side block riding along, tile-DAG schedule against the recursive one:
  python tools/gpu_dag.py [n r]...      (default: 8192 2049, 8192 0, 16385 1)"""
from cvx_b200 import _lib
h = _lib.default_handle()
pk_ms, pk_fl = h.bench_kernel(0, 20000, 0, 1)
print("dmma peak %.2f TFLOP/s" % (pk_fl / pk_ms / 1e9), flush=True)
args = [int(a) for a in sys.argv[1:]]
shapes = list(zip(args[0::2], args[1::2])) or [(8192, 2049), (8192, 0), (16385, 1)]
configs = [(0, 0)] + [(b, r) for b in (1024, 2048) for r in (4, 8, 12)]
for n, r in shapes:
    h.set_schedule()
    ms, fl = h.bench_kernel(9 if r > 0 else 3, n, r, 3)
    print("n=%d r=%d library defaults: %.3f ms, %.2f TFLOP/s" % (n, r, ms, fl / ms / 1e9), flush=True)
    for blk, res in configs:
        h.set_schedule(blk, 3 * 1024 if blk else -1, res if blk else -1)
        ms, fl = h.bench_kernel(9 if r > 0 else 3, n, r, 3)
        print("n=%d r=%d dag_block=%d reserve=%d: %.3f ms, %.2f TFLOP/s" % (n, r, blk, res, ms, fl / ms / 1e9), flush=True)

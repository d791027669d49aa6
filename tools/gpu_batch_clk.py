import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import cvx_b200 as cb
import synthetic as P
h = cb.default_handle()
lib = h.lib
lib.cvxb_debug_batch_clocks.argtypes = [C.POINTER(C.c_longlong), C.c_int]
buf = (C.c_longlong * 12)()
B = 600
s = cb.BatchedBarrierSolver(cb.pack_problems([P.batched_problem(i, 64, 128, 1000) for i in range(B)]))
s.solve()
lib.cvxb_debug_batch_clocks(buf, 1)
sol = s.solve()
lib.cvxb_debug_batch_clocks(buf, 1)
names = ["(loop/eval tail)", "hessian", "store H", "load H", "ruiz", "scale", "potrf", "solve+resid", "Gd + line search", "update + eval"]
tot = sum(buf[:10])
for n, v in zip(names, buf[:10]):
    print("  %-18s %12d cycles %5.1f%%" % (n, v, 100.0 * v / tot))
print("CTA 0 total %.2f ms of kernel %.2f ms" % (tot / 1.965e6, sol.solve_ms))

import os, sys; sys.path.insert(0, "/root/repo")
os.environ["CVXB_POTRF_CLOCKS"]="1"
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
print("potrf sub-phases: outside", buf[0], "diag", buf[1], "panel", buf[2], "update", buf[3], "blockstart", buf[4])

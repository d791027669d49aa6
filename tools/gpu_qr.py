"""Time the device SolutionSpace (Householder QR of A', explicit Q, z0) against LAPACK dgeqrf+dorgqr on the host cores.
usage: python tools/gpu_qr.py [p n]..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.linalg as sla
import cvx_b200 as cb

h = cb.default_handle()
sizes = [(500, 2000), (2048, 8192)]
if len(sys.argv) > 2:
    sizes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
for p, n in sizes:
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.uniform(-1, 1, (p, n)))
    b = rng.uniform(-1, 1, p)
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        sol = cb.SolutionSpace(A, b, h)
        dt = time.perf_counter() - t
        best = min(best, dt)
        if rep < 2:
            sol.close()
    z0 = sol.z0
    F = sol.F
    t = time.perf_counter()
    Q, R = sla.qr(A.T, mode="full")
    t_cpu = time.perf_counter() - t
    print("p=%d n=%d: device SolutionSpace %.1f ms (incl. upload of A), host LAPACK qr(full) %.1f ms on %d cores;  "
          "||A F||=%.2e ||F'F-I||=%.2e ||A z0-b||=%.2e  max|F-F_lapack|=%.2e"
          % (p, n, best * 1e3, t_cpu * 1e3, os.cpu_count(), np.linalg.norm(A @ F), np.linalg.norm(F.T @ F - np.eye(n - p)),
             np.linalg.norm(A @ z0 - b), np.abs(F - Q[:, p:]).max()), flush=True)

import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Randomised differential test of seam B (KKTSystem.solve with its whole fallback chain, choleskySolve,
SymmetricLinearSystem) against the CPU oracle: random sizes, condition numbers, bad row scalings, rank-deficient H
(path 1), indefinite H (path 2), dependent equality rows, inconsistent systems.
usage: python tools/gpu_fuzz_kkt.py [cases] [seed]"""
import time
import numpy as np
import cvx_b200 as cb
from cvx_b200 import KKTSystem, MatrixUtils, _lib
from oracle import cvx_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = cb.default_handle()
rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
bad = paths_differ = 0
tally = {}
t0 = time.time()
for it in range(N):
    n = int(rng.integers(2, 400))
    p = int(rng.integers(0, max(1, min(n - 1, 120))))
    kind = str(rng.choice(["pd", "pd_illcond", "pd_scaled", "semidef", "indef", "dep_rows", "inconsistent"]))
    Qm, _ = np.linalg.qr(rng.normal(size=(n, n)))
    if kind in ("pd", "dep_rows", "inconsistent"):
        lam = rng.uniform(0.5, 5.0, n)
    elif kind == "pd_illcond":
        lam = 10.0 ** rng.uniform(-float(rng.integers(2, 10)), 0.0, n)
    elif kind == "pd_scaled":
        lam = rng.uniform(0.5, 5.0, n)
    elif kind == "semidef":
        lam = rng.uniform(0.5, 5.0, n)
        lam[: max(1, min(p, n // 3))] = 0.0
    else:
        lam = rng.uniform(0.5, 5.0, n) * rng.choice([-1.0, 1.0], n)
    H = (Qm * lam) @ Qm.T
    if kind == "pd_scaled":
        s = 10.0 ** rng.uniform(-3, 3, n)
        H = H * np.outer(s, s)
    H = (H + H.T) * 0.5
    A = rng.uniform(-1, 1, (p, n))
    if kind == "semidef" and p:
        A[: min(p, n // 3)] = Qm[:, : min(p, n // 3)].T          # the null directions of H are in the row space of A
    if kind in ("dep_rows", "inconsistent") and p >= 2:
        A[-1] = A[0]
    x, w = rng.uniform(-1, 1, n), rng.uniform(-1, 1, p)
    q, b = -(H @ x + A.T @ w), A @ x
    if kind == "inconsistent" and p >= 2:
        b[-1] += 1.0
    tol = float(10.0 ** rng.uniform(-9, -3))
    try:
        info0 = O.KKTInfo()
        if p:
            x0, w0 = O.kkt_solve(H, A, q, b, tol, info0)
        else:
            x0, w0 = O.choleskySolve(H, -q, tol), np.zeros(0)
        r0 = ("ok", info0.path if p else 0)
    except Exception as e:
        r0 = (type(e).__name__, None)
    try:
        if p:
            K = KKTSystem(H, A, q, b, h)
            x1, w1 = K.solve(1e-6, None, tol, 0)
            path1 = K.info.path
        else:
            x1, w1, path1 = MatrixUtils.choleskySolve(H, -q, None, tol, 0, h), np.zeros(0), 0
        r1 = ("ok", path1)
    except cb.CvxbError as e:
        r1 = (type(e).__name__, None)
    key = (kind, r0[0] if r0[0] == "ok" else "fail", r1[0] if r1[0] == "ok" else "fail")
    tally[key] = tally.get(key, 0) + 1
    if (r0[0] == "ok") != (r1[0] == "ok"):
        bad += 1
        print("OUTCOME", it, kind, n, p, "tol %.1e" % tol, r0, r1, flush=True)
    elif r0[0] == "ok":
        def backward(xx, ww):
            res = np.concatenate([H @ xx + (A.T @ ww if p else 0.0) + q, (A @ xx - b) if p else np.zeros(0)])
            return np.linalg.norm(res) / max(np.linalg.norm(np.concatenate([q, b])), 1e-300)
        e0, e1 = backward(x0, w0), backward(x1, w1)
        if r0[1] != r1[1]:
            # a singular Schur complement (dependent rows) or a singular H: whether dpotrf gets through on a pivot of pure
            # rounding noise, and with it which path of the chain answers, is decided by rounding; both answers are accepted
            # by the reference's own test (error <= tol)
            paths_differ += 1
        if e1 > max(1e-10, 20 * e0) and e1 > tol:
            bad += 1
            print("MISMATCH", it, kind, n, p, "tol %.1e" % tol, "paths", r0[1], r1[1], "residuals %.2e %.2e" % (e0, e1), flush=True)
for k in sorted(tally):
    print("  %-13s oracle %-5s device %-5s %4d" % (k[0], k[1], k[2], tally[k]))
print("seam-B fuzz: %d cases, %d disagreements (%d more solved on a different path of the fallback chain, both within tol), %.1f s"
      % (N, bad, paths_differ, time.time() - t0))

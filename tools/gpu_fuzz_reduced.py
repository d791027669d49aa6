import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Randomised differential test of the equality elimination (SolutionSpace by device QR, BarrierSolver.reduced /
PrimalDualSolver.reduced for all objective families) against the oracle (LAPACK QR + the oracle's solve of the transformed
problem), and of the reduced solve against the equality-constrained solve of the same problem.
usage: python tools/gpu_fuzz_reduced.py [cases] [seed]"""
import time
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P

N = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = cb.default_handle()
rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
bad = 0
tally = {}
t0 = time.time()
for it in range(N):
    fam = str(rng.choice(["space", "slab_lp", "slab_qp", "kl", "pnorm", "quad"]))
    solver = str(rng.choice(["BR", "PD"]))
    seed = int(rng.integers(0, 10**6))
    try:
        if fam == "space":
            n = int(rng.integers(2, 300))
            p = int(rng.integers(1, n))
            if rng.integers(0, 3) == 0:
                p = min(n - 1, max(1, 32 * int(rng.integers(1, 6)) + int(rng.integers(-1, 2))))      # around the QR panel width
            r2 = np.random.default_rng(seed)
            A = r2.uniform(-1, 1, (p, n)) * 10.0 ** r2.uniform(-2, 2, (p, 1))
            b = r2.uniform(-1, 1, p)
            z0, F = cb.MatrixUtils.solveUnderdetermined(A, b, h)
            zr, Fr = O.solveUnderdetermined(A, b)
            k = n - p
            e = [np.linalg.norm(A @ F) / (np.linalg.norm(A) * np.sqrt(k)), np.linalg.norm(F.T @ F - np.eye(k)) / np.sqrt(k),
                 np.linalg.norm(A @ z0 - b) / max(1.0, np.linalg.norm(b)), rel(z0, zr), np.linalg.norm(F - Fr) / np.sqrt(k)]
            ok = e[0] < 1e-12 and e[1] < 1e-12 and e[2] < 1e-10 and e[3] < 1e-8 and e[4] < 1e-8
            tally[(fam, "ok" if ok else "BAD")] = tally.get((fam, "ok" if ok else "BAD"), 0) + 1
            if not ok:
                bad += 1
                print("SPACE", it, p, n, seed, ["%.1e" % v for v in e], flush=True)
            continue
        n = int(rng.integers(6, 70))
        p = int(rng.integers(1, max(2, n // 3)))
        if fam == "slab_lp":
            prob = P.slab_lp(n, int(rng.integers(n, 2 * n)), p, seed)
        elif fam == "slab_qp":
            prob = P.slab_qp(n, int(rng.integers(n // 2 + 1, 2 * n)), p, seed)
        elif fam == "kl":
            prob = P.kl_random(n, int(rng.integers(2, n + 1)), p, seed)
            prob["x0"] = prob["qstar"].copy()
        elif fam == "pnorm":
            r2 = np.random.default_rng(seed)
            x0 = r2.uniform(0.5, 1.5, n)
            A = r2.uniform(-1, 1, (p, n))
            prob = dict(kind="pnorm", n=n, a=None, r=0.0, P=None, pow=float(rng.choice([2.0, 2.5, 3.0, 4.0])),
                        G=np.vstack([np.eye(n), -np.eye(n)]), rvec=np.zeros(2 * n), ub=np.concatenate([x0 + 2.0, -x0 + 2.0]),
                        A=A, b=A @ x0, x0=x0, xdef=x0.copy())
        else:
            prob = P.lin_quad_set(n, int(rng.integers(0, n)), int(rng.integers(1, 4)), p, seed, "quadratic", True)
        objF, cnts, eqs = P.to_oracle(prob)

        def outcome(f, value):
            try:
                s_ = f()
                return ("ok", value(s_), s_)
            except (cb.CvxbError, O.InfeasibleProblemException, O.NotStrictlyFeasible, O.NotConvergedException,
                    O.UnsolvableSystemException, O.LineSearchFailedException) as e_:
                return (type(e_).__name__, None, None)
        # every solve against its oracle counterpart (the reference's primal-dual solver with equalities fails its residual
        # line search on most LPs and runs into maxIter on some quadratic problems -- so does the device)
        full0 = outcome(lambda: O.solveProblem(objF, cnts, eqs, solver)[0], lambda s_: objF.valueAt(s_.x))
        full1 = outcome(lambda: cb.from_dict(prob, solver, None, h).solve(), lambda s_: objF.valueAt(s_.x))
        z0, F = O.solveUnderdetermined(prob["A"], prob["b"])
        o_u, c_u = O.affineTransformedProblem(objF, cnts, z0, F)
        red0 = outcome(lambda: O.solveProblem(o_u, c_u, None, solver)[0], lambda s_: o_u.valueAt(s_.x))
        noeq = dict(prob)
        noeq["A"] = noeq["b"] = None
        red = cb.from_dict(noeq, solver, None, h).solver.reduced(cb.SolutionSpace(prob["A"], prob["b"], h))
        red1 = outcome(lambda: red.solve(), lambda s_: s_.objective)
        msgs = []
        for tag, r0, r1 in (("equality-constrained", full0, full1), ("reduced", red0, red1)):
            if (r0[0] == "ok") != (r1[0] == "ok") or (r0[0] != "ok" and r0[0] != r1[0]):
                msgs.append("%s: oracle %s device %s" % (tag, r0[0], r1[0]))
            elif r0[0] == "ok":
                if abs(r0[1] - r1[1]) > 1e-7 * max(1.0, abs(r0[1])): msgs.append("%s objective %.3e" % (tag, r1[1] - r0[1]))
                if solver == "BR" and r0[2].outer_stages != r1[2].outer_stages: msgs.append("%s stages" % tag)
                if solver == "PD" and abs(r0[2].newton_steps - r1[2].newton_steps) > 1: msgs.append("%s pd steps" % tag)
        if red1[0] == "ok":
            x = red.point(red1[2].x)
            if np.linalg.norm(prob["A"] @ x - prob["b"]) > 1e-9 * max(1.0, np.linalg.norm(prob["b"])): msgs.append("Ax=b")
            if not cnts.isSatisfiedStrictlyBy(x): msgs.append("infeasible")
        key = (fam + " " + solver, ("ok" if full1[0] == "ok" else "exc") + "/" + ("ok" if red1[0] == "ok" else "exc") if not msgs else "BAD")
        tally[key] = tally.get(key, 0) + 1
        if msgs:
            bad += 1
            print("REDUCED", it, fam, solver, n, p, seed, msgs, flush=True)
    except Exception as e:
        bad += 1
        tally[(fam + " " + solver, "EXC")] = tally.get((fam + " " + solver, "EXC"), 0) + 1
        print("EXC", it, fam, solver, seed, type(e).__name__, str(e)[:140], flush=True)
for k in sorted(tally):
    print("  %-14s %-8s %4d      (equality-constrained / reduced solve: ok or the reference's exception, device = oracle)" % (k[0], k[1], tally[k]))
print("reduction fuzz: %d cases, %d disagreements, %.1f s" % (N, bad, time.time() - t0))

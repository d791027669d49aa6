import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Randomised differential test of the dual route (Duality.solveDual on random Dist_KL problems, barrier and primal-dual
solver) and of the feasibility analyses (phase_I_Analysis, phase_I_Analysis_SOI, withFeasiblePoint) against the oracle.
usage: python tools/gpu_fuzz_dual_feas.py [cases] [seed]"""
import time
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P
from tests.test_feasibility_gpu import mirror_set

N = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = cb.default_handle()
rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
bad = 0
tally = {}
t0 = time.time()
EXC = (cb.CvxbError, O.InfeasibleProblemException, O.NotStrictlyFeasible, O.NotConvergedException, O.UnsolvableSystemException,
       O.LineSearchFailedException)
for it in range(N):
    fam = str(rng.choice(["dual_BR", "dual_PD", "phase1", "soi"]))
    seed = int(rng.integers(0, 10**6))
    msgs = []
    try:
        if fam.startswith("dual"):
            n = int(rng.integers(6, 150))
            mh = int(rng.integers(0, n))
            pe = int(rng.integers(0, max(1, n // 4)))
            pr = P.kl_random(n, max(mh, 1), pe, seed)
            H, u = (pr["G"][:mh], pr["ub"][:mh]) if mh else (None, None)
            A, r = (pr["A"][:pe], pr["b"][:pe]) if pe else (None, None)
            prob = cb.Dist_KL(n, H, u, A, r, "BR", None, None, 0, h)
            st = fam[-2:]
            try:
                if st == "BR":
                    s0 = O.solveDual(n, H, u, A, r); x0 = s0.x
                else:
                    objF, cnts, mI = O.dist_KL_dual_problem(n, H, u, A, r)
                    s0 = O.PrimalDual(objF, cnts, None, O.SolverParams.standardParams()).solve(); x0 = objF.primalOptimum(s0.x)
                r0 = "ok"
            except EXC as e:
                r0 = type(e).__name__
            try:
                s1 = prob.solveDual(st); r1 = "ok"
            except EXC as e:
                r1 = type(e).__name__
            if (r0 == "ok") != (r1 == "ok"): msgs.append("oracle %s device %s" % (r0, r1))
            elif r0 == "ok":
                if rel(s1.x, x0) > 1e-6: msgs.append("x %.2e" % rel(s1.x, x0))
                if st == "BR" and s1.outer_stages != s0.outer_stages: msgs.append("stages")
                if st == "PD" and abs(s1.newton_steps - s0.newton_steps) > 1: msgs.append("pd steps %d %d" % (s1.newton_steps, s0.newton_steps))
            key = (fam, r1 if r1 == "ok" else "exc")
        else:
            n = int(rng.integers(3, 60))
            kind = str(rng.choice(["lp", "kl", "quad", "infeasible"]))
            if kind == "lp": pr = P.slab_lp(n, int(rng.integers(n, 2 * n)), int(rng.integers(0, max(1, n // 4))), seed, feasible_start=False)
            elif kind == "kl": pr = P.kl_random(max(n, 4), int(rng.integers(1, n + 1)), int(rng.integers(0, max(1, n // 4))), seed)
            elif kind == "quad": pr = P.lin_quad_set(n, int(rng.integers(0, n)), int(rng.integers(1, 4)), int(rng.integers(0, max(1, n // 4))), seed, "quadratic", False)
            else: pr = P.infeasible_kl_1(max(n, 10) + (max(n, 10) % 2))
            _, cnts0, eqs0 = P.to_oracle(pr)
            cnts, eqs = mirror_set(cb, pr)
            if fam == "soi":
                try:
                    rep0, sol0 = O.phase_I_Analysis_SOI(cnts0, eqs0, O.SolverParams()); r0 = "ok"
                except EXC as e:
                    r0 = type(e).__name__
                try:
                    rep = cnts.phase_I_Analysis_SOI(eqs, None, 0, h); r1 = "ok"
                except EXC as e:
                    r1 = type(e).__name__
                if (r0 == "ok") != (r1 == "ok"): msgs.append("oracle %s device %s" % (r0, r1))
                elif r0 == "ok":
                    if rep.isFeasible(1e-9) != rep0.isFeasible(1e-9): msgs.append("isFeasible")
                    if abs(rep.s.sum() - rep0.s.sum()) > 1e-7 * max(1.0, rep0.s.sum()): msgs.append("sum s %.3e" % (rep.s.sum() - rep0.s.sum()))
                    if rep.solution.outer_stages != sol0.outer_stages: msgs.append("stages")
            else:
                try:
                    x0, s0, sol0 = O.phase_I_Analysis(cnts0, eqs0, O.SolverParams()); r0 = "ok"
                except EXC as e:
                    r0 = type(e).__name__
                try:
                    rep = cnts.phase_I_Analysis(eqs, None, 0, h); r1 = "ok"
                except EXC as e:
                    r1 = type(e).__name__
                if (r0 == "ok") != (r1 == "ok"): msgs.append("oracle %s device %s" % (r0, r1))
                elif r0 == "ok":
                    if (rep.s[0] < 0) != (s0 < 0): msgs.append("sign of s")
                    if rep.solution.outer_stages != sol0.outer_stages: msgs.append("stages %d %d" % (rep.solution.outer_stages, sol0.outer_stages))
                    if rep.isFeasible(1e-9) != (s0 < 1e-9): msgs.append("isFeasible")
            key = (fam + " " + kind, r1 if r1 == "ok" else "exc")
        if msgs:
            key = (key[0], "BAD")
            bad += 1
            print("CASE", it, fam, seed, msgs, flush=True)
        tally[key] = tally.get(key, 0) + 1
    except Exception as e:
        bad += 1
        print("HARNESS", it, fam, seed, type(e).__name__, str(e)[:160], flush=True)
for k in sorted(tally):
    print("  %-18s %-4s %4d" % (k[0], k[1], tally[k]))
print("dual / feasibility fuzz: %d cases, %d disagreements, %.1f s" % (N, bad, time.time() - t0))

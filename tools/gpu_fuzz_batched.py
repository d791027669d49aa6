import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Randomised differential test of the batched one-CTA-per-problem kernel against the CPU oracle: random shapes
(n <= 63, m <= 126), KL / QP / LP problems with and without an equality, with a feasible start or only a point where the
problem is defined (phase I inside the CTA), convex and a few nonconvex QPs (decomposition last resort).
usage: python tools/gpu_fuzz_batched.py [batches] [seed]"""
import time
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = cb.default_handle()
# Ill-posed inputs (unbounded LPs with fewer slab rows than variables, empty feasible sets, nonconvex objectives) end in
# one of the reference's exceptions; WHICH one the iteration runs into first at the edge of the barrier's domain is decided
# by rounding (the one-problem device path differs from the oracle there in the same way), so they form one class.
STATUS = {O.InfeasibleProblemException: "failed", O.NotStrictlyFeasible: "failed", O.NotConvergedException: "failed",
          O.UnsolvableSystemException: "failed"}
bad = cases = 0
tally = {}
t0 = time.time()
for b in range(NB):
    n = int(rng.integers(4, 64))
    mh = int(rng.integers(max(2, n // 2), n + 1))          # KL: mh rows + n positivity rows; QP/LP: 2 * mh rows
    m = n + mh
    if m % 2:
        mh += 1
        m += 1
    if m > 126:
        continue
    probs = []
    for i in range(12):
        seed = int(rng.integers(0, 10**6))
        fam = rng.choice(["kl", "kl_phase1", "qp", "qp_phase1", "qp_eq", "lp", "nonconvex"])
        if fam == "kl":
            pr = P.kl_small(n, m - n, seed)
        elif fam == "kl_phase1":
            pr = P.kl_random(n, m - n, 0, seed)
        elif fam in ("qp", "qp_phase1", "qp_eq", "nonconvex"):
            pr = P.slab_qp(n, m // 2, 1 if fam in ("qp_eq", "nonconvex") else 0, seed, scale=True)
            if fam == "qp_phase1":
                pr["xdef"] = pr["x0"] + rng.uniform(0.3, 1.0, n)
                pr["x0"] = None
            if fam == "nonconvex":
                pr["P"] = pr["P"] - float(rng.choice([2.0, 10.0, 50.0])) * np.eye(n)
        else:
            pr = P.slab_lp(n, m // 2, 0, seed)
        if pr["G"].shape[0] != m:
            continue
        probs.append((fam, seed, pr))
    if not probs:
        continue
    sol = cb.BatchedBarrierSolver(cb.pack_problems([p_[2] for p_ in probs]), None, h).solve()
    for i, (fam, seed, pr) in enumerate(probs):
        cases += 1
        objF, cnts, eqs = P.to_oracle(pr)
        try:
            s0, _ = O.solveProblem(objF, cnts, eqs, "BR")
            want = ("ok", objF.valueAt(s0.x), s0.outer_stages)
        except tuple(STATUS) as e:
            want = (STATUS[type(e)], None, None)
        st = int(sol.status[i])
        got = "ok" if st == 0 else ("failed" if st in (cb._lib.EUNSOLVABLE, cb._lib.EINFEASIBLE, cb._lib.ENOTFEASIBLE,
                                                         cb._lib.ELINESEARCH) else "status %d" % st)
        tally[(fam, want[0], got)] = tally.get((fam, want[0], got), 0) + 1
        if want[0] != got:
            bad += 1
            print("OUTCOME", b, i, fam, n, m, seed, want[0], got, flush=True)
        elif got == "ok":
            if abs(sol.objective[i] - want[1]) > 1e-7 * max(1.0, abs(want[1])) or sol.outer_stages[i] != want[2]:
                bad += 1
                print("MISMATCH", b, i, fam, n, m, seed, want, sol.objective[i], sol.outer_stages[i], flush=True)
for k in sorted(tally):
    print("  %-10s oracle %-7s device %-9s %4d" % (k[0], k[1], k[2], tally[k]))
print("batched fuzz: %d cases, %d disagreements, %.1f s" % (cases, bad, time.time() - t0))

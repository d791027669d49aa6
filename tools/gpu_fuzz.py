import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Randomised differential test: device path vs CPU oracle on many small random problems of every family."""
import time
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
NMIN = int(sys.argv[3]) if len(sys.argv) > 3 else 2          # python tools/gpu_fuzz.py cases seed [nmin nmax]: medium sizes
NMAX = int(sys.argv[4]) if len(sys.argv) > 4 else 45         # exercise the multi-leaf Cholesky schedules and the wavefront solves
h = cb.default_handle()
bad = explained = 0
t0 = time.time()
for it in range(N):
    fam = rng.choice(["slab_lp", "slab_qp", "kl", "quad", "pnorm", "lp_phase1", "kl_phase1"])
    n = int(rng.integers(NMIN, NMAX))
    seed = int(rng.integers(0, 10**6))
    solver = str(rng.choice(["BR", "PD"]))
    try:
        if fam == "slab_lp":
            prob = P.slab_lp(n, int(rng.integers(n, 2 * n + 2)), int(rng.integers(0, max(1, min(6, n - 1)))), seed)
        elif fam == "slab_qp":
            prob = P.slab_qp(n, int(rng.integers(n // 2 + 1, 2 * n + 2)), int(rng.integers(0, max(1, min(6, n - 1)))), seed)
        elif fam == "kl":
            prob = P.kl_small(n, int(rng.integers(1, n + 2)), seed)
        elif fam == "quad":
            prob = P.lin_quad_set(n, int(rng.integers(0, n + 3)), int(rng.integers(1, 5)), int(rng.integers(0, max(1, min(4, n - 1)))), seed,
                                  str(rng.choice(["quadratic", "linear"])) if rng.integers(0, 2) else "quadratic", bool(rng.integers(0, 2)))
            if prob["kind"] == "linear":
                # bounded: add a box so the LP has a finite optimum
                prob["G"] = np.vstack([prob["G"], np.eye(n), -np.eye(n)])
                prob["ub"] = np.concatenate([prob["ub"], prob["xdef"] + 5.0 if prob["x0"] is None else prob["x0"] + 5.0,
                                             -(prob["xdef"] if prob["x0"] is None else prob["x0"]) + 5.0])
                prob["rvec"] = np.zeros(prob["G"].shape[0])
        elif fam == "pnorm":
            prob = P.min_pNorm(max(n, 2), float(rng.choice([2.0, 2.5, 3.0, 4.0])))
        elif fam == "lp_phase1":
            prob = P.slab_lp(n, int(rng.integers(n, 2 * n + 2)), 0, seed, feasible_start=False)
        else:
            prob = P.kl_random(max(n, 4), int(rng.integers(1, n + 2)), int(rng.integers(0, max(1, min(5, n - 2)))), seed)
        objF, cnts, eqs = P.to_oracle(prob)
        try:
            sol0, _ = O.solveProblem(objF, cnts, eqs, solver)
            r0 = ("ok", objF.valueAt(sol0.x))
        except Exception as e:
            r0 = (type(e).__name__, None)
        try:
            sol = cb.from_dict(prob, solver, None, h).solve()
            r1 = ("ok", sol.objective)
        except cb.CvxbError as e:
            r1 = (type(e).__name__, None)
        if r0[0] == "ok" and r1[0] == "ok":
            if abs(r0[1] - r1[1]) > 1e-7 * max(1.0, abs(r0[1])):
                bad += 1
                print("MISMATCH", it, fam, n, seed, solver, r0, r1, flush=True)
        elif (r0[0] == "ok") != (r1[0] == "ok"):
            # is it the one as-if deviation with numerical consequences?  The device forms the Schur complement as Y'Y
            # (positive semidefinite by construction), the reference as A (H^-1 A') symmetrised
            O.BLOCK_ELIMINATION = "one_trsm"
            try:
                sol2, _ = O.solveProblem(objF, cnts, eqs, solver)
                r2 = ("ok", objF.valueAt(sol2.x))
            except Exception as e:
                r2 = (type(e).__name__, None)
            finally:
                O.BLOCK_ELIMINATION = "reference"
            if (r2[0] == r1[0] == "ok" and abs(r2[1] - r1[1]) <= 1e-7 * max(1.0, abs(r2[1]))) or (r2[0] == r1[0] != "ok"):
                explained += 1
                print("EXPLAINED", it, fam, n, seed, solver, r0, r1, "oracle with the device's block elimination:", r2, flush=True)
            else:
                bad += 1
                print("OUTCOME", it, fam, n, seed, solver, r0, r1, "with the device's block elimination:", r2, flush=True)
    except Exception as e:
        bad += 1
        print("HARNESS", it, fam, n, seed, solver, repr(e), flush=True)
print("fuzz: %d cases, %d disagreements, %d more where the outcome depends on the formulation of the block elimination (reference: "
      "A (H^-1 A') symmetrised; device: Y'Y) -- the oracle agrees with the device once it uses the device's formulation, %.1f s"
      % (N, bad, explained, time.time() - t0))

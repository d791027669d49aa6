import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Replays tools/gpu_fuzz_reduced.py's generator to given iterations and reports each solve separately."""
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P
fseed = int(sys.argv[1]); wanted = [int(a) for a in sys.argv[2:]]
rng = np.random.default_rng(fseed)
h = cb.default_handle()
for it in range(max(wanted) + 1):
    fam = str(rng.choice(["space", "slab_lp", "slab_qp", "kl", "pnorm", "quad"]))
    solver = str(rng.choice(["BR", "PD"]))
    seed = int(rng.integers(0, 10**6))
    prob = None
    if fam == "space":
        n = int(rng.integers(2, 300)); p = int(rng.integers(1, n))
        if rng.integers(0, 3) == 0:
            p = min(n - 1, max(1, 32 * int(rng.integers(1, 6)) + int(rng.integers(-1, 2))))
        continue
    n = int(rng.integers(6, 70)); p = int(rng.integers(1, max(2, n // 3)))
    if fam == "slab_lp": prob = P.slab_lp(n, int(rng.integers(n, 2 * n)), p, seed)
    elif fam == "slab_qp": prob = P.slab_qp(n, int(rng.integers(n // 2 + 1, 2 * n)), p, seed)
    elif fam == "kl":
        prob = P.kl_random(n, int(rng.integers(2, n + 1)), p, seed); prob["x0"] = prob["qstar"].copy()
    elif fam == "pnorm":
        pw = float(rng.choice([2.0, 2.5, 3.0, 4.0]))
    else:
        prob = P.lin_quad_set(n, int(rng.integers(0, n)), int(rng.integers(1, 4)), p, seed, "quadratic", True)
    if it not in wanted or prob is None:
        continue
    print("== it", it, fam, solver, "n", n, "p", p, "m", prob["G"].shape[0], "mq", len(prob.get("quad") or []), "seed", seed, flush=True)
    objF, cnts, eqs = P.to_oracle(prob)
    def run(tag, f):
        try:
            s = f()
            print("  %-28s ok objective %.12g steps %s" % (tag, objF.valueAt(s.x) if len(s.x) == n else float("nan"), s.newton_steps), flush=True)
            return s
        except Exception as e:
            print("  %-28s %s %s" % (tag, type(e).__name__, str(e)[:100]), flush=True)
            return None
    run("oracle equality-constrained", lambda: O.solveProblem(objF, cnts, eqs, solver)[0])
    full = run("device equality-constrained", lambda: cb.from_dict(prob, solver, None, h).solve())
    z0, F = O.solveUnderdetermined(prob["A"], prob["b"])
    o_u, c_u = O.affineTransformedProblem(objF, cnts, z0, F)
    try:
        s0, _ = O.solveProblem(o_u, c_u, None, solver)
        print("  oracle reduced               ok objective %.12g steps %d" % (o_u.valueAt(s0.x), s0.newton_steps))
    except Exception as e:
        print("  oracle reduced              ", type(e).__name__, str(e)[:100])
    try:
        noeq = dict(prob); noeq["A"] = noeq["b"] = None
        red = cb.from_dict(noeq, solver, None, h).solver.reduced(cb.SolutionSpace(prob["A"], prob["b"], h))
        su = red.solve()
        print("  device reduced               ok objective %.12g steps %d" % (su.objective, su.newton_steps))
    except Exception as e:
        print("  device reduced              ", type(e).__name__, str(e)[:100])

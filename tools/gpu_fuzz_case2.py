import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Replays tools/gpu_fuzz.py's generator to given iterations and prints what the oracle and the device did there.
usage: python tools/gpu_fuzz_case2.py <fuzz seed> <iteration>..."""
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P
from collections import Counter
fseed = int(sys.argv[1]); wanted = [int(a) for a in sys.argv[2:]]
rng = np.random.default_rng(fseed)
h = cb.default_handle()
for it in range(max(wanted) + 1):
    fam = rng.choice(["slab_lp", "slab_qp", "kl", "quad", "pnorm", "lp_phase1", "kl_phase1"])
    n = int(rng.integers(2, 45)); seed = int(rng.integers(0, 10**6)); solver = str(rng.choice(["BR", "PD"]))
    prob = None
    if fam == "slab_lp":
        a = (int(rng.integers(n, 2 * n + 2)), int(rng.integers(0, max(1, min(6, n - 1)))))
        prob = P.slab_lp(n, a[0], a[1], seed)
    elif fam == "slab_qp": a = (int(rng.integers(n // 2 + 1, 2 * n + 2)), int(rng.integers(0, max(1, min(6, n - 1)))))
    elif fam == "kl": a = (int(rng.integers(1, n + 2)),)
    elif fam == "quad":
        a = (int(rng.integers(0, n + 3)), int(rng.integers(1, 5)), int(rng.integers(0, max(1, min(4, n - 1)))))
        obj = str(rng.choice(["quadratic", "linear"])) if rng.integers(0, 2) else "quadratic"
        feas = bool(rng.integers(0, 2))
    elif fam == "pnorm": a = (float(rng.choice([2.0, 2.5, 3.0, 4.0])),)
    elif fam == "lp_phase1": a = (int(rng.integers(n, 2 * n + 2)),)
    else: a = (int(rng.integers(1, n + 2)), int(rng.integers(0, max(1, min(5, n - 2)))))
    if it not in wanted or prob is None:
        continue
    print("== it", it, fam, "n", n, "m_half, p", a, "seed", seed, solver, flush=True)
    objF, cnts, eqs = P.to_oracle(prob)
    stats = []
    try:
        if solver == "BR":
            s0 = O.barrierSolve(objF, cnts, eqs, O.SolverParams.standardParams(), None, False, kkt_stats=stats)
        else:
            s0 = O.PrimalDual(objF, cnts, eqs, O.SolverParams.standardParams(), False, False).solve()
        print("  oracle ok", objF.valueAt(s0.x), s0.outer_stages, s0.stage_newton_steps)
    except Exception as e:
        print("  oracle", type(e).__name__, str(e)[:100])
    print("  oracle kkt paths", Counter((s_.path, s_.regularized) for s_ in stats), "steps", len(stats))
    try:
        s1 = cb.from_dict(prob, solver, None, h).solve()
        print("  device ok", s1.objective, s1.outer_stages, s1.stage_newton_steps[:s1.outer_stages], "fallbacks", s1.kkt_fallbacks, "regularized", s1.kkt_regularized)
    except cb.CvxbError as e:
        print("  device", type(e).__name__, str(e)[:160])

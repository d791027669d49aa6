#!/usr/bin/env python
"""Are the OUTCOMES (solution vs exception) of the reference's solvers reproducible under rounding-level changes on the
problems where the device and the oracle disagree?  CPU oracle only, no GPU.

The randomised differential tests (tools/gpu_fuzz.py) disagree with the oracle on a few slab LPs in their last stages: the
oracle throws UnsolvableSystemException (defect D3 at the end of KKTSystem.solve's fallback chain) or fails / passes the
primal-dual residual line search where the device does the opposite.  This script re-runs exactly those problems on the CPU
oracle with mathematically neutral changes (tools/iteration_noise_experiment.py's variants: OpenBLAS on one thread, the
Hessian accumulated constraint by constraint as the reference does, every entry of G and ub multiplied by
1 + k*2^-53) and records which variants end in a solution and which in which exception.

usage: python tools/outcome_noise_experiment.py [--out tests/golden/outcome_noise.json]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cvx_oracle as O      # noqa: E402
from oracle import problems as P        # noqa: E402
from tools.iteration_noise_experiment import perturbed   # noqa: E402

# (name, generator arguments of synthetic.slab_lp(n, m_half, p, seed), solver): the disagreements of
# profiles/r2_fuzz_large_path.log and r2_fuzz_reduced.log
CASES = [("fuzz11_it18", (27, 54, 1, 20152), "BR"), ("fuzz11_it68", (34, 48, 1, 616934), "BR"),
         ("fuzz11_it135", (31, 39, 2, 83398), "BR"), ("fuzz11_it31", (30, 58, 4, 915843), "PD"),
         ("reduced4_it54", (46, None, 3, 676326), "PD")]


def run(prob, solver, literal=False):
    objF, cnts, eqs = P.to_oracle(prob)
    try:
        sol, _ = O.solveProblem(objF, cnts, eqs, solver, literal=literal)
        return {"outcome": "ok", "objective": float(objF.valueAt(sol.x)), "newton_steps": int(sol.newton_steps)}
    except Exception as e:          # the reference's exceptions
        return {"outcome": type(e).__name__}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "outcome_noise.json"))
    args = ap.parse_args()
    from threadpoolctl import threadpool_limits
    res = {}
    for name, (n, mh, p, seed), solver in CASES:
        if mh is None:
            continue            # (the generator arguments of that case are replayed by tools/gpu_fuzz_reduced_case.py)
        prob = P.slab_lp(n, mh, p, seed)
        v = {"base": run(prob, solver)}
        with threadpool_limits(limits=1):
            v["threads1"] = run(prob, solver)
        v["literal"] = run(prob, solver, literal=True)
        for s in range(8):
            v["ulp%d" % s] = run(perturbed(prob, 200 + s), solver)
        res[name] = {"problem": "slab_lp(n=%d, m_half=%d, p=%d, seed=%d), %s" % (n, mh, p, seed, solver), "variants": v}
        print(name, solver, {k: r["outcome"] for k, r in v.items()}, flush=True)
    flips = {k: sorted({r["outcome"] for r in v["variants"].values()}) for k, v in res.items()}
    summary = {"outcomes_seen_per_problem": flips, "problems_whose_outcome_flips": sum(len(f) > 1 for f in flips.values()),
               "problems": len(flips),
               "note": "outcomes of the CPU ORACLE ALONE under neutral changes (BLAS threads, literal Hessian accumulation, "
                       "half-ulp input changes) on the problems where device and oracle disagree"}
    print(json.dumps(summary, indent=1))
    json.dump({"summary": summary, "runs": res, "numpy": np.__version__}, open(args.out, "w"), indent=0)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""How reproducible are the reference's per-stage Newton iteration counts under rounding-level changes?

north_star asks for "the same iteration count +-1".  The inner loops of the reference stop on
`newtonDecrement > tol && normGrad > tol` (EqualityConstrainedSolver.scala:49, UnconstrainedSolver.scala:45) with
tol = 1e-8, evaluated on t*f0(x) - sum log d_i.  From t ~ 1e4 on the barrier value is ~1e4..1e12 and its rounding error
(1e-16 relative) is of the order of the decrement being tested, so WHICH iteration first passes the test depends on
rounding.  This script measures that on the CPU oracle alone -- no GPU involved -- by re-running the same solves with
changes that are mathematically neutral:
   threads1   OpenBLAS restricted to 1 thread (different dgemm / dgemv summation order)
   literal    the Hessian accumulated constraint by constraint as the reference does (BarrierSolver.scala:303-315)
              instead of one dgemm
   ulp        every entry of G and ub multiplied by (1 + k*2^-53), k in {-1, 0, 1} (a half-ulp input change)
and records, per problem and stage, the Newton counts of every variant.  The band the GPU tests allow is derived from
the largest deviation between two CPU variants (tests/golden/iteration_noise.json, read by tests/test_barrier_gpu.py).

usage: python tools/iteration_noise_experiment.py [--out tests/golden/iteration_noise.json]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cvx_oracle as O      # noqa: E402
from oracle import problems as P        # noqa: E402


def problems():
    out = {
        "slab_qp_64": P.slab_qp(64, 64, 0, 1),
        "kl_small_64": P.kl_small(64, 64, 2),
        "slab_qp_eq": P.slab_qp(48, 60, 6, 3),
        "min_dot_product": P.min_dot_product(np.linspace(0.5, 2, 10)),
        "kl_1A": P.kl_1A(20),
        "kl_2A": P.kl_2A(20),
        "kl_random_120": P.kl_random(120, 120, 29, 5),
        "slab_lp_phase1": P.slab_lp(40, 60, 0, 7, feasible_start=False),
        "slab_lp_100_eq20": P.slab_lp(100, 100, 20, 0),
    }
    for i in range(16):
        out["batched_%d" % i] = P.batched_problem(i, 64, 128, 1000)
    return out


def perturbed(prob, seed):
    rng = np.random.default_rng(seed)
    q = dict(prob)
    for k in ("G", "ub"):
        a = np.array(prob[k], dtype=np.float64)
        q[k] = a * (1.0 + rng.integers(-1, 2, a.shape) * 2.0 ** -53)
    return q


def run(prob, literal=False):
    objF, cnts, eqs = P.to_oracle(prob)
    sol, ph = O.solveProblem(objF, cnts, eqs, "BR", literal=literal)
    return {"stages": [int(s) for s in sol.stage_newton_steps], "objective": float(objF.valueAt(sol.x)),
            "phase1": None if ph is None else [int(s) for s in ph.stage_newton_steps]}


def summarise(res):
    """Largest deviation of the per-stage Newton counts between ANY TWO CPU variants of the same problem, per stage index,
    and the band derived from it for the GPU tests: stage k may differ from the oracle by at most
        band[k] = 1                                    while no CPU variant pair differs at any stage <= k  (north_star's +-1)
        band[k] = 2 * max(deviation at stages <= k)    afterwards (the GPU is one more rounding variant; factor 2 = margin
                                                       for the small sample)
    Stages in which a variant runs to maxIter (the ||b-Ax|| > tol spin of EqualityConstrainedSolver.scala:49) are
    counted separately: whether a stage spins is itself rounding-decided."""
    dev, spin, obj_rel, nst = {}, 0, 0.0, 0
    for name, v in res.items():
        runs = list(v.values())
        nst = max(nst, len(runs[0]["stages"]))
        for a_ in range(len(runs)):
            for b_ in range(a_ + 1, len(runs)):
                ra, rb = runs[a_], runs[b_]
                if len(ra["stages"]) != len(rb["stages"]):
                    raise SystemExit("%s: different number of outer stages between variants" % name)
                obj_rel = max(obj_rel, abs(ra["objective"] - rb["objective"]) / max(1.0, abs(rb["objective"])))
                for st, (a, b) in enumerate(zip(ra["stages"], rb["stages"])):
                    if a >= 1000 or b >= 1000:
                        spin += int(a != b)
                        continue
                    dev[st] = max(dev.get(st, 0), abs(a - b))
    by_stage = [dev.get(k, 0) for k in range(nst)]
    band, run_max = [], 0
    for k in range(nst):
        run_max = max(run_max, by_stage[k])
        band.append(1 if run_max == 0 else 2 * run_max)
    return {"pairwise_max_deviation_by_stage": by_stage, "band_by_stage": band, "spin_stage_flips": spin,
            "max_rel_objective_deviation": obj_rel, "problems": len(res), "variants_per_problem": len(next(iter(res.values()))),
            "note": "deviations of the per-stage Newton counts BETWEEN CPU ORACLE VARIANTS (BLAS threads, literal Hessian, "
                    "half-ulp input changes); stage k has barrier parameter t = 10^k; 'spin' stages run to maxIter because "
                    "||b-Ax|| stays above 1e-8 (EqualityConstrainedSolver.scala:49)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "iteration_noise.json"))
    ap.add_argument("--summarise-only", action="store_true", help="recompute the summary of an existing result file")
    args = ap.parse_args()
    if args.summarise_only:
        old = json.load(open(args.out))
        old["summary"] = summarise(old["runs"])
        print(json.dumps(old["summary"], indent=1))
        json.dump(old, open(args.out, "w"), indent=0)
        return
    from threadpoolctl import threadpool_limits
    res = {}
    for name, prob in problems().items():
        v = {"base": run(prob)}
        with threadpool_limits(limits=1):
            v["threads1"] = run(prob)
        v["literal"] = run(prob, literal=True)
        for s in range(3):
            v["ulp%d" % s] = run(perturbed(prob, 100 + s))
        res[name] = v
        print(name, {k: r["stages"] for k, r in v.items()}, flush=True)
    summary = summarise(res)
    print(json.dumps(summary, indent=1))
    json.dump({"summary": summary, "runs": res, "numpy": np.__version__}, open(args.out, "w"), indent=0)


if __name__ == "__main__":
    main()

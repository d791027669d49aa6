"""Problem dicts -> CPU-oracle objects, plus every generator of synthetic.py re-exported for the tests.

TEST INFRASTRUCTURE ONLY (like the rest of oracle/): only tests/, __graft_entry__.smoke() and the CPU legs of
bench.py import this module.  The generators themselves (pure numpy input generation, shared with bench.py's
GPU arm and tools/) live in synthetic.py at the repo root."""
from __future__ import annotations

import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from synthetic import *  # noqa: F401,F403,E402  (slab_lp, slab_qp, kl_random, batched_problem, kkt_planted_*, ...)
import numpy as np  # noqa: E402


# ---------------- conversion to oracle objects ----------------


def to_oracle(prob):
    from . import cvx_oracle as O
    if prob["kind"] == "linear":
        objF = O.LinearObjective(prob["a"], prob["r"])
    elif prob["kind"] == "quadratic":
        objF = O.QuadraticObjective(prob["P"], prob["a"], prob["r"])
    elif prob["kind"] == "kl":
        objF = O.KLObjective(prob["n"])
    elif prob["kind"] == "pnorm":
        objF = O.PNormObjective(prob["n"], prob["pow"])
    else:
        raise ValueError(prob["kind"])
    quad = [O.QuadCnt(q["P"], q["a"], q["r"], q["ub"]) for q in prob.get("quad") or []]
    cnts = O.ConstraintSet(prob["G"], prob["rvec"], prob["ub"], quad or None, prob["xdef"])
    if prob.get("x0") is not None:
        cnts = cnts.addFeasiblePoint(prob["x0"])
    eqs = O.EqualityConstraint(prob["A"], prob["b"]) if prob.get("A") is not None else None
    return objF, cnts, eqs

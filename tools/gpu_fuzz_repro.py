import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cvx_b200 as cb
from oracle import cvx_oracle as O, problems as P
# case: quad n=15 seed=446736 PD (fuzz seed 7, it 31): regenerate by replaying the generator choices
rng = np.random.default_rng(7)
for it in range(32):
    fam = rng.choice(["slab_lp", "slab_qp", "kl", "quad", "pnorm", "lp_phase1", "kl_phase1"])
    n = int(rng.integers(2, 45)); seed = int(rng.integers(0, 10**6)); solver = str(rng.choice(["BR", "PD"]))
    if fam == "slab_lp": a = (int(rng.integers(n, 2 * n + 2)), int(rng.integers(0, max(1, min(6, n - 1)))))
    elif fam == "slab_qp": a = (int(rng.integers(n // 2 + 1, 2 * n + 2)), int(rng.integers(0, max(1, min(6, n - 1)))))
    elif fam == "kl": a = (int(rng.integers(1, n + 2)),)
    elif fam == "quad":
        a = (int(rng.integers(0, n + 3)), int(rng.integers(1, 5)), int(rng.integers(0, max(1, min(4, n - 1)))))
        obj = str(rng.choice(["quadratic", "linear"])) if rng.integers(0, 2) else "quadratic"
        feas = bool(rng.integers(0, 2))
    elif fam == "pnorm": a = (float(rng.choice([2.0, 2.5, 3.0, 4.0])),)
    elif fam == "lp_phase1": a = (int(rng.integers(n, 2 * n + 2)),)
    else: a = (int(rng.integers(1, n + 2)), int(rng.integers(0, max(1, min(5, n - 2)))))
print(fam, n, seed, solver, a, obj, feas)
prob = P.lin_quad_set(n, a[0], a[1], a[2], seed, obj, feas)
if prob["kind"] == "linear":
    prob["G"] = np.vstack([prob["G"], np.eye(n), -np.eye(n)])
    base = prob["xdef"] if prob["x0"] is None else prob["x0"]
    prob["ub"] = np.concatenate([prob["ub"], base + 5.0, -base + 5.0]); prob["rvec"] = np.zeros(prob["G"].shape[0])
objF, cnts, eqs = P.to_oracle(prob)
sol0, ph0 = O.solveProblem(objF, cnts, eqs, solver)
print("oracle", objF.valueAt(sol0.x), sol0.newton_steps, None if ph0 is None else ph0.stage_newton_steps)
op = cb.from_dict(prob, solver, None, cb.default_handle())
try:
    sol = op.solve(); print("gpu", sol.objective, sol.newton_steps, sol.phase1_newton_steps)
except Exception as e:
    print("gpu failed:", e)
    if prob["x0"] is None:
        xf, ph = op.solver.phase_I()
        print("phase1 gpu s", ph.phase1_s, "stages", ph.outer_stages, ph.stage_newton_steps, "feasible by oracle test:", cnts.isSatisfiedStrictlyBy(xf))
        x0, s0, solp = O.phase_I_Analysis(cnts, eqs, O.SolverParams())
        print("phase1 oracle s", s0, solp.stage_newton_steps, np.linalg.norm(xf - x0))
        print("min slack at gpu xf:", (cnts.ub_all() - cnts.valuesAt(xf)).min(), " oracle:", (cnts.ub_all() - cnts.valuesAt(x0)).min())

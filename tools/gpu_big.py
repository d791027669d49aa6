import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
"""Large-configuration timing (BASELINE.json configs[3] and configs[4] shapes): a few Newton steps each."""
import json, sys, time
import numpy as np
import cvx_b200 as cb
import synthetic as P

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
h = cb.default_handle()
out = {}
def run(tag, prob, solver, steps):
    t0 = time.time()
    op = cb.from_dict(prob, solver, cb.SolverParams(stepLimit=2), h)
    up = time.time() - t0
    op.solve()                                  # warm-up (2 steps)
    op.solver.pars.stepLimit = steps
    h.profile_enable(True)
    sol = op.solve()
    n_syrk, syrk_ms, syrk_fl = h.profile_read()
    h.profile_enable(False)
    done = sol.executed_newton_steps + sol.phase1_executed_steps
    out[tag] = dict(steps=done, ms_per_step=sol.solve_ms / max(done, 1), steps_per_s=1e3 * done / sol.solve_ms,
                    upload_s=up, syrk_ms=syrk_ms / max(n_syrk, 1), syrk_tflops=syrk_fl / max(syrk_ms, 1e-9) / 1e9,
                    syrk_share=syrk_ms / sol.solve_ms)
    print(tag, json.dumps(out[tag]), flush=True)
    op.solver.problem.close()

if which == "c4":
    n, mh, p = 8192, 8192, 2048
    t0 = time.time(); prob = P.slab_qp(n, mh, p, 0); print("gen %.1fs" % (time.time() - t0), flush=True)
    F = mh * 2 * n * (n + 1) + n**3 / 3 + n * n * p + p * p * n + p**3 / 3
    run("c4_barrier_n8192_m16384_p2048", prob, "BR", 6)
    run("c4_primal_dual_n8192_m16384_p2048", prob, "PD", 6)
    for k in list(out):
        out[k]["tflops_algorithmic"] = F / (out[k]["ms_per_step"] * 1e-3) / 1e12
else:
    n, mh = 16384, 16384
    t0 = time.time(); prob = P.slab_lp(n, mh, 0, 0, feasible_start=False); print("gen %.1fs" % (time.time() - t0), flush=True)
    run("c5_phase1_n16385_m32768", prob, "BR", 4)
    m = 2 * mh
    F = m * (n + 1) * (n + 2) + (n + 1)**3 / 3
    for k in list(out):
        out[k]["tflops_algorithmic"] = F / (out[k]["ms_per_step"] * 1e-3) / 1e12
print(json.dumps(out, indent=1))

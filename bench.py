#!/usr/bin/env python
"""bench.py -- FP64 Newton (KKT) steps/s of the interior-point hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c5|...] [--impl cvxb|reference]

Headline workload (default, `config.workload`): BASELINE.json configs[3] -- the configuration north_star's target
"FP64 KKT Newton steps/sec at n=8192" is quoted on: one dense QP, n = 8192 variables, m = 16384 inequality rows,
p = 2048 equality rows, PRIMAL-DUAL solver (PrimalDualSolver.solve_withEQs, PrimalDualSolver.scala:550-621), FP64,
synthetic seeded data (3.4 GB of device state: fits one GPU).  A "step" is ONE primal-dual Newton iteration: Hessian
H_pd = hess f + G' diag(-lam/f) G assembled as one weighted SYRK on FP64 DMMA, residual GEMVs, Ruiz equilibration,
Cholesky of H with the forward substitution of [DA', Dq] riding along, Schur complement + its Cholesky, back
substitution, residual check, delta-lambda, residual line search, update.  The K timed steps are taken from the start
of full solves of pre-uploaded seeded instances (device time by CUDA events on the library's stream, barrier +
synchronize on both sides, max over ranks).  N > 1: the path shards only across independent problems, so each rank
runs its own instances (weak scaling, "replicas only", no data-path collective).

The JSON line also carries
  roofline               the dominant kernel (Hessian-assembly SYRK), timed per launch with CUDA events inside the
                         timed region, against the FP64 DMMA issue-rate peak measured in the same run
  roofline_chol_trailing the Cholesky trailing updates A22 -= A21 A21' (north_star's >= 60 % target), same method
  step_breakdown         share of the step spent in each timed range (SYRK / Ruiz / factorisation+TRSM / Schur / GEMV)
  hbm_kernels            the HBM-bound GEMVs over G (1.07 GB, far above the 126 MB L2) against MEASURED_PEAKS.json
  cpu_baseline           the CPU oracle (numpy/LAPACK restatement of the reference) on this box's host cores
  e2e                    ONE FULL primal-dual solve through the public API with HOST (pinned) buffers -- problem
                         creation, upload, every Newton step to termination, download -- divided by its steps
  legs                   the other single-GPU configs (C2 KL barrier with phase I, C5 LP phase I) as extra records
  batched                configs[2]: B = 8192 problems n=64, m=128 sharded over the ranks + one NCCL all-gather
--impl reference times the reference's CPU algorithm (oracle/) on the host cores, same workload, same kind of steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: description, generator (synthetic.py), its arguments, solver type
    "c4": dict(desc="dense QP n=8192 m=16384 p=2048, primal-dual solver (BASELINE.json configs[3], the north-star config)",
               gen="slab_qp", args=dict(n=8192, m_half=8192, p=2048), solver="PD"),
    "c4s": dict(desc="dense QP n=1024 m=2048 p=256, primal-dual solver (smoke size of c4)", gen="slab_qp",
                args=dict(n=1024, m_half=1024, p=256), solver="PD"),
    "c4b": dict(desc="dense QP n=8192 m=16384 p=2048, barrier solver", gen="slab_qp", args=dict(n=8192, m_half=8192, p=2048),
                solver="BR"),
    "c1": dict(desc="small dense LP n=100 m=200 p=20, barrier", gen="slab_lp", args=dict(n=100, m_half=100, p=20), solver="BR"),
    "c2": dict(desc="KL-distance minimisation n=2000 m=4000 p=500, barrier, phase I included", gen="kl_random",
               args=dict(n=2000, m_h=2000, p_extra=499), solver="BR"),
    "c2s": dict(desc="KL-distance minimisation n=400 m=800 p=100 (smoke size)", gen="kl_random",
                args=dict(n=400, m_h=400, p_extra=99), solver="BR"),
    "c5": dict(desc="random dense LP n=16384 m=32768 p=0, phase I then barrier", gen="slab_lp",
               args=dict(n=16384, m_half=16384, p=0, feasible_start=False), solver="BR"),
    "c5s": dict(desc="random dense LP n=2048 m=4096 p=0, phase I then barrier (smoke size of c5)", gen="slab_lp",
                args=dict(n=2048, m_half=2048, p=0, feasible_start=False), solver="BR"),
}


def make_problem(workload, seed):
    import synthetic as P                  # input generation only (numpy; no oracle code on the GPU arm)
    w = WORKLOADS[workload]
    return getattr(P, w["gen"])(seed=seed, **w["args"])


def f_step(n, m, p):
    """Algorithmic flops of one Newton step."""
    return m * n * (n + 1) + n ** 3 / 3 + n * n * p + p * p * n + p ** 3 / 3 + 4 * m * n + 4 * n * n + 8 * p * n


class ClockSampler:
    """SM clock + throttle reasons while the timed region runs (B200_PROFILING.md's clocks line).  Sampled in-process
    through NVML every 20 ms (a query costs microseconds); `nvidia-smi -lms` as a fallback -- its start-up (NVML
    init in a second process, under the driver's locks) overlapped the short timed region of this latency-bound
    workload and showed up as occasional slow runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None
        self.nv, self.nvh, self.stop_flag, self.samples = None, None, False, []
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = None
            try:
                import torch
                u = getattr(torch.cuda.get_device_properties(device), "uuid", None)
                if u is not None:
                    hnd = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(u)).encode())
            except Exception:
                hnd = None
            if hnd is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = device
                if vis:
                    try:
                        idx = int(vis.split(",")[device])
                    except Exception:
                        idx = device
                hnd = pynvml.nvmlDeviceGetHandleByIndex(idx)
            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)      # probe
            self.nv, self.nvh = pynvml, hnd
        except Exception:
            self.nv = None

    def _sample_nvml(self):
        nv, hnd = self.nv, self.nvh
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(hnd, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(hnd, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(hnd)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
                self.samples.append((float(sm), float(mx), {k for k, b_ in bits.items() if rs & b_}))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._sample_nvml, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            sm = [s_[0] for s_ in self.samples]
            mx = [s_[1] for s_ in self.samples]
            reasons = set()
            for s_ in self.samples:
                reasons |= s_[2]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def steps_of(sol):
    return int(sol.executed_newton_steps + sol.phase1_executed_steps)


def run_steps(instances, budget, stats, cycle):
    """Consume `budget` Newton steps from the start of full solves; returns device ms (CUDA events on the library's
    stream, cvxb_solution.solve_ms).  cycle: instances with a given feasible start can be solved again from scratch
    (every solve recomputes everything from the uploaded problem data)."""
    ms = 0.0
    while budget > 0:
        progressed = False
        for op in instances:
            if budget <= 0:
                break
            op.solver.pars.stepLimit = budget
            sol = op.solve()
            done = steps_of(sol)
            ms += sol.solve_ms
            budget -= done
            stats["steps"] += done
            stats["solves"] += 1
            stats["last"] = sol
            progressed = progressed or done > 0
        if not cycle or not progressed:
            break
    if budget > 0:
        raise RuntimeError("not enough problem instances for the requested number of steps")
    return ms


def _blas_threads(n=None):
    """BLAS threads of the CPU arm: torchrun exports OMP_NUM_THREADS=1 to every rank, which would throttle the
    reference's LAPACK calls; the CPU arm uses every host core the box has."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n or os.cpu_count())
    except Exception:
        pass
    return n or os.cpu_count()


def cpu_reference_leg(workload, steps, warmup, seed=0):
    """The reference's CPU algorithm (oracle restatement: numpy + LAPACK dpotrf/dtrtrs/dgemm, all host threads the
    BLAS uses) on a bounded sample of the SAME workload and the SAME kind of steps as the GPU arm: `warmup` untimed +
    `steps` timed Newton steps from the start of a solve.  Returns (steps/s, steps done, seconds, sample text)."""
    from oracle import cvx_oracle as O
    from oracle import problems as P
    cores = _blas_threads()
    w = WORKLOADS[workload]
    prob = make_problem(workload, seed)
    objF, cnts, eqs = P.to_oracle(prob)
    if w["solver"] == "PD":
        # primal-dual iterations (corrected solver: bug_compat=False, like the GPU arm) from the strictly feasible start
        pd = O.PrimalDual(objF, cnts, eqs, O.SolverParams(), False, False)
        marks = []
        inner = pd.newton_direction

        def stamped(*a, **k):
            marks.append(time.perf_counter())
            return inner(*a, **k)
        pd.newton_direction = stamped
        sol = pd.solve(max_steps=warmup + steps)
        marks.append(time.perf_counter())
        done_all = len(marks) - 1
        w_ = min(warmup, max(done_all - 1, 0))
        done = done_all - w_
        dt = marks[-1] - marks[w_]
        return done / dt, done, dt, ("primal-dual Newton iterations %d..%d of a solve from the strictly feasible start "
                                     "(%d untimed warm-up iterations before them)" % (w_ + 1, done_all, w_)), cores
    if cnts.feasiblePoint is None and "qstar" not in prob:      # workloads that really need phase I (c5)
        pars = O.SolverParams(maxIter=max(1, steps + warmup))
        t0 = time.perf_counter()
        x0, s, sol = O.phase_I_Analysis(cnts, eqs, pars)
        dt = time.perf_counter() - t0
        return sol.newton_steps / dt, sol.newton_steps, dt, "phase-I Newton steps from pointWhereDefined", cores
    # barrier workloads: K/2 phase-I steps + K/2 main-phase steps, as the GPU arm does (c2), or all main-phase
    done, dt = 0, 0.0
    sample = []
    if cnts.feasiblePoint is None:
        k1 = steps // 2
        if k1 > 0:
            pars = O.SolverParams(maxIter=k1)
            t0 = time.perf_counter()
            x0, s, sol = O.phase_I_Analysis(cnts, eqs, pars)
            dt += time.perf_counter() - t0
            done += sol.newton_steps
            sample.append("%d phase-I Newton steps from pointWhereDefined" % sol.newton_steps)
        steps -= k1
        cnts = cnts.addFeasiblePoint(prob["qstar"])
    bf = O.BarrierFunctions(objF, cnts)
    x = np.array(cnts.feasiblePoint)
    t = 1.0
    todo_w, todo, d2 = warmup, steps, 0
    while todo > 0:
        k = todo_w if todo_w > 0 else todo
        pars = O.SolverParams(maxIter=k)
        t0 = time.perf_counter()
        sol = O.equalityConstrainedSolve(bf, t, x, eqs.A, eqs.b, pars) if eqs is not None else O.unconstrainedSolve(bf, t, x, pars)
        el = time.perf_counter() - t0
        x = sol.x
        if todo_w > 0:
            todo_w -= max(1, sol.newton_steps)
        else:
            todo -= max(1, sol.newton_steps)
            d2 += sol.newton_steps
            dt += el
        if not sol.maxedOut:
            t *= 10.0
    done += d2
    sample.append("%d barrier-stage Newton steps from the strictly feasible point" % d2)
    return done / dt, done, dt, " + ".join(sample), cores


def cpu_modes_leg():
    """the CPU baseline in its two modes on the configs where the literal one is feasible --
    `literal` = the reference's actual per-constraint foldLeft Hessian (BarrierSolver.scala:303-315: m rank-one updates
    with fresh n x n temporaries), `vectorised` = one dgemm.  C1 (n=100, m=200, p=20, barrier) as Newton steps/s, C3
    (n=64, m=128 batched mix, 8 problems) as solves/s.  At C4 / C5 the literal loop would move ~44 TB per Hessian."""
    from oracle import cvx_oracle as O
    from oracle import problems as P
    _blas_threads()
    out = {}
    pr = make_problem("c1", 0)
    objF, cnts, eqs = P.to_oracle(pr)
    for mode, lit in (("literal", True), ("vectorised", False)):
        t0 = time.perf_counter()
        # maxIter 100 instead of 1000: this instance has a stage in which the reference repeats one identical step until
        # maxIter (||b-Ax|| stays above tol, EqualityConstrainedSolver.scala:49); every repeat is a full Newton step on the CPU
        sol, _ = O.solveProblem(objF, cnts, eqs, "BR", pars=O.SolverParams(maxIter=100), literal=lit)
        dt = time.perf_counter() - t0
        steps = int(sol.newton_steps)
        out["c1_" + mode] = {"value": steps / dt, "unit": "steps/s", "steps": steps, "seconds": dt}
    import synthetic as S
    probs = [S.batched_problem(i, 64, 128, 1000) for i in range(8)]
    for mode, lit in (("literal", True), ("vectorised", False)):
        t0 = time.perf_counter()
        for q in probs:
            o_, c_, e_ = P.to_oracle(q)
            O.solveProblem(o_, c_, e_, "BR", literal=lit)
        dt = time.perf_counter() - t0
        out["c3_" + mode] = {"value": len(probs) / dt, "unit": "solves/s", "problems": len(probs), "seconds": dt}
    return out


def pinned_problem(prob):
    """The problem's matrices as column-major (Breeze layout) views on pinned host memory: what a long-running caller
    hands to the C ABI.  Returns (problem dict, bytes uploaded per cvxb_problem_create, keep-alive list)."""
    import torch
    out, keep, nbytes = dict(prob), [], 0
    for k_ in ("G", "A", "P"):
        if prob.get(k_) is not None:
            a = np.asarray(prob[k_], dtype=np.float64)
            tpin = torch.empty(a.shape[::-1], dtype=torch.float64).pin_memory()    # (cols, rows) C-order == (rows, cols) F-order
            tpin.numpy()[...] = a.T
            keep.append(tpin)
            out[k_] = tpin.numpy().T
            nbytes += a.nbytes
    for k_ in ("a", "ub", "b", "rvec", "xdef", "x0"):
        if isinstance(prob.get(k_), np.ndarray):
            nbytes += prob[k_].nbytes
    return out, nbytes, keep


def roofline_from_range(rng, kernel, peak, peak_source, dev_ms, extra=None):
    """rng = (count, total ms, total algorithmic flops) of a timed range (Handle.profile_read_range)."""
    cnt, ms, work = rng
    ach = work / (ms / 1e3) / 1e12 if ms > 0 else 0.0
    r = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
         "frac": ach / peak if peak else None, "traffic": None, "launches": cnt, "avg_launch_ms": ms / max(cnt, 1),
         "flops_per_launch": work / max(cnt, 1), "peak_source": peak_source, "share_of_step": ms / dev_ms if dev_ms > 0 else None}
    if extra:
        r.update(extra)
    return r


def extra_leg(cb, h, workload, steps, warmup_steps):
    """One of the other single-GPU configs as an extra record: `steps` Newton steps from the start of a full solve."""
    w = WORKLOADS[workload]
    pr = make_problem(workload, 7)
    n, m = pr["n"], pr["G"].shape[0]
    p = 0 if pr.get("A") is None else pr["A"].shape[0]
    two_kinds = pr.get("x0") is None and "qstar" in pr
    ops = [cb.from_dict(pr, w["solver"], cb.SolverParams(), h)]
    if two_kinds:
        q = dict(pr)
        q["x0"] = pr["qstar"].copy()
        ops.append(cb.from_dict(q, w["solver"], cb.SolverParams(), h))
    for op in ops:                                   # warm-up: same kernels, same shapes
        op.solver.pars.stepLimit = warmup_steps
        op.solve()
    if two_kinds:                                    # the warm-up left a feasible point behind: start again from scratch
        ops[0].solver.problem.close()
        ops[0] = cb.from_dict(pr, w["solver"], cb.SolverParams(), h)
    h.profile_enable(True)
    ms, done, kinds = 0.0, 0, []
    per = max(1, steps // len(ops))
    for op in ops:
        op.solver.pars.stepLimit = per
        sol = op.solve()
        ms += sol.solve_ms
        done += steps_of(sol)
        kinds.append("%d phase-I + %d main-phase" % (sol.phase1_executed_steps, sol.executed_newton_steps))
    cnt, sms, sfl = h.profile_read()
    h.profile_enable(False)
    for op in ops:
        op.solver.problem.close()
    nn = n + 1 if (pr.get("x0") is None and not two_kinds) else n      # phase-I dimension for c5
    return {"workload": workload + ": " + w["desc"], "metric": "newton_steps_per_sec", "value": done / (ms / 1e3), "unit": "steps/s",
            "steps": done, "ms_per_step": ms / max(done, 1), "n": n, "m": m, "p": p, "steps_counted": " ; ".join(kinds),
            "tflops": f_step(nn, m + (2 * p if nn != n else 0), 0 if nn != n else p) * done / (ms / 1e3) / 1e12,
            "hessian_syrk": {"launches": cnt, "avg_launch_ms": sms / max(cnt, 1),
                             "achieved_tflops": sfl / (sms / 1e3) / 1e12 if sms > 0 else None,
                             "share_of_step": sms / ms if ms > 0 else None}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="cvxb", choices=["cvxb", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the extra single-GPU legs (c2, c5)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--legs", default="c1,c2,c5", help="comma-separated extra workloads run at N=1 after the headline")
    ap.add_argument("--batch", type=int, default=8192, help="problems in the batched leg (configs[2])")
    ap.add_argument("--cpu-steps", type=int, default=0, help="Newton steps of the CPU-baseline sample (0 = auto)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(1, args.steps), max(0, args.warmup)
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return 0
        val, done, dt, sample, cores = cpu_reference_leg(args.workload, K, W)
        line = {"impl": "reference", "metric": "newton_steps_per_sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1000.0 * dt / max(done, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload + ": " + w["desc"], **w["args"], "solver": w["solver"]},
                "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
                                 "sample": "%d %s; %.1f s (numpy/LAPACK oracle of the Scala reference, BLAS threads = %d; no JVM "
                                           "on this image)" % (done, sample, dt, cores)},
                "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the cvxb path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import cvx_b200 as cb
    from cvx_b200 import _lib
    h = _lib.Handle(local_rank)

    # ---- inputs: seeded instances, generated on the host and uploaded BEFORE the timed region ----------
    base_seed = 100 * rank
    prob0 = make_problem(args.workload, base_seed)
    n, m = prob0["n"], prob0["G"].shape[0]
    p = 0 if prob0.get("A") is None else prob0["A"].shape[0]
    solver_type = w["solver"]
    # two kinds of instance for configs whose solve has two halves (c2: phase I in dimension n+1 without equalities, then
    # the main phase with p equalities): K/2 steps from each, each taken from the start of full solves
    two_kinds = prob0.get("x0") is None and "qstar" in prob0
    cycle = prob0.get("x0") is not None           # a given feasible start: the same instance can be solved again from scratch

    def make_instances(seed0, count, feasible, first=None):
        out = []
        for i in range(count):
            pr = first if (first is not None and i == 0) else make_problem(args.workload, seed0 + i)
            if feasible:
                pr = dict(pr)
                pr["x0"] = pr["qstar"].copy()
            out.append(cb.from_dict(pr, solver_type, cb.SolverParams(), h))
        return out

    K1 = K // 2 if two_kinds else 0
    K2 = K - K1
    if cycle:
        inst1, inst2 = [], make_instances(base_seed, 2 if n <= 8192 else 1, False, prob0)
        warm = inst2                       # every timed instance has run once (first-use allocations, pool growth)
    else:
        inst1 = make_instances(base_seed, (K1 + 89) // 90 + 1, False) if K1 else []
        inst2 = make_instances(base_seed + 40, (K2 + 29) // 30 + 1, two_kinds)
        warm = make_instances(base_seed + 90, 1, False) + (make_instances(base_seed + 91, 1, True) if two_kinds else [])
    del prob0
    h.synchronize()

    # ---- warm-up: W untimed Newton steps (same kernels, same shapes) ------------------------------------
    st = {"steps": 0, "solves": 0, "last": None}
    for wi in warm:
        run_steps([wi], max(W, 3), st, False)
    h.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K Newton steps ------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    h.profile_enable(True)
    l0 = h.launches
    st = {"steps": 0, "solves": 0, "last": None}
    t0 = time.perf_counter()
    dev_ms = (run_steps(inst1, K1, st, False) if K1 else 0.0) + run_steps(inst2, K2, st, cycle)
    h.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = h.launches - l0
    n_syrk, syrk_ms, syrk_flops = h.profile_read()
    ranges = {k_: h.profile_read_range(k_) for k_ in ("chol_trailing_update", "factor_h_with_trsm", "schur_syrk", "ruiz", "gemv_g",
                                                         "chol_lookahead_phases", "chol_trsm_right")}
    h.profile_enable(False)
    barrier()
    clk = clocks.stop()
    steps_done = st["steps"]
    tmax = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(steps_done), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = tmax.tolist()
    total_steps, total_launches = tot.tolist()
    value = total_steps / (dev_ms_max / 1e3)
    last = st["last"]
    for op_ in {id(o): o for o in inst1 + inst2 + warm}.values():
        op_.solver.problem.close()
    del inst1, inst2, warm

    # ---- e2e: ONE FULL solve through the public API with host (pinned) buffers -----------------------------------------
    # Timed: problem creation (device allocation from the handle's pool), upload of every matrix and vector, every Newton
    # step to termination, download of the solution, a host read of the result.  Steady state of a long-running caller:
    # earlier problems were destroyed, their device memory sits in the handle's pool.
    e2e = None
    if not args.no_e2e:
        # the SAME seeded instance on every rank ("replicas only"): equal iteration counts, so that total steps / slowest
        # rank's time measures the hardware and not the spread of iteration counts between different random instances
        e2e_prob, h2d, keep = pinned_problem(make_problem(args.workload, 50))
        barrier()
        t0 = time.perf_counter()
        op = cb.from_dict(e2e_prob, solver_type, cb.SolverParams(), h)
        sol = op.solve()
        xsum = float(np.sum(sol.x))                 # result read back on the host
        h.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_steps = steps_of(sol)
        d2h = sol.x.nbytes + (0 if sol.lam is None else sol.lam.nbytes) + (0 if sol.nu is None else sol.nu.nbytes)
        d2h += e2e_steps * (128 * 8 + 64 * 4)       # the status block the host reads once per Newton step
        e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        e2e_n = torch.tensor([float(e2e_steps)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(e2e_n, op=dist.ReduceOp.SUM)
        e2e = {"value": e2e_n.item() / e2e_t.item(), "unit": "steps/s", "h2d_bytes_per_step": h2d / max(e2e_steps, 1),
               "d2h_bytes_per_step": d2h / max(e2e_steps, 1), "steps": e2e_steps, "seconds": e2e_s, "device_ms": sol.solve_ms,
               "what": "one full %s solve to termination through the public API: cvxb_problem_create with pinned host buffers "
                       "(%.2f GB uploaded), all Newton steps, solution downloaded; value = steps of all ranks / wall seconds of "
                       "the slowest rank (every rank solves the same seeded instance)" % (solver_type, h2d / 1e9),
               "solution": {"objective": sol.objective, "dualityGap": sol.dualityGap, "equalityGap": sol.equalityGap,
                            "normDualResidual": sol.normDualResidual, "newton_steps": sol.newton_steps,
                            "phase1_newton_steps": sol.phase1_newton_steps, "maxedOut": sol.maxedOut},
               "checksum": xsum}
        op.solver.problem.close()
        del op, e2e_prob, keep

    # ---- batched leg (BASELINE.json configs[2]): B = 8192 independent n=64, m=128 problems (half KL with p=1, half QP
    # with p=0), contiguous shards over the ranks, one CTA per problem, results left on the device, then ONE NCCL
    # all-gather of the packed records (x + objective + gap + status + steps + stages) and one device-to-host copy.
    # Strong scaling: B is fixed as N grows.
    batched = None
    if not args.no_batched:
        import synthetic as P
        Bt = args.batch
        lo, hi = cb.shard_range(Bt, rank, world)
        bprobs = [P.batched_problem(i, 64, 128, 1000) for i in range(lo, hi)]
        solver = cb.BatchedBarrierSolver(cb.pack_problems(bprobs), cb.SolverParams(), h)
        solver.solve(download=False)                      # warm-up (same shapes)
        if world > 1:
            cb.gather_solutions(solver, Bt, 64)           # and of the exchange: NCCL connects its rings on first use
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            bsol = solver.solve(download=False)
            g = cb.gather_solutions(solver, Bt, 64)
            conv, mxs, nsteps = g["converged"], g["max_newton_steps"], float(g["newton_steps"].sum())
        else:
            bsol = solver.solve()
            conv, mxs, nsteps = int((bsol.status == 0).sum()), int(bsol.newton_steps.max()), float(bsol.newton_steps.sum())
        torch.cuda.synchronize()
        bt = torch.tensor([time.perf_counter() - t0, bsol.solve_ms / 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
        wall_s, dev_s = bt.tolist()
        fl_b = 6.69e5 * nsteps                  # algorithmic flops of one n=64, m=128 Newton step
        batched = {"metric": "batched_barrier_solves_per_sec", "value": Bt / wall_s, "unit": "solves/s", "B": Bt, "n": 64,
                   "m": 128, "scaling": "strong", "device_only_value": Bt / dev_s, "converged": conv,
                   "max_newton_steps": mxs, "newton_steps_per_sec": nsteps / wall_s,
                   "tflops_algorithmic": fl_b / dev_s / 1e12,
                   "roofline": {"bound": "latency (FP64 dependent-issue chains inside one CTA per problem; not tensor- or HBM-bound)",
                                "achieved": fl_b / dev_s / 1e12, "unit": "TFLOP/s",
                                "note": "fraction of the FP64 tensor peak is filled in below; per-phase clock64 split (in-CTA Cholesky "
                                        "36 %, of it the 64 sequential pivots 67 %; Ruiz 19 %; Hessian DMMA 15 %; triangular solves 15 %), "
                                        "residency scaling and the ncu capture (issue slots 28 % busy, top stall barrier): "
                                        "profiles/r2_batched_phase_split.txt, profiles/r2_ncu_full_batched_1024.txt"},
                   "includes": "kernel + ONE NCCL all-gather of the packed device records + one device-to-host copy"
                               if world > 1 else "kernel + device-to-host of the results"}
        solver.close()
        if world == 1:
            # the phase-I variant: no feasible starts, ConstraintSet.phase_I_Analysis runs inside each CTA first
            try:
                Bp = min(Bt, 2048)
                pprobs = [P.batched_problem_phase1(i, 63, 126) for i in range(Bp)]
                psolver = cb.BatchedBarrierSolver(cb.pack_problems(pprobs), cb.SolverParams(), h)
                psolver.solve(download=False)
                t0 = time.perf_counter()
                psol = psolver.solve()
                pwall = time.perf_counter() - t0
                batched["phase1_variant"] = {
                    "B": Bp, "n": 63, "m": 126, "value": Bp / pwall, "unit": "solves/s", "device_only_value": Bp / (psol.solve_ms / 1e3),
                    "converged": int((psol.status == 0).sum()), "phase1_newton_steps_mean": float(psol.phase1_newton_steps.mean()),
                    "newton_steps_mean": float(psol.newton_steps.mean()),
                    "what": "problems without a feasible start: phase I (n + 1 = 64 variables, m + 2p <= 128 rows) then the barrier "
                            "solve, all inside the problem's CTA"}
                psolver.close()
            except Exception as e:      # never lose the headline to a leg
                batched["phase1_variant"] = {"failed": repr(e)}

    # ---- the other single-GPU configs as extra legs (N = 1 only) -------------------------------------------------------
    legs = {}
    if world == 1 and not args.no_legs:
        for name in [s_ for s_ in args.legs.split(",") if s_ and s_ != args.workload]:
            try:
                legs[name] = extra_leg(cb, h, name, 20 if WORKLOADS[name]["args"].get("n", 0) <= 4096 else 4, 3 if name != "c5" else 1)
            except Exception as e:          # never lose the headline to a leg
                legs[name] = {"workload": name, "failed": repr(e)}

    if rank == 0:
        # ---- rooflines -------------------------------------------------------------------------------------------------
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_file)) if os.path.exists(peaks_file) else {}
        pk_ms, pk_fl = h.bench_kernel(0, 30000, 0, 1)
        fp64_peak = pk_fl / pk_ms / 1e9
        peak_source = ("FP64 DMMA issue-rate probe measured in this run (MEASURED_PEAKS.json has HBM and bf16 only: hbm_gbs=%s); "
                       "cuBLAS DGEMM 8192^3 on this pool measured 35.5 TFLOP/s" % peaks.get("hbm_gbs"))
        achieved = syrk_flops / (syrk_ms / 1e3) / 1e12 if syrk_ms > 0 else 0.0
        big = n >= 8192
        roofline = {"bound": "tensor",
                    "kernel": "gemm_dmma_streamk_kernel<TN> (Hessian assembly G' diag(w) G, FP64 DMMA m8n8k4, persistent stream-K grid)",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                    "traffic": 7.54e9 + 0.54e9 if big else (111.9e6 if args.workload == "c2" else None),
                    "traffic_source": ("dram__bytes_read+write per launch, ncu --set full, profiles/r2_ncu_full_syrk_hessian_c4_tile_order.txt "
                                       "(algorithmic: 1.07 GB read of G + 0.54 GB write of H; the rest: every wave of 148 tiles streams its "
                                       "24 operand panels of 16.8 MB once -- K = 16384 is far longer than L2 can keep across waves; 13.7 GB "
                                       "before the supertile order and the short stream-K tail; 3 % of DRAM peak: the kernel is "
                                       "tensor-bound)") if big else None,
                    "launches": n_syrk, "avg_launch_ms": syrk_ms / max(n_syrk, 1), "flops_per_launch": syrk_flops / max(n_syrk, 1),
                    "peak_source": peak_source, "share_of_step": syrk_ms / dev_ms if dev_ms > 0 else None}
        chol = roofline_from_range(
            ranges["chol_trailing_update"],
            "gemm_dmma_persist_kernel<NT> (Cholesky trailing updates A22 -= A21 A21' of the tile-DAG schedule, lower triangle, "
            "K = 1024..2048, n = %d: persistent grid on 140 of 148 SMs, timed on the bulk stream WHILE the chain of the next "
            "diagonal block runs on the other SMs; recursive schedule below n = 5120: gemm_dmma_streamk_kernel<NT>)" % n,
            fp64_peak, peak_source, dev_ms)
        # the same contraction timed alone on the whole machine at three shapes of the n = 8192 factorisation
        alone = {}
        for tag, (nn_, kk_) in (("trailing_6144x6144_K2048", (6144, 2048)), ("top_level_4096x4096_K4096", (4096, 4096)),
                                ("rank128_update_8064x8064_K128", (8064, 128))):
            try:
                ms_k, fl_k = h.bench_kernel(2, nn_, kk_, 5)
                alone[tag] = {"ms": ms_k, "tflops": fl_k / ms_k / 1e9, "frac": fl_k / ms_k / 1e9 / fp64_peak}
            except Exception as e:
                alone[tag] = {"failed": repr(e)}
        chol["timed_alone"] = alone
        if batched and "roofline" in batched:
            batched["roofline"]["peak"] = fp64_peak
            batched["roofline"]["frac"] = batched["roofline"]["achieved"] / fp64_peak if fp64_peak else None
        fcnt, fms, fwk = ranges["factor_h_with_trsm"]
        breakdown = {"hessian_syrk": syrk_ms / dev_ms if dev_ms > 0 else None}
        for k_, (c_, ms_, wk_) in ranges.items():
            breakdown[k_] = {"share_of_step": ms_ / dev_ms if dev_ms > 0 else None, "ms_per_step": ms_ / max(steps_done, 1),
                             "count": c_}
        if fms > 0:
            breakdown["factor_h_with_trsm"]["tflops"] = fwk / fms / 1e9
        gc_, gms_, gby_ = ranges["gemv_g"]
        # the HBM-bound kernels of the step (slack / gradient / line-search GEMVs over G): in-solve and timed alone
        hbm_peak = peaks.get("hbm_gbs")
        hbm = {}
        if gms_ > 0:
            gbs = gby_ / gms_ / 1e6
            hbm["gemv over G inside the solve (G x and G'(1/f), CUDA events in the timed region)"] = {
                "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else None,
                "bytes_per_launch": gby_ / max(gc_, 1), "launches": gc_}
        for which, name in ((7, "gemv_n (G x, G d) alone"), (8, "gemv_t (G' w) alone")):
            try:
                ms_k, by_k = h.bench_kernel(which, n, m, 10)
                gbs = by_k / ms_k / 1e6
                hbm[name] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": gbs / hbm_peak if hbm_peak else None, "bytes_per_launch": by_k,
                             "note": "G is %.0f MB (L2 is 126 MB)" % (by_k / 1e6)}
            except Exception as e:
                hbm[name] = {"failed": repr(e)}
        cpu = None
        if not args.no_cpu_baseline:
            cs = args.cpu_steps or (2 if big else 20)
            try:
                val, done, dt, sample, cores = cpu_reference_leg(args.workload, cs, 1)
                cpu = {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
                       "sample": "%d %s, %.1f s of CPU time (numpy/LAPACK oracle, BLAS threads = %d; the Scala reference needs a JVM "
                                 "this image lacks)" % (done, sample, dt, cores)}
            except Exception as e:      # never lose the GPU line to a CPU-side problem
                cpu = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        cpu_modes = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu_modes = cpu_modes_leg()
            except Exception as e:
                cpu_modes = {"failed": repr(e)}
        line = {"metric": "newton_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms_max / max(steps_done, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload + ": " + w["desc"], "n": n, "m": m, "p": p, "solver": solver_type,
                           "parallelism": "replicas only (one independent problem stream per GPU, no data-path collective)",
                           "l2": "working set per step (G, scaled G, H, L, RHS: %.0f MB) exceeds the 126 MB L2; no flush needed"
                                 % ((2 * m * n + 3 * n * n + n * (p + 1)) * 8 / 1e6),
                           "solves_started": st["solves"],
                           "steps_counted": ("%d primal-dual Newton iterations, each from the start of full solves" % K2) if solver_type == "PD"
                           else ("%d phase-I Newton steps (dimension n+1, p = 0) + %d main-phase Newton steps (p equalities, Schur "
                                 "complement), each from the start of full solves" % (K1, K2))},
                "clocks": clk, "wall_ms_per_step": wall_ms_max / max(steps_done, 1),
                "flops_per_step": f_step(n, m, p), "tflops": f_step(n, m, p) * value / world / 1e12,
                "tflops_frac_of_fp64_tensor_peak": f_step(n, m, p) * value / world / 1e12 / fp64_peak if fp64_peak else None,
                "e2e": e2e, "gpu_launches": int(total_launches), "roofline": roofline, "roofline_chol_trailing": chol,
                "step_breakdown": breakdown, "hbm_kernels": hbm, "cpu_baseline": cpu, "cpu_baseline_modes": cpu_modes, "batched": batched, "legs": legs,
                "last_solution": {"objective": last.objective, "outer_stages": last.outer_stages,
                                  "newton_steps": last.newton_steps, "phase1_newton_steps": last.phase1_newton_steps}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

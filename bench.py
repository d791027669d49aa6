#!/usr/bin/env python
"""bench.py -- FP64 Newton (KKT) steps/s of the interior-point hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c4|c5] [--impl cvxb|reference]

A "step" is ONE Newton step of the barrier method (Hessian + gradient assembly over the m inequality
constraints, KKT solve by Cholesky + Schur complement, backtracking line search, update) on the
workload BASELINE.json's metric is quoted on at one GPU: configs[1], the KL-distance minimisation
n=2000, m=4000 inequalities, p=500 equalities, barrier solver, phase I included (Dist_KL.apply
semantics), synthetic seeded data, FP64.  K timed steps are taken from the start of full solves
(several seeded instances, uploaded before the timed region).  N > 1: the path shards only across
independent problems, so each rank runs its own instances (weak scaling, "replicas only", no data-path
collective); time = max over ranks.

The JSON line also carries
  roofline      the dominant kernel (Hessian-assembly SYRK on FP64 DMMA), timed per launch with CUDA
                events on the library's stream inside the timed region
  cpu_baseline  the CPU oracle (numpy/LAPACK restatement of the reference) on this box's host cores
  e2e           the same metric through the public API with host buffers (upload + solve + download timed)
--impl reference times the reference's CPU algorithm (oracle/) on the host cores for the same config.
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
    # name: (generator name, kwargs, solver)
    "c1": dict(desc="small dense LP n=100 m=200 p=20, barrier", gen="slab_lp", args=dict(n=100, m_half=100, p=20)),
    "c2": dict(desc="KL-distance minimisation n=2000 m=4000 p=500, barrier, phase I included", gen="kl_random",
               args=dict(n=2000, m_h=2000, p_extra=499)),
    "c2s": dict(desc="KL-distance minimisation n=400 m=800 p=100 (smoke size)", gen="kl_random",
                args=dict(n=400, m_h=400, p_extra=99)),
    "c4b": dict(desc="dense QP n=8192 m=16384 p=2048, barrier solver", gen="slab_qp", args=dict(n=8192, m_half=8192, p=2048)),
    "c5": dict(desc="random dense LP n=16384 m=32768 p=0, phase I then barrier", gen="slab_lp",
               args=dict(n=16384, m_half=16384, p=0, feasible_start=False)),
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


def run_steps(instances, budget, stats):
    """Consume `budget` Newton steps from the start of full solves; returns device ms."""
    ms = 0.0
    for op in instances:
        if budget <= 0:
            break
        op.solver.pars.stepLimit = budget
        sol = op.solve()
        done = sol.executed_newton_steps + sol.phase1_executed_steps
        ms += sol.solve_ms
        budget -= done
        stats["steps"] += done
        stats["solves"] += 1
        stats["last"] = sol
    if budget > 0:
        raise RuntimeError("not enough problem instances for the requested number of steps")
    return ms


def cpu_reference_leg(workload, steps, warmup, seed=0):
    """The reference's CPU algorithm (oracle restatement: numpy + LAPACK dpotrf/dtrtrs/dgemm, all host
    threads the BLAS uses) on a bounded sample: `steps` Newton steps of the barrier stage loop started at
    the strictly feasible point the generator knows (no phase I needed on the CPU side)."""
    from oracle import cvx_oracle as O
    from oracle import problems as P
    prob = make_problem(workload, seed)
    if prob.get("x0") is None and "qstar" in prob:
        prob["x0"] = prob["qstar"].copy()
    if prob.get("x0") is None:
        prob["x0"] = None
    objF, cnts, eqs = P.to_oracle(prob)
    if cnts.feasiblePoint is None:      # workloads that really need phase I (c5): include it in the sample
        pars = O.SolverParams(maxIter=max(1, steps + warmup))
        t0 = time.perf_counter()
        x0, s, sol = O.phase_I_Analysis(cnts, eqs, pars)
        dt = time.perf_counter() - t0
        return sol.newton_steps / dt, sol.newton_steps, dt, "phase-I Newton steps from pointWhereDefined"
    bf = O.BarrierFunctions(objF, cnts)
    x = np.array(cnts.feasiblePoint)
    t = 1.0
    done, dt = 0, 0.0
    todo_w, todo = warmup, steps
    while todo > 0:
        k = todo_w if todo_w > 0 else todo
        pars = O.SolverParams(maxIter=k)
        t0 = time.perf_counter()
        if eqs is not None:
            sol = O.equalityConstrainedSolve(bf, t, x, eqs.A, eqs.b, pars)
        else:
            sol = O.unconstrainedSolve(bf, t, x, pars)
        el = time.perf_counter() - t0
        x = sol.x
        if todo_w > 0:
            todo_w -= max(1, sol.newton_steps)
        else:
            todo -= max(1, sol.newton_steps)
            done += sol.newton_steps
            dt += el
        if not sol.maxedOut:
            t *= 10.0
    return done / dt, done, dt, "barrier-stage Newton steps from the generator's strictly feasible point"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="cvxb", choices=["cvxb", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--batch", type=int, default=8192, help="problems in the batched leg (configs[2])")
    ap.add_argument("--cpu-steps", type=int, default=0, help="Newton steps of the CPU-baseline sample (0 = auto)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(1, args.steps), max(0, args.warmup)
    w = WORKLOADS[args.workload]
    prob0 = None

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count()
        val, done, dt, sample = cpu_reference_leg(args.workload, K, W)
        line = {"impl": "reference", "metric": "newton_steps_per_sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1000.0 * dt / max(done, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload + ": " + w["desc"], **w["args"]},
                "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
                                 "sample": "%d %s (numpy/LAPACK oracle of the Scala reference; no JVM on this image)" % (done, sample)},
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
    # Two kinds of instance so that the K timed steps cover both halves of the solve of this config:
    #   "phase1"  starts at pointWhereDefined = 1/n: phase-I Newton steps (dimension n+1, no equalities,
    #             m + 2p inequality rows, UnconstrainedSolver / choleskySolve path)
    #   "main"    starts at the generator's strictly feasible point: barrier Newton steps with the p
    #             equalities (EqualityConstrainedSolver / KKTSystem Schur-complement path)
    # K/2 steps are taken from each kind (all from "main" when the workload has a feasible start).
    base_seed = 100 * rank
    prob0 = make_problem(args.workload, base_seed)
    n, m = prob0["n"], prob0["G"].shape[0]
    p = 0 if prob0.get("A") is None else prob0["A"].shape[0]
    two_kinds = prob0.get("x0") is None and "qstar" in prob0

    def feasible_variant(pr):
        q = dict(pr)
        q["x0"] = pr["qstar"].copy()
        return q

    def make_instances(seed0, count, feasible):
        out = []
        for i in range(count):
            pr = make_problem(args.workload, seed0 + i)
            if feasible:
                pr = feasible_variant(pr)
            out.append(cb.from_dict(pr, "BR", cb.SolverParams(), h))
        return out

    K1 = K // 2 if two_kinds else 0            # phase-I steps
    K2 = K - K1                                 # main-phase steps
    inst1 = make_instances(base_seed, (K1 + 89) // 90 + 1, False) if K1 else []       # >= 90 phase-I steps per solve
    inst2 = make_instances(base_seed + 40, (K2 + 29) // 30 + 1, two_kinds)            # >= 30 main-phase steps per solve
    warm = make_instances(base_seed + 90, 1, False) + (make_instances(base_seed + 91, 1, True) if two_kinds else [])
    h.synchronize()

    # ---- warm-up: W untimed Newton steps (same kernels, same shapes) ------------------------------------
    st = {"steps": 0, "solves": 0, "last": None}
    for wi in warm:
        run_steps([wi], min(max(W, 3), 25), st)      # a single solve has > 25 steps; more warm-up adds nothing
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
    dev_ms = (run_steps(inst1, K1, st) if K1 else 0.0) + run_steps(inst2, K2, st)
    h.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = h.launches - l0
    n_syrk, syrk_ms, syrk_flops = h.profile_read()
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

    # ---- e2e: public API with host buffers; upload + solve + download inside the timed region ------------
    # steady state of a long-running service: earlier problems have been destroyed, their device memory sits in the
    # handle's pool; the timed call still creates its problem, uploads, solves and downloads
    for op_ in inst1 + inst2 + warm:
        op_.solver.problem.close()
    del inst1, inst2, warm
    e2e_prob = make_problem(args.workload, base_seed + 50)
    if two_kinds and K2 >= K1:
        pass        # e2e runs the user's call as is: full problem from pointWhereDefined (phase I first)
    pinned = {}
    for k_ in ("G", "A", "P"):
        if e2e_prob.get(k_) is not None:
            a = np.asfortranarray(e2e_prob[k_])
            tpin = torch.empty(a.shape[::-1], dtype=torch.float64).pin_memory()      # (cols, rows) C-order == F-order (rows, cols)
            tpin.numpy()[...] = a.T
            pinned[k_] = tpin
            e2e_prob[k_] = tpin.numpy().T                                             # column-major view on pinned memory
    h2d = sum(int(np.asarray(v).nbytes) for k_, v in e2e_prob.items()
              if isinstance(v, np.ndarray) and k_ in ("G", "A", "P", "a", "ub", "b", "rvec", "xdef", "x0"))
    barrier()
    e2e_budget = K
    t0 = time.perf_counter()
    op = cb.from_dict(e2e_prob, "BR", cb.SolverParams(stepLimit=e2e_budget), h)
    sol = op.solve()
    xsum = float(np.sum(sol.x))                 # result read back on the host
    h.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_steps = sol.executed_newton_steps + sol.phase1_executed_steps
    d2h = sol.x.nbytes + 128 * 8
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    e2e_n = torch.tensor([float(e2e_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_n, op=dist.ReduceOp.SUM)
    e2e_value = e2e_n.item() / e2e_t.item()

    # ---- batched leg (BASELINE.json configs[2]): B = 8192 independent n=64, m=128 problems (half KL with
    # p=1, half QP with p=0), contiguous shards over the ranks, one CTA per problem, then the NCCL gather
    # of the solutions + convergence all-reduce.  Strong scaling: B is fixed as N grows.
    batched = None
    if not args.no_batched:
        import synthetic as P
        Bt = args.batch
        lo, hi = cb.shard_range(Bt, rank, world)
        bprobs = [P.batched_problem(i, 64, 128, 1000) for i in range(lo, hi)]
        solver = cb.BatchedBarrierSolver(cb.pack_problems(bprobs), cb.SolverParams(), h)
        wsol = solver.solve()                             # warm-up (same shapes)
        if world > 1:
            cb.gather_solutions(wsol, Bt, 64)             # and of the exchange: NCCL connects its all-gather rings on first use
        barrier()
        t0 = time.perf_counter()
        bsol = solver.solve()
        if world > 1:
            g = cb.gather_solutions(bsol, Bt, 64)
            conv, mxs = g["converged"], g["max_newton_steps"]
        else:
            conv, mxs = int((bsol.status == 0).sum()), int(bsol.newton_steps.max())
        torch.cuda.synchronize()
        bt = torch.tensor([time.perf_counter() - t0, bsol.solve_ms / 1e3], dtype=torch.float64, device="cuda")
        bn = torch.tensor([float(bsol.newton_steps.sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bn, op=dist.ReduceOp.SUM)
        wall_s, dev_s = bt.tolist()
        batched = {"metric": "batched_barrier_solves_per_sec", "value": Bt / wall_s, "unit": "solves/s", "B": Bt, "n": 64,
                   "m": 128, "scaling": "strong", "device_only_value": Bt / dev_s, "converged": conv,
                   "max_newton_steps": mxs, "newton_steps_per_sec": bn.item() / wall_s,
                   "includes": "kernel + device-to-host of the shard + NCCL all-gather of x and status + all-reduce"
                               if world > 1 else "kernel + device-to-host of the results"}

    if rank == 0:
        # ---- roofline of the dominant kernel ---------------------------------------------------------------
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_file)) if os.path.exists(peaks_file) else {}
        pk_ms, pk_fl = h.bench_kernel(0, 30000, 0, 1)
        fp64_peak = pk_fl / pk_ms / 1e9
        achieved = syrk_flops / (syrk_ms / 1e3) / 1e12 if syrk_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "gemm_dmma_streamk_kernel<TN> (Hessian assembly G' diag(w) G, FP64 DMMA m8n8k4, persistent stream-K grid)",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                    "traffic": 111.9e6 if args.workload == "c2" else None, "traffic_source": "dram__bytes_read+write per launch, ncu --set full, "
                    "profiles/r1_ncu_full_syrk_hessian_c2_n2000_m4000.txt (algorithmic: 64 MB read of G + 32 MB write of H; "
                    "the rest is the stream-K partial tiles, 148 x 128 KB written and read back, and re-reads of G that "
                    "miss L2 in ncu's cold-cache replay)" if args.workload == "c2" else None, "launches": n_syrk, "avg_launch_ms": syrk_ms / max(n_syrk, 1),
                    "flops_per_launch": syrk_flops / max(n_syrk, 1),
                    "peak_source": "FP64 DMMA issue-rate probe measured in this run (MEASURED_PEAKS.json has HBM and bf16 "
                                   "only: hbm_gbs=%s); cuBLAS DGEMM 8192^3 on this pool measured 35.5 TFLOP/s" % peaks.get("hbm_gbs"),
                    "share_of_step": syrk_ms / dev_ms if dev_ms > 0 else None}
        # the HBM-bound kernels of the step (slack / gradient / line-search GEMVs over G), timed alone on this handle
        hbm_peak = peaks.get("hbm_gbs")
        hbm = {}
        if args.workload == "c2":
            for which, name in ((7, "gemv_n (G x, G d)"), (8, "gemv_t (G' (1/d))")):
                ms_k, by_k = h.bench_kernel(which, 2000, 4000, 20)
                gbs = by_k / ms_k / 1e6
                hbm[name] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": gbs / hbm_peak if hbm_peak else None, "bytes_per_launch": by_k,
                             "note": "G is 64 MB: partly L2-resident between repeated launches"}
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count()
            cs = args.cpu_steps or (40 if args.workload == "c2" else 20)
            try:
                val, done, dt, sample = cpu_reference_leg(args.workload, cs, 2)
                cpu = {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
                       "sample": "%d %s, %.1f s of CPU time (numpy/LAPACK oracle; the Scala reference needs a JVM this image lacks)"
                                 % (done, sample, dt)}
            except Exception as e:      # never lose the GPU line to a CPU-side problem
                cpu = {"value": None, "unit": "steps/s", "cores": cores, "kind": "port", "sample": "failed: %r" % (e,)}
        last = st["last"]
        line = {"metric": "newton_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms_max / max(steps_done, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload + ": " + w["desc"], "n": n, "m": m, "p": p,
                           "parallelism": "replicas only (one independent problem stream per GPU, no data-path collective)",
                           "l2": "working set per step (G, scaled G, H, L, RHS: %.0f MB) exceeds the 126 MB L2; no flush needed"
                                 % ((2 * m * n + 3 * n * n + n * (p + 1)) * 8 / 1e6),
                           "solves_started": st["solves"],
                           "steps_counted": "%d phase-I Newton steps (dimension n+1, p = 0) + %d main-phase Newton steps "
                                            "(p equalities, Schur complement), each from the start of full solves" % (K1, K2)},
                "clocks": clk, "wall_ms_per_step": wall_ms_max / max(steps_done, 1),
                "flops_per_step": f_step(n, m, p), "tflops": f_step(n, m, p) * value / world / 1e12,
                "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d / max(e2e_steps, 1),
                        "d2h_bytes_per_step": d2h / max(e2e_steps, 1), "steps": e2e_steps, "seconds": e2e_s, "checksum": xsum},
                "gpu_launches": int(total_launches), "roofline": roofline, "hbm_kernels": hbm, "cpu_baseline": cpu, "batched": batched,
                "last_solution": {"objective": last.objective, "outer_stages": last.outer_stages,
                                  "newton_steps": last.newton_steps, "phase1_newton_steps": last.phase1_newton_steps}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Batched barrier solves of many small independent problems (BASELINE.json configs[2]: n = 64, m = 128,
p in {0,1}).  One CTA per problem on the device (cvxb_batch_barrier_solve);
across GPUs the batch is split into contiguous blocks, one block per rank, with no data-path collective:
the only communication is the final gather of the solutions and a convergence reduction.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import BatchDesc, BatchResult, check, dptr, OBJ_KL, OBJ_LINEAR, OBJ_QUADRATIC

_KIND = {"linear": OBJ_LINEAR, "quadratic": OBJ_QUADRATIC, "kl": OBJ_KL}


@dataclass
class BatchSolution:
    x: np.ndarray              # B x n
    status: np.ndarray         # B   cvxb_status per problem (0 = ok)
    newton_steps: np.ndarray   # B
    outer_stages: np.ndarray   # B
    objective: np.ndarray      # B
    dualityGap: np.ndarray     # B
    equalityGap: np.ndarray    # B
    solve_ms: float
    stage_newton_steps: Optional[np.ndarray] = None    # B x 16: Newton steps per outer stage
    cycles: Optional[np.ndarray] = None                # B: SM clock cycles each problem occupied its CTA
    phase1_newton_steps: Optional[np.ndarray] = None   # B: Newton steps / stages / final slack of the phase-I analysis
    phase1_stages: Optional[np.ndarray] = None
    phase1_s: Optional[np.ndarray] = None


def pack_problems(probs: Sequence[dict]):
    """Problem dictionaries (oracle/problems.py layout) -> packed column-major arrays of cvxb_batch_desc.
    Every problem must have the same n and m and at most one equality.  A problem without a strictly feasible x0 is
    started from its `xdef` (ConstraintSet.pointWhereDefined) and flagged for the phase-I analysis."""
    B = len(probs)
    n = probs[0]["n"]
    m = probs[0]["G"].shape[0]
    obj = np.zeros(B, dtype=np.int32)
    pcount = np.zeros(B, dtype=np.int32)
    obj_a = np.zeros((B, n))
    obj_r = np.zeros(B)
    any_quad = any(p["kind"] == "quadratic" for p in probs)
    obj_P = np.zeros((B, n, n)) if any_quad else None
    G = np.empty((B, n, m))          # [b, j, i] = G_b(i, j): column-major per problem
    ub = np.empty((B, m))
    A = np.zeros((B, n))
    b = np.zeros(B)
    x0 = np.empty((B, n))
    phase1 = np.zeros(B, dtype=np.int32)
    for k, pr in enumerate(probs):
        assert pr["n"] == n and pr["G"].shape == (m, n), "batched problems must share n and m"
        assert pr.get("x0") is not None or pr.get("xdef") is not None, "need a feasible start x0 or a point xdef"
        obj[k] = _KIND[pr["kind"]]
        if pr["kind"] != "kl":
            obj_a[k] = pr["a"]
        obj_r[k] = pr.get("r", 0.0) or 0.0
        if pr["kind"] == "quadratic":
            obj_P[k] = np.asarray(pr["P"]).T
        rv = pr.get("rvec")
        G[k] = np.asarray(pr["G"]).T
        ub[k] = pr["ub"] - (rv if rv is not None else 0.0)     # r + Gx <= ub  <=>  Gx <= ub - r
        if pr.get("A") is not None:
            assert pr["A"].shape[0] == 1, "batched solver supports p in {0, 1}"
            pcount[k] = 1
            A[k] = pr["A"][0]
            b[k] = pr["b"][0]
        if pr.get("x0") is not None:
            x0[k] = pr["x0"]
        else:
            x0[k] = pr["xdef"]
            phase1[k] = 1
    return dict(B=B, n=n, m=m, objective=obj, pcount=pcount, obj_a=obj_a, obj_r=obj_r, obj_P=obj_P, G=G, ub=ub, A=A, b=b, x0=x0,
                phase1=phase1 if phase1.any() else None)


class BatchedBarrierSolver:
    def __init__(self, packed: dict, pars=None, handle=None):
        from .solvers import SolverParams
        self.handle = handle if handle is not None else _lib.default_handle()
        self.pars = pars if pars is not None else SolverParams()
        self.packed = packed          # keeps the host arrays alive during the upload
        d = BatchDesc()
        d.B, d.n, d.m = packed["B"], packed["n"], packed["m"]
        d.p = int(packed["pcount"].max()) if packed["B"] else 0
        d.objective = packed["objective"].ctypes.data_as(C.POINTER(C.c_int))
        d.pcount = packed["pcount"].ctypes.data_as(C.POINTER(C.c_int))
        d.obj_a, d.obj_r = dptr(packed["obj_a"]), dptr(packed["obj_r"])
        d.obj_P = dptr(packed["obj_P"]) if packed["obj_P"] is not None else None
        d.G, d.ub, d.x0 = dptr(packed["G"]), dptr(packed["ub"]), dptr(packed["x0"])
        d.A, d.b = dptr(packed["A"]), dptr(packed["b"])
        ph = packed.get("phase1")
        d.phase1 = ph.ctypes.data_as(C.POINTER(C.c_int)) if ph is not None else None
        self.B, self.n = d.B, d.n
        self._b = C.c_void_p()
        check(self.handle.lib.cvxb_batch_create(self.handle._h, C.byref(d), C.byref(self._b)))

    def solve(self, download: bool = True) -> BatchSolution:
        """One launch of the batched barrier kernel.  download=False leaves every result on the device (the packed
        records of cvxb_batch_device_records, which `gather_solutions` ships with one all-gather)."""
        B, n = self.B, self.n
        r = BatchResult()
        cp = self.pars.to_c(self.handle)
        if not download:
            check(self.handle.lib.cvxb_batch_barrier_solve(self.handle._h, self._b, C.byref(cp), C.byref(r)))
            return BatchSolution(None, None, None, None, None, None, None, float(r.solve_ms))
        x = np.empty((B, n))
        status = np.empty(B, dtype=np.int32)
        steps = np.empty(B, dtype=np.int32)
        stages = np.empty(B, dtype=np.int32)
        objv, gap, eqg = np.empty(B), np.empty(B), np.empty(B)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        r.x, r.status, r.newton_steps, r.outer_stages = dptr(x), ip(status), ip(steps), ip(stages)
        r.objective, r.duality_gap, r.equality_gap = dptr(objv), dptr(gap), dptr(eqg)
        stage_steps = np.zeros((B, 16), dtype=np.int32)
        r.stage_newton_steps = ip(stage_steps)
        cycles = np.zeros(B, dtype=np.int64)
        r.cycles = cycles.ctypes.data_as(C.POINTER(C.c_longlong))
        ph_steps, ph_stages, ph_s = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32), np.zeros(B)
        r.phase1_newton_steps, r.phase1_stages, r.phase1_s = ip(ph_steps), ip(ph_stages), dptr(ph_s)
        check(self.handle.lib.cvxb_batch_barrier_solve(self.handle._h, self._b, C.byref(cp), C.byref(r)))
        return BatchSolution(x, status, steps, stages, objv, gap, eqg, float(r.solve_ms), stage_steps, cycles, ph_steps,
                             ph_stages, ph_s)

    def device_records(self):
        """(device pointer, doubles per row) of the packed results of the last solve:
        rows [x(n), objective, duality gap, status, newton steps, outer stages]."""
        p = C.c_void_p()
        row = C.c_int()
        check(self.handle.lib.cvxb_batch_device_records(self._b, C.byref(p), C.byref(row)))
        return int(p.value), int(row.value)

    def device_records_tensor(self):
        """The packed results as a torch CUDA tensor that aliases the library's buffer (no copy)."""
        import torch
        ptr_, row = self.device_records()

        class _View:
            __cuda_array_interface__ = {"shape": (self.B, row), "typestr": "<f8", "data": (ptr_, False), "version": 3,
                                        "strides": None}
        return torch.as_tensor(_View(), device=torch.device("cuda", self.handle.device))

    def close(self):
        if getattr(self, "_b", None):
            self.handle.lib.cvxb_batch_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_range(B: int, rank: int, world: int):
    """Contiguous block of ceil(B / world) problems per rank."""
    per = (B + world - 1) // world
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


RECORD_EXTRA = 5     # CVXB_BATCH_RECORD_EXTRA: objective, duality gap, status, newton steps, outer stages


def pack_records(local: BatchSolution) -> np.ndarray:
    """Host-side twin of the device record layout (used on the gloo path of the CPU tests)."""
    k, n = local.x.shape
    rec = np.empty((k, n + RECORD_EXTRA))
    rec[:, :n] = local.x
    rec[:, n] = local.objective
    rec[:, n + 1] = local.dualityGap
    rec[:, n + 2] = local.status
    rec[:, n + 3] = local.newton_steps
    rec[:, n + 4] = local.outer_stages
    return rec


_PINNED = {}


def _pinned_buffer(rows: int, cols: int):
    """Pinned host staging buffer for the gathered records, kept between calls (pinning memory costs more than the copy)."""
    import torch
    key = (rows, cols)
    if key not in _PINNED:
        _PINNED.clear()
        _PINNED[key] = torch.empty(rows, cols, dtype=torch.float64).pin_memory()
    return _PINNED[key]


def gather_solutions(local, B: int, n: int, group=None):
    """Final exchange of the sharded batch: ONE all-gather of the packed per-problem records
    [x(n), objective, gap, status, newton steps, outer stages]; the global convergence figures (converged count, max
    Newton steps) are computed from the gathered status / step columns, so no separate reduction is needed.
    `local` is a BatchedBarrierSolver whose results are still on the device (NCCL over NVLink: the library's record
    buffer is handed to the all-gather as is, then one device-to-host copy of the gathered block) or a BatchSolution
    on the host (gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (B + world - 1) // world
    row = n + RECORD_EXTRA
    nccl = dist.get_backend(group) == "nccl"
    if isinstance(local, BatchedBarrierSolver):
        rec = local.device_records_tensor()
        dev = rec.device
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
        rec = torch.from_numpy(pack_records(local)).to(dev) if local.x.shape[0] else torch.zeros(0, row, dtype=torch.float64, device=dev)
    k = rec.shape[0]
    if k != per:          # the last shard may be short: pad with status -1 rows
        pad = torch.zeros(per, row, dtype=torch.float64, device=dev)
        pad[:, n + 2] = -1.0
        pad[:k] = rec
        rec = pad
    out = torch.empty(world * per, row, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    if world * per == B:
        # even split (the BASELINE shapes): the gathered block IS the result, one copy into a pinned host buffer
        if dev.type == "cuda":
            host = _pinned_buffer(world * per, row)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            g = host.numpy()
        else:
            g = out.numpy()
    else:
        # drop the padding rows of short shards, keep problem order
        keep = torch.cat([torch.arange(r_ * per, r_ * per + (shard_range(B, r_, world)[1] - shard_range(B, r_, world)[0]))
                          for r_ in range(world)]).to(dev)
        g = out.index_select(0, keep).cpu().numpy()
    status = g[:, n + 2].astype(np.int32)
    steps = g[:, n + 3].astype(np.int32)
    return dict(x=np.ascontiguousarray(g[:, :n]), objective=g[:, n].copy(), dualityGap=g[:, n + 1].copy(), status=status,
                newton_steps=steps, outer_stages=g[:, n + 4].astype(np.int32), converged=int((status == 0).sum()),
                max_newton_steps=int(steps.max()) if steps.size else 0)

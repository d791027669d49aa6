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


def pack_problems(probs: Sequence[dict]):
    """Problem dictionaries (oracle/problems.py layout) -> packed column-major arrays of cvxb_batch_desc.
    Every problem must have the same n and m, a strictly feasible x0 and at most one equality."""
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
    for k, pr in enumerate(probs):
        assert pr["n"] == n and pr["G"].shape == (m, n), "batched problems must share n and m"
        assert pr.get("x0") is not None, "batched solver needs a strictly feasible start (no phase I)"
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
        x0[k] = pr["x0"]
    return dict(B=B, n=n, m=m, objective=obj, pcount=pcount, obj_a=obj_a, obj_r=obj_r, obj_P=obj_P, G=G, ub=ub, A=A, b=b, x0=x0)


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
        self.B, self.n = d.B, d.n
        self._b = C.c_void_p()
        check(self.handle.lib.cvxb_batch_create(self.handle._h, C.byref(d), C.byref(self._b)))

    def solve(self) -> BatchSolution:
        B, n = self.B, self.n
        x = np.empty((B, n))
        status = np.empty(B, dtype=np.int32)
        steps = np.empty(B, dtype=np.int32)
        stages = np.empty(B, dtype=np.int32)
        objv, gap, eqg = np.empty(B), np.empty(B), np.empty(B)
        r = BatchResult()
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        r.x, r.status, r.newton_steps, r.outer_stages = dptr(x), ip(status), ip(steps), ip(stages)
        r.objective, r.duality_gap, r.equality_gap = dptr(objv), dptr(gap), dptr(eqg)
        cp = self.pars.to_c(self.handle)
        check(self.handle.lib.cvxb_batch_barrier_solve(self.handle._h, self._b, C.byref(cp), C.byref(r)))
        return BatchSolution(x, status, steps, stages, objv, gap, eqg, float(r.solve_ms))

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


def gather_solutions(local: BatchSolution, B: int, n: int, group=None):
    """Final exchange of the sharded batch: all-gather of x (B x n doubles) and of the per-problem
    status / step counts, plus one all-reduce of (converged count, max Newton steps).  Works on any
    torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (B + world - 1) // world
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    xbuf = torch.zeros(per, n, dtype=torch.float64, device=dev)
    ibuf = torch.zeros(per, 3, dtype=torch.int32, device=dev)
    k = local.x.shape[0]
    if k:
        xbuf[:k] = torch.from_numpy(local.x).to(dev)
        ibuf[:k, 0] = torch.from_numpy(local.status).to(dev)
        ibuf[:k, 1] = torch.from_numpy(local.newton_steps).to(dev)
        ibuf[:k, 2] = torch.from_numpy(local.outer_stages).to(dev)
    xs = torch.empty(world * per, n, dtype=torch.float64, device=dev)
    is_ = torch.empty(world * per, 3, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(xs, xbuf, group=group)
    dist.all_gather_into_tensor(is_, ibuf, group=group)
    conv = torch.tensor([int((local.status == 0).sum())], dtype=torch.int64, device=dev)
    mx = torch.tensor([int(local.newton_steps.max()) if k else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(conv, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    xs, is_ = xs[:B].cpu().numpy(), is_[:B].cpu().numpy()
    return dict(x=xs, status=is_[:, 0], newton_steps=is_[:, 1], outer_stages=is_[:, 2], converged=int(conv.item()),
                max_newton_steps=int(mx.item()))

"""The multi-GPU exchange of the batched path (contiguous sharding, ONE all-gather of the packed per-problem records,
convergence figures derived from the gathered status / step columns) on the gloo backend with world_size 2 -- host
logic only, no GPU."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cvx_b200 as cb
    lo, hi = cb.shard_range(B, rank, world)
    k = hi - lo
    # stand-in for the device result of this rank's block: x_b = b (problem index), steps = 10 + b
    idx = np.arange(lo, hi)
    local = cb.BatchSolution(x=np.repeat(idx[:, None].astype(float), n, 1), status=np.where(idx == 3, 1, 0).astype(np.int32),
                             newton_steps=(10 + idx).astype(np.int32), outer_stages=np.full(k, 12, dtype=np.int32),
                             objective=np.zeros(k), dualityGap=np.zeros(k), equalityGap=np.zeros(k), solve_ms=1.0)
    out = cb.gather_solutions(local, B, n)
    assert out["x"].shape == (B, n) and out["outer_stages"].tolist() == [12] * B
    q.put((rank, out["x"][:, 0].tolist(), out["status"].tolist(), out["newton_steps"].tolist(), out["converged"],
           out["max_newton_steps"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8])
def test_gather_world2(B):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, x0, status, steps, conv, mx in res:
        assert x0 == [float(i) for i in range(B)]
        assert status == [1 if i == 3 else 0 for i in range(B)]
        assert steps == [10 + i for i in range(B)]
        assert conv == B - 1 and mx == 10 + B - 1

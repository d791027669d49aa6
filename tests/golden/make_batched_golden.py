#!/usr/bin/env python
"""Golden results of the CPU oracle for a 256-problem sample of the batched config (BASELINE.json configs[2]:
B = 8192 problems n = 64, m = 128, even index KL with p = 1, odd index QP with p = 0; synthetic.batched_problem(i, 64,
128, 1000)).  The GPU test tests/test_batched_gpu.py::test_batched_8192_against_golden_sample runs all 8192 problems on
the device and compares the sampled ones with this file; tests/test_oracle_cpu.py re-derives a few entries live.

usage: python tests/golden/make_batched_golden.py     (about a minute on 8 cores; writes batched_8192_sample.npz)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cvx_oracle as O      # noqa: E402
from oracle import problems as P        # noqa: E402

B, N, M, BASE_SEED, SAMPLE = 8192, 64, 128, 1000, 256


def sample_indices():
    rng = np.random.default_rng(20261019)
    idx = np.sort(rng.choice(B, SAMPLE - 2, replace=False))
    return np.unique(np.concatenate([[0, B - 1], idx]))          # both ends of the batch are always in


def main():
    from threadpoolctl import threadpool_limits
    idx = sample_indices()
    x = np.zeros((len(idx), N))
    obj = np.zeros(len(idx))
    gap = np.zeros(len(idx))
    stages = np.zeros(len(idx), dtype=np.int32)
    stage_steps = np.zeros((len(idx), 16), dtype=np.int32)
    with threadpool_limits(limits=1):
        for k, i in enumerate(idx):
            pr = P.batched_problem(int(i), N, M, BASE_SEED)
            objF, cnts, eqs = P.to_oracle(pr)
            sol, _ = O.solveProblem(objF, cnts, eqs, "BR")
            x[k] = sol.x
            obj[k] = objF.valueAt(sol.x)
            gap[k] = sol.dualityGap
            stages[k] = sol.outer_stages
            ss = sol.stage_newton_steps[:16]
            stage_steps[k, :len(ss)] = ss
            if k % 32 == 0:
                print(k, i, obj[k], ss, flush=True)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "batched_8192_sample.npz")
    np.savez_compressed(out, index=idx.astype(np.int32), x=x, objective=obj, dualityGap=gap, outer_stages=stages,
                        stage_newton_steps=stage_steps)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

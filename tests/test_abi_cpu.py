"""The C-ABI library loads without a GPU, exports every symbol include/cvxb.h declares, its POD
structs have the layout the ctypes binding assumes, and the product path refuses to run without CUDA."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cvxb.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvxb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cvx_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libcvxb.so does not export %s" % n
        assert n in _lib.SYMBOLS, "ctypes binding lacks %s" % n
    assert set(_lib.SYMBOLS) <= set(names)


def test_struct_layouts_match_the_header(tmp_path):
    from cvx_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cvxb.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(cvxb_params),sizeof(cvxb_kkt_info),sizeof(cvxb_problem_desc),sizeof(cvxb_solution),'
                    'sizeof(cvxb_batch_desc),sizeof(cvxb_batch_result),offsetof(cvxb_solution,stage_newton_steps),'
                    'offsetof(cvxb_solution,solve_ms),offsetof(cvxb_params,stepLimit));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.Params), C.sizeof(_lib.KktInfo), C.sizeof(_lib.ProblemDesc), C.sizeof(_lib.SolutionC),
            C.sizeof(_lib.BatchDesc), C.sizeof(_lib.BatchResult), _lib.SolutionC.stage_newton_steps.offset,
            _lib.SolutionC.solve_ms.offset, _lib.Params.stepLimit.offset]
    assert got == want


def test_default_params_are_the_reference_constants():
    from cvx_b200 import _lib
    p = _lib.Params()
    assert _lib.load().cvxb_default_params(C.byref(p)) == 0
    # SolverParams.standardParams (SolverParams.scala:35-46) and the constants hard-coded in the solvers
    assert (p.maxIter, p.alpha, p.beta, p.tolSolver, p.tolEqSolve, p.tolFeas, p.delta) == (1000, .04, .8, 1e-8, 1e-1, 1e-7, 1e-6)
    assert (p.mu, p.t0, p.ruizMaxSweeps, p.ruizTol, p.cholRegDelta, p.cholMinDiag) == (10.0, 1.0, 20, 1e-6, 1e-10, 1e-7)
    assert (p.newtonRegDelta, p.phase1EqTol, p.pdStepFraction, p.bugCompat, p.stepLimit) == (1e-9, 1e-6, 0.99, 0, 0)


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cvx_b200 import _lib
    with pytest.raises(_lib.CudaError):
        _lib.Handle(0)
    import cvx_b200 as cb
    with pytest.raises(_lib.CudaError):
        cb.KKTSystem(np.eye(2), np.ones((1, 2)), np.ones(2), np.ones(1))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cvx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_constructor_dimension_asserts_do_not_need_a_gpu():
    import cvx_b200 as cb
    with pytest.raises(AssertionError):
        cb.KKTSystem(np.eye(3), np.ones((1, 4)), np.ones(3), np.ones(1), handle=object())
    with pytest.raises(AssertionError):
        cb.KKTSystem(np.ones((3, 2)), np.ones((1, 2)), np.ones(2), np.ones(1), handle=object())


def test_pack_problems_layout():
    import cvx_b200 as cb
    from oracle import problems as P
    probs = [P.batched_problem(i, 8, 16, 5) for i in range(4)]
    pk = cb.pack_problems(probs)
    assert pk["G"].shape == (4, 8, 16) and pk["pcount"].tolist() == [1, 0, 1, 0]
    # column-major per problem: element (i, j) of G_b at flat index j*m + i
    flat = pk["G"][1].reshape(-1)
    assert flat[3 * 16 + 5] == probs[1]["G"][5, 3]
    assert np.array_equal(pk["obj_P"][1], probs[1]["P"].T)
    assert cb.shard_range(10, 0, 4) == (0, 3) and cb.shard_range(10, 3, 4) == (9, 10) and cb.shard_range(2, 3, 4) == (2, 2)


def test_pack_problems_flags_phase1_for_problems_without_a_feasible_start():
    """A problem that only has a point where it is defined (ConstraintSet.pointWhereDefined) is started there and flagged
    for the in-kernel phase-I analysis; a batch of feasible starts carries no flag array at all."""
    import cvx_b200 as cb
    from oracle import problems as P
    probs = [P.batched_problem_phase1(i, 7, 14, 3) for i in range(3)] + [P.slab_qp(7, 7, 0, 9)]
    pk = cb.pack_problems(probs)
    assert pk["phase1"].tolist() == [1, 1, 1, 0]
    assert np.array_equal(pk["x0"][0], probs[0]["xdef"]) and np.array_equal(pk["x0"][3], probs[3]["x0"])
    assert cb.pack_problems([P.slab_qp(7, 7, 0, 9)])["phase1"] is None
    from cvx_b200 import _lib
    assert _lib.BatchDesc.phase1.offset == C.sizeof(_lib.BatchDesc) - C.sizeof(C.c_void_p)      # appended: old callers stay valid


def test_tile_dag_block_schedule():
    """Host logic of the tile-DAG factorisation schedule (factor.cu: dag_blocks): the diagonal blocks tile [0, n) in order,
    every start is a multiple of the 128-column leaf, the first block is half a block when there are at least three
    (nothing can overlap the first chain), the blocks shrink geometrically at the end down to <= 5 leaves, and no block is
    wider than the nominal width (or than the 5-leaf tail).  C4 (n = 8192, 2048): 1024, 2048, 2048, 1536, 768, 384, 384."""
    from cvx_b200 import _lib
    lib = _lib.load()
    buf = (C.c_int * 256)()

    def blocks(n, nbk):
        k = lib.cvxb_debug_dag_blocks(n, nbk, buf, 256)
        assert 2 <= k <= 256
        st = list(buf[:k])
        assert st[0] == 0 and st[-1] == n and all(b > a for a, b in zip(st, st[1:]))
        assert all(s_ % 128 == 0 for s_ in st[:-1])
        return [b - a for a, b in zip(st, st[1:])]

    assert blocks(8192, 2048) == [1024, 2048, 2048, 1536, 768, 384, 384]
    for n, nbk in [(5120, 2048), (5121, 2048), (16385, 2048), (16385, 1024), (12289, 2048), (1000, 256), (1537, 384), (2049, 512),
                   (32768, 2048), (6145, 1024)]:
        sz = blocks(n, nbk)
        assert sum(sz) == n and max(sz) <= max(nbk, 5 * 128 + 1) and sz[-1] <= 5 * 128 + 1      # (a tail of <= 5 leaves is one block)
        if n >= 3 * nbk:
            assert sz[0] == max(128, nbk // 2 // 128 * 128)
        if nbk >= 1024:                                              # (with test-sized blocks the 5-leaf tail exceeds a block)
            tail = sz[1:]
            assert all(b <= a for a, b in zip(tail, tail[1:]))      # non-increasing after the first block
    assert lib.cvxb_debug_dag_blocks(0, 2048, buf, 256) == -1 and lib.cvxb_debug_dag_blocks(100, 64, buf, 256) == -1


def test_bench_gpu_arm_never_imports_the_oracle():
    """bench.py may execute oracle/ only in its CPU legs (cpu_baseline, --impl reference): every import of it sits
    inside cpu_reference_leg / cpu_modes_leg; synthetic.py (input generation for both arms) does not import it at all."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())

    def oracle_imports(node):
        out = []
        for sub in ast.walk(node):
            if isinstance(sub, ast.ImportFrom) and (sub.module or "").split(".")[0] == "oracle":
                out.append(sub.lineno)
            if isinstance(sub, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in sub.names):
                out.append(sub.lineno)
        return out

    allowed = []
    for node in tree.body:       # the two CPU legs: the reference arm / cpu_baseline, and the literal-vs-vectorised modes
        if isinstance(node, ast.FunctionDef) and node.name in ("cpu_reference_leg", "cpu_modes_leg"):
            allowed += oracle_imports(node)
    assert allowed, "the CPU legs should be the only places that import oracle/"
    assert sorted(oracle_imports(tree)) == sorted(allowed)
    syn = ast.parse(open(os.path.join(root, "synthetic.py")).read())
    assert oracle_imports(syn) == []
    for name in ("gpu_big.py", "gpu_batch.py", "gpu_syrk.py", "gpu_potrf.py", "gpu_gemm_sweep.py"):
        assert oracle_imports(ast.parse(open(os.path.join(root, "tools", name)).read())) == []

"""Seam B (per-step dense linear algebra) through the C ABI against the CPU oracle.  The test designs
are the reference's own (src/test/scala/cvx/KktTest.scala, MatrixUtilsTests.scala) with seeds added."""
import numpy as np
import pytest

from oracle import cvx_oracle as O
from oracle import problems as P

pytestmark = pytest.mark.gpu
RTOL = 1e-10     # BASELINE.json north_star: dx, nu within 1e-10 relative


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def spd(n, seed, cond_pow=0):
    rng = np.random.default_rng(seed)
    M = rng.uniform(-1, 1, (n, n))
    H = M @ M.T + n * 1e-3 * np.eye(n)
    s = 10.0 ** rng.uniform(-cond_pow, cond_pow, n)       # bad row/column scaling, Ruiz must undo it
    H = H * np.outer(s, s)
    return (H + H.T) * 0.5


def test_ruiz_zero_row(handle):
    """MatrixUtilsTests.scala:16-26: a zero row keeps d_i = 1 (MatrixUtils.scala:259)."""
    from cvx_b200 import MatrixUtils
    A = np.array([[.5, 0, .5], [0, 0, 0], [-.5, 0, -.5]])
    d, Q, sweeps = MatrixUtils.ruizEquilibrate(A, handle, return_sweeps=True)
    d0, Q0 = O.ruizEquilibrate(A)
    assert d[1] == 1.0
    assert sweeps == O.ruizEquilibrate.last_sweeps
    assert np.allclose(d, d0, rtol=1e-14, atol=0) and np.allclose(Q, Q0, rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("n,cond_pow", [(50, 0), (257, 2), (1000, 3)])
def test_ruiz_matches_oracle(handle, n, cond_pow):
    from cvx_b200 import MatrixUtils
    H = spd(n, n, cond_pow)
    d, Q, sweeps = MatrixUtils.ruizEquilibrate(H, handle, return_sweeps=True)
    d0, Q0 = O.ruizEquilibrate(H)
    assert abs(sweeps - O.ruizEquilibrate.last_sweeps) <= 1
    if sweeps == O.ruizEquilibrate.last_sweeps:
        assert rel(d, d0) < 1e-12 and rel(Q, Q0) < 1e-12
    assert np.array_equal(Q, Q.T)


@pytest.mark.parametrize("n,cond_pow", [(4096, 1), (4229, 2)])
def test_ruiz_symmetric_half_sweeps(handle, n, cond_pow, monkeypatch):
    """Beyond L2 (n >= 4096) the sweeps read only the lower triangle (ruiz_sym_kernel: tile-wise column and row partials,
    deterministic order): same d, same scaled matrix and the same number of sweeps as the oracle (MatrixUtils.scala:240-268)
    and as the full-column kernel, with a ragged last tile (4229 = 33 * 128 + 5) and a zero row."""
    from cvx_b200 import MatrixUtils
    rng = np.random.default_rng(n)
    M = rng.uniform(-1, 1, (n, 64))
    H = M @ M.T + 1e-3 * n * np.eye(n)
    sc = 10.0 ** rng.uniform(-cond_pow, cond_pow, n)
    H = H * np.outer(sc, sc)
    H = (H + H.T) * 0.5
    H[7, :] = 0.0
    H[:, 7] = 0.0
    d, Q, sweeps = MatrixUtils.ruizEquilibrate(H, handle, return_sweeps=True)
    d0, Q0 = O.ruizEquilibrate(H)
    assert d[7] == 1.0
    assert abs(sweeps - O.ruizEquilibrate.last_sweeps) <= 1
    if sweeps == O.ruizEquilibrate.last_sweeps:
        assert rel(d, d0) < 1e-12 and rel(Q, Q0) < 1e-12
    assert np.array_equal(Q, Q.T)
    d_a, _, sweeps_a = MatrixUtils.ruizEquilibrate(H, handle, return_sweeps=True)
    assert np.array_equal(d, d_a) and sweeps == sweeps_a            # deterministic: no floating-point atomics


@pytest.mark.parametrize("n", [1, 7, 128, 129, 300, 1025])
def test_regularized_cholesky(handle, n):
    from cvx_b200 import MatrixUtils
    H = spd(n, 100 + n)
    L = MatrixUtils.regularizedCholesky(H, handle)
    L0 = O.regularizedCholesky(H)
    assert np.array_equal(np.triu(L, 1), np.zeros((n, n)))
    assert rel(L, L0) < 1e-11
    assert rel(L @ L.T, H) < 1e-13


def test_regularized_cholesky_semidefinite(handle):
    """rank-deficient Q: plain dpotrf fails or min diag <= 1e-7 -> Q + 1e-10 I (MatrixUtils.scala:452-461)."""
    from cvx_b200 import MatrixUtils, _lib
    rng = np.random.default_rng(3)
    B = rng.uniform(-1, 1, (40, 10))
    Q = B @ B.T
    Q = (Q + Q.T) / 2
    info = _lib.KktInfo()
    try:
        L = MatrixUtils.regularizedCholesky(Q, handle, info)
        ok = True
    except _lib.LinSolveException:
        ok = False
    try:
        L0 = O.regularizedCholesky(Q)
        ok0 = True
    except Exception:
        ok0 = False
    assert ok == ok0
    if ok:
        assert info.regularized == 1 and O.regularizedCholesky.last_regularized
        assert rel(L @ L.T, Q + 1e-10 * np.eye(40)) < 1e-9


@pytest.mark.parametrize("n,nrhs", [(5, 1), (200, 3), (513, 70), (1000, 1)])
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_triangular_solve_planted(handle, n, nrhs, uplo):
    """MatrixUtilsTests.testTriangularSolve (:36-93): L = tril(U(-5,5)) + 20 I, planted X."""
    from cvx_b200 import MatrixUtils
    rng = np.random.default_rng(n + nrhs)
    T = np.tril(rng.uniform(-5, 5, (n, n))) + 20 * np.eye(n)
    if uplo == "U":
        T = T.T.copy()
    X = rng.uniform(0, 1, (n, nrhs))
    B = T @ X
    X1 = MatrixUtils.triangularSolve(T, uplo, B, handle)
    X0 = O.triangularSolve(T, uplo, B)
    assert rel(X1, X0) < 1e-10 and rel(X1, X) < 1e-10
    if nrhs == 1:
        f = MatrixUtils.forwardSolve if uplo == "L" else MatrixUtils.backSolve
        assert rel(f(T, B[:, 0], handle), X[:, 0]) < 1e-10


def test_triangular_solve_singular(handle):
    from cvx_b200 import MatrixUtils, _lib
    T = np.tril(np.ones((10, 10)))
    T[4, 4] = 0.0
    with pytest.raises(_lib.LinSolveException):
        MatrixUtils.triangularSolve(T, "L", np.ones(10), handle)


@pytest.mark.parametrize("n", [3, 100, 700])
def test_cholesky_solve(handle, n):
    """MatrixUtilsTests.testSolveWithPreconditioning (:165-198)."""
    from cvx_b200 import MatrixUtils, _lib
    H = spd(n, 7 * n, 2)
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, n)
    b = H @ x
    info = _lib.KktInfo()
    x1 = MatrixUtils.choleskySolve(H, b, None, 1e-9, 0, handle, info)
    x0 = O.choleskySolve(H, b, 1e-9)
    assert np.linalg.norm(H @ x1 - b) / np.linalg.norm(b) < 1e-10
    assert rel(x1, x0) < 1e-7      # forward error is cond-limited; the residual above is the contract


def test_cholesky_solve_not_pd(handle):
    from cvx_b200 import MatrixUtils, _lib
    H = np.diag([1.0, -1.0, 2.0])
    with pytest.raises(_lib.LinSolveException):
        MatrixUtils.choleskySolve(H, np.ones(3), None, 1e-1, 0, handle)
    with pytest.raises(Exception):
        O.choleskySolve(H, np.ones(3), 1e-1)


@pytest.mark.parametrize("n,p,seed", [(10, 2, 0), (100, 20, 1), (300, 40, 2), (1000, 100, 3), (513, 129, 4)])
def test_kkt_planted_pd(handle, n, p, seed):
    """KktTest.testPositiveDefinite (:197-272): planted (x,w); intended size n=1000, p=100, tol 1e-10
    (Runner.scala:79-80)."""
    from cvx_b200 import KKTSystem
    s = P.kkt_planted_pd(n, p, seed)
    K = KKTSystem(s["H"], s["A"], s["q"], s["b"], handle)
    tol = 1e-7          # the reference's acceptance tolerance inside solveWithCholFactor (not the parity bar)
    x, w = K.solve(1e-6, None, tol, 0)
    info0 = O.KKTInfo()
    x0, w0 = O.kkt_solve(s["H"], s["A"], s["q"], s["b"], tol, info0)
    assert K.info.path == info0.path == 0
    assert abs(K.info.ruiz_sweeps - info0.ruiz_sweeps) <= 1
    H, A = s["H"], s["A"]

    def backward(xx, ww):
        res = np.linalg.norm(np.concatenate([H @ xx + A.T @ ww + s["q"], A @ xx - s["b"]]))
        return res / np.linalg.norm(np.concatenate([s["q"], s["b"]]))

    # parity bar: relative residual 1e-10 (or as good as the LAPACK oracle where cond(H) makes 1e-10 unreachable)
    assert backward(x, w) < max(RTOL, 10 * backward(x0, w0))
    # forward error against the planted solution no worse than the oracle's (KktTest :170-182)
    assert rel(x, s["x"]) < 10 * rel(x0, s["x"]) + 1e-12
    assert rel(w, s["w"]) < 10 * rel(w0, s["w"]) + 1e-12


@pytest.mark.parametrize("n,p,seed", [(300, 40, 5), (1000, 100, 6), (2304, 300, 7), (4800, 130, 8)])
def test_kkt_fused_forward_substitution(handle, n, p, seed, monkeypatch):
    """The forward substitutions Y = L^-1 [DA', Dq], K^-1 z that ride along with the Cholesky (potrf_lower_rhs: updates
    on the look-ahead stream, recursive split above the look-ahead limit) give the same solution as separate triangular
    sweeps after the factorisation (CVXB_NO_FUSED_TRSM), to rounding; both meet the planted-solution bar."""
    from cvx_b200 import KKTSystem, MatrixUtils
    s = P.kkt_planted_pd(n, p, seed)
    K = KKTSystem(s["H"], s["A"], s["q"], s["b"], handle)
    x, w = K.solve(1e-6, None, 1e-7, 0)
    c = MatrixUtils.choleskySolve(s["H"], s["q"], None, 1e-7, 0, handle)
    monkeypatch.setenv("CVXB_NO_FUSED_TRSM", "1")
    x1, w1 = KKTSystem(s["H"], s["A"], s["q"], s["b"], handle).solve(1e-6, None, 1e-7, 0)
    c1 = MatrixUtils.choleskySolve(s["H"], s["q"], None, 1e-7, 0, handle)
    monkeypatch.delenv("CVXB_NO_FUSED_TRSM")
    assert rel(x, x1) < 1e-7 and rel(w, w1) < 1e-7 and rel(c, c1) < 1e-7
    assert rel(x, s["x"]) < 1e-5 and rel(w, s["w"]) < 1e-5
    assert rel(s["H"] @ c, s["q"]) < 1e-9


@pytest.mark.parametrize("n,p,block,seed", [(1000, 100, 256, 11), (1537, 200, 384, 12), (2049, 129, 512, 13),
                                            (1300, 0, 256, 14)])
def test_tile_dag_schedule_equals_recursive(handle, n, p, block, seed):
    """The tile-DAG schedule of the big factorisations (potrf_dag: chain of diagonal-block factorisations beside the bulk
    updates on a third stream, persistent GEMM grids on a restricted number of SMs) is the same right-looking blocked
    Cholesky in another order: factor, KKT solution and choleskySolve agree with the recursive schedule to rounding and
    meet the planted-solution bar.  Small blocks bring the schedule (default: n >= 5120, blocks of 2048) down to test
    sizes, with ragged last blocks (1537 = 4*384 + 1, 2049 = 4*512 + 1)."""
    from cvx_b200 import KKTSystem, MatrixUtils
    H = spd(n, seed)
    try:
        handle.set_schedule(0, -1, -1)
        L0 = MatrixUtils.regularizedCholesky(H, handle)
        handle.set_schedule(block, block + 1, 8)
        L1 = MatrixUtils.regularizedCholesky(H, handle)
        assert np.array_equal(np.triu(L1, 1), np.zeros((n, n)))
        assert rel(L1, L0) < 1e-12 and rel(L1 @ L1.T, H) < 1e-13
        if p:
            s = P.kkt_planted_pd(n, p, seed)
            handle.set_schedule(0, -1, -1)
            x0, w0 = KKTSystem(s["H"], s["A"], s["q"], s["b"], handle).solve(1e-6, None, 1e-7, 0)
            c0 = MatrixUtils.choleskySolve(s["H"], s["q"], None, 1e-7, 0, handle)
            handle.set_schedule(block, block + 1, 8)
            x1, w1 = KKTSystem(s["H"], s["A"], s["q"], s["b"], handle).solve(1e-6, None, 1e-7, 0)
            c1 = MatrixUtils.choleskySolve(s["H"], s["q"], None, 1e-7, 0, handle)
            assert rel(x1, x0) < 1e-7 and rel(w1, w0) < 1e-7 and rel(c1, c0) < 1e-7
            assert rel(x1, s["x"]) < 1e-5 and rel(w1, s["w"]) < 1e-5
            assert rel(s["H"] @ c1, s["q"]) < 1e-9
    finally:
        handle.set_schedule()          # back to the defaults


@pytest.mark.parametrize("n,p,seed", [(50, 5, 0), (400, 60, 1)])
def test_solve_with_chol_factor(handle, n, p, seed):
    """KktTest.testSolutionWithCholFactor (:117-184)."""
    from cvx_b200 import KKTSystem
    s = P.kkt_planted_chol(n, p, seed)
    x, w = KKTSystem.solveWithCholFactor(s["L"], s["A"], s["q"], s["b"], None, 1e-10, 0, handle)
    x0, w0 = O.solveWithCholFactor(s["L"], s["A"], s["q"], s["b"], 1e-10)
    assert rel(x, s["x"]) < 1e-9 and rel(w, s["w"]) < 1e-9
    assert rel(x, x0) < 1e-8 and rel(w, w0) < 1e-8


def test_kkt_fallback_path1(handle):
    """H singular on null(A)-complement directions but H + A'A positive definite: path 0 fails,
    path 1 (KKTSystem.scala:57-59) succeeds -- same path in the oracle."""
    from cvx_b200 import KKTSystem
    rng = np.random.default_rng(11)
    n, p = 60, 10
    A = rng.uniform(-1, 1, (p, n))
    # H = projector-like PSD matrix vanishing on row space of A
    Qm, _ = np.linalg.qr(A.T, mode="complete")
    N = Qm[:, p:]
    H = N @ N.T
    H = (H + H.T) / 2
    x = rng.uniform(-1, 1, n)
    w = rng.uniform(-1, 1, p)
    q = -(H @ x + A.T @ w)
    b = A @ x
    info0 = O.KKTInfo()
    x0, w0 = O.kkt_solve(H, A, q, b, 1e-6, info0)
    K = KKTSystem(H, A, q, b, handle)
    x1, w1 = K.solve(1e-6, None, 1e-6, 0)
    assert K.info.path == info0.path
    assert rel(x1, x0) < 1e-6 and rel(w1, w0) < 1e-6


def test_symmetric_linear_system(handle):
    from cvx_b200 import SymmetricLinearSystem
    H = spd(200, 42, 2)
    rng = np.random.default_rng(1)
    r = rng.uniform(-1, 1, 200)
    x1 = SymmetricLinearSystem(H, r, None, handle).solve(1e-9, 0)
    x0 = O.symmetricLinearSystemSolve(H, r, 1e-9)
    assert np.linalg.norm(H @ x1 - r) / np.linalg.norm(r) < 1e-10
    assert rel(x1, x0) < 1e-7


def test_dimension_asserts(handle):
    from cvx_b200 import KKTSystem, _lib
    with pytest.raises(AssertionError):
        KKTSystem(np.eye(3), np.ones((1, 4)), np.ones(3), np.ones(1), handle)


# ---- decomposition fallbacks (kktSymSolve / symSolve / svdSolve, MatrixUtils.scala:603-751) ------------------

def test_kkt_fallback_path2_eigen(handle):
    """H negative definite: Cholesky fails on H and on H + A'A, the KKT matrix itself is nonsingular ->
    KKTSystem.kktSymSolve (KKTSystem.scala:63, 283-310)."""
    from cvx_b200 import KKTSystem
    rng = np.random.default_rng(21)
    n, p = 12, 3
    H = -np.eye(n) * 50.0
    A = rng.uniform(-1, 1, (p, n))
    x, w = rng.uniform(-1, 1, n), rng.uniform(-1, 1, p)
    q, b = -(H @ x + A.T @ w), A @ x
    info0 = O.KKTInfo()
    x0, w0 = O.kkt_solve(H, A, q, b, 1e-8, info0)
    K = KKTSystem(H, A, q, b, handle)
    x1, w1 = K.solve(1e-6, None, 1e-8, 0)
    assert info0.path == 2 and K.info.path == 2
    assert rel(x1, x) < 1e-9 and rel(w1, w) < 1e-9
    assert rel(x1, x0) < 1e-9 and rel(w1, w0) < 1e-9


@pytest.mark.parametrize("n,p,seed", [(103, 20, 1003), (164, 35, 1006), (119, 18, 1008), (40, 6, 5)])
def test_kkt_dependent_equality_rows(handle, n, p, seed):
    """Two identical rows in A make the Schur complement -- and the KKT matrix -- singular, the system stays consistent.
    Wherever the fallback chain ends (a Cholesky that gets through on a pivot of rounding noise, or kktSymSolve,
    KKTSystem.scala:283-310, on the singular KKT matrix), the reference answers; so does the device: the decomposition
    solve takes its coefficients from the orthonormal V and solves eigenvalues at rounding level as zeros (found by
    tools/gpu_fuzz_kkt.py: the left vectors w_j / s_j of a one-sided Jacobi SVD are noise for a zero singular value)."""
    from cvx_b200 import KKTSystem
    rng = np.random.default_rng(seed)
    Qm, _ = np.linalg.qr(rng.normal(size=(n, n)))
    H = (Qm * rng.uniform(0.5, 5.0, n)) @ Qm.T
    H = (H + H.T) * 0.5
    A = rng.uniform(-1, 1, (p, n))
    A[-1] = A[0]
    x, w = rng.uniform(-1, 1, n), rng.uniform(-1, 1, p)
    q, b = -(H @ x + A.T @ w), A @ x
    x0, w0 = O.kkt_solve(H, A, q, b, 1e-6)
    x1, w1 = KKTSystem(H, A, q, b, handle).solve(1e-6, None, 1e-6, 0)
    res = lambda xx, ww: np.linalg.norm(np.concatenate([H @ xx + A.T @ ww + q, A @ xx - b])) / np.linalg.norm(np.concatenate([q, b]))
    assert res(x1, w1) < max(RTOL, 10 * res(x0, w0))
    assert rel(x1, x) < 1e-8 and rel(x1, x0) < 1e-8          # x is unique; w is not (the two equal rows share their multiplier)
    assert abs((w1[0] + w1[-1]) - (w[0] + w[-1])) < 1e-6


@pytest.mark.parametrize("n", [9, 130])
def test_symmetric_system_indefinite_uses_symsolve(handle, n):
    """Symmetric indefinite, nonsingular: choleskySolve throws, symSolve answers (SymmetricLinearSystem.scala:31-34)."""
    from cvx_b200 import SymmetricLinearSystem
    rng = np.random.default_rng(n)
    Qm, _ = np.linalg.qr(rng.normal(size=(n, n)))
    lam = np.concatenate([rng.uniform(1, 3, n // 2), -rng.uniform(1, 3, n - n // 2)])
    H = (Qm * lam) @ Qm.T
    H = (H + H.T) / 2
    x = rng.uniform(-1, 1, n)
    r = H @ x
    S = SymmetricLinearSystem(H, r, None, handle)
    x1 = S.solve(1e-8, 0)
    x0 = O.symmetricLinearSystemSolve(H, r, 1e-8)
    assert S.info.path == 2
    assert rel(x1, x) < 1e-9 and rel(x1, x0) < 1e-9


def test_symmetric_system_asymmetric_uses_svdsolve(handle):
    """||Q - Q'|| >= 1e-13 -> svdSolve (SymmetricLinearSystem.scala:28-29)."""
    from cvx_b200 import SymmetricLinearSystem
    rng = np.random.default_rng(5)
    n = 40
    H = spd(n, 77) + 1e-3 * rng.uniform(-1, 1, (n, n))
    x = rng.uniform(-1, 1, n)
    r = H @ x
    S = SymmetricLinearSystem(H, r, None, handle)
    x1 = S.solve(1e-8, 0)
    x0 = O.symmetricLinearSystemSolve(H, r, 1e-8)
    assert S.info.path == 3
    assert rel(x1, x) < 1e-8 and rel(x1, x0) < 1e-8


def test_unsolvable_system(handle):
    """Singular matrix, right-hand side outside its range: UnsolvableSystemException (MatrixUtils.scala:627-633)."""
    from cvx_b200 import SymmetricLinearSystem, _lib
    H = np.diag([1.0, 2.0, 0.0, -1.0])
    r = np.array([1.0, 1.0, 1.0, 1.0])
    with pytest.raises(O.UnsolvableSystemException):
        O.symmetricLinearSystemSolve(H, r, 1e-6)
    with pytest.raises(_lib.UnsolvableSystemException):
        SymmetricLinearSystem(H, r, None, handle).solve(1e-6, 0)
    # in range: the pseudo-inverse solution
    r2 = np.array([1.0, 1.0, 0.0, 1.0])
    x1 = SymmetricLinearSystem(H, r2, None, handle).solve(1e-6, 0)
    assert np.allclose(x1, [1.0, 0.5, 0.0, -1.0], atol=1e-12)


@pytest.mark.parametrize("null", [(0,), (3, 7), (1, 2, 11), ()])
def test_kkt_system_reduction(handle, null):
    """KktTest.testKktSystemReduction (KktTest.scala:52-104): wipe rows/columns `null` of H, columns of A and entries
    of g; reduce, solve, pad with zeros; the padded solution solves the ORIGINAL system."""
    import cvx_b200 as cb
    rng = np.random.default_rng(len(null))
    dim = (null[-1] if null else 6) + 12        # the reference uses last + 5 with indices large enough that dim - |null| > 6 equations
    Q = rng.uniform(-1, 1, (dim, dim))
    H = Q.T @ Q
    A = rng.uniform(-1, 1, (6, dim))
    g = rng.uniform(-2, 2, dim)
    r = rng.uniform(-1, 1, 6)
    for j in null:
        H[:, j] = 0.0
        H[j, :] = 0.0
        A[:, j] = 0.0
        g[j] = 0.0
    kd = cb.KKTData(H, A, g, r, None, handle)
    dx, nu = kd.solveReduced(1e-6, None, 1e-10)
    assert (kd.nullIndices or []) == list(null)
    assert np.linalg.norm(H @ dx + A.T @ nu + g) < 1e-9 * max(1.0, np.linalg.norm(g))
    assert np.linalg.norm(A @ dx - r) < 1e-9
    assert all(dx[j] == 0.0 for j in null)
    # against the oracle's reduced -> solve -> pad
    Hr, Ar, gr, rr, nul0 = O.kktDataReduced(H, A, g, r)
    xr, nu0 = O.kkt_solve(Hr, Ar, gr, rr, 1e-10)
    x0 = O.paddVector(xr, nul0) if nul0 else xr
    assert np.linalg.norm(dx - x0) < 1e-8 * max(1.0, np.linalg.norm(x0)) and np.linalg.norm(nu - nu0) < 1e-8 * max(1.0, np.linalg.norm(nu0))
    assert np.allclose(cb.KKTData.paddVector(xr, list(null)), x0) if null else True


def test_kkt_system_reduction_unsolvable(handle):
    """A zero row with a nonzero right-hand side: UnsolvableSystemException (KKTData.scala:78-81)."""
    import cvx_b200 as cb
    rng = np.random.default_rng(5)
    Q = rng.uniform(-1, 1, (9, 9))
    H = Q.T @ Q
    A = rng.uniform(-1, 1, (3, 9))
    g = rng.uniform(-2, 2, 9)
    H[:, 4] = 0.0
    H[4, :] = 0.0
    A[:, 4] = 0.0
    g[4] = 0.5
    with pytest.raises(cb.UnsolvableSystemException):
        cb.KKTData(H, A, g, rng.uniform(-1, 1, 3), None, handle).solveReduced()
    with pytest.raises(O.UnsolvableSystemException):
        O.kktDataReduced(H, A, g, np.zeros(3))

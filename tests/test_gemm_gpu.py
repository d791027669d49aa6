"""DMMA GEMM / SYRK kernel (cvx_b200/csrc/gemm_dmma.cu) against numpy float64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(a_kc, b_kc, A, B):
    Aop = A.T if a_kc else A          # stored (K x M) if a_kc else (M x K)
    Bop = B if b_kc else B.T          # stored (K x N) if b_kc else (N x K)
    return Aop @ Bop


@pytest.mark.parametrize("a_kc,b_kc", [(1, 1), (0, 1), (0, 0), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (130, 70, 37), (257, 129, 300), (64, 1, 200), (1, 5, 3)])
def test_gemm_layouts(handle, a_kc, b_kc, M, N, K):
    from cvx_b200.linalg import dgemm
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    A = rng.uniform(-1, 1, (K, M) if a_kc else (M, K))
    B = rng.uniform(-1, 1, (K, N) if b_kc else (N, K))
    C0 = rng.uniform(-1, 1, (M, N))
    out = dgemm(a_kc, b_kc, M, N, K, 0.7, A, B, -1.3, C0, 0, handle)
    ref = 0.7 * _ref(a_kc, b_kc, A, B) - 1.3 * C0
    assert np.max(np.abs(out - ref)) <= 1e-13 * max(1.0, K ** 0.5) * np.max(np.abs(ref))


@pytest.mark.parametrize("n,k", [(100, 200), (257, 1000), (384, 64)])
def test_syrk_tn_mirrored(handle, n, k):
    """Hessian assembly H = Gs'Gs (tri=2): exactly symmetric, equal to numpy to rounding."""
    from cvx_b200.linalg import dgemm
    rng = np.random.default_rng(n + k)
    G = rng.uniform(-1, 1, (k, n))
    H = dgemm(1, 1, n, n, k, 1.0, G, G, 0.0, np.zeros((n, n)), 2, handle)
    ref = G.T @ G
    assert np.array_equal(H, H.T)
    assert np.max(np.abs(H - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_syrk_nt_lower_only(handle):
    """Cholesky trailing update C -= A A' (tri=1): strict upper triangle untouched."""
    from cvx_b200.linalg import dgemm
    rng = np.random.default_rng(5)
    n, k = 300, 128
    A = rng.uniform(-1, 1, (n, k))
    C0 = rng.uniform(-1, 1, (n, n))
    out = dgemm(0, 0, n, n, k, -1.0, A, A, 1.0, C0, 1, handle)
    ref = C0 - A @ A.T
    iu = np.triu_indices(n, 1)
    assert np.array_equal(out[iu], C0[iu])
    il = np.tril_indices(n)
    assert np.max(np.abs(out[il] - ref[il])) <= 1e-12


@pytest.mark.parametrize("n,k", [(2000, 700), (1700, 513), (2001, 2500), (2300, 520)])
def test_syrk_streamk_shapes(handle, n, k):
    """Tile grids with a partial last wave run the persistent stream-K kernel (split tiles summed in a fixed
    order): H = G'G exactly symmetric and equal to numpy; the NT lower update leaves the upper triangle alone;
    two launches give bitwise identical results (deterministic fix-up)."""
    from cvx_b200.linalg import dgemm
    rng = np.random.default_rng(n * 7 + k)
    G = rng.uniform(-1, 1, (k, n))
    C0 = rng.uniform(-1, 1, (n, n))
    C0 = C0 + C0.T
    H = dgemm(1, 1, n, n, k, 1.0, G, G, 0.5, C0, 2, handle)
    ref = G.T @ G + 0.5 * C0
    assert np.array_equal(H, H.T)
    assert np.max(np.abs(H - ref)) <= 1e-13 * np.max(np.abs(ref))
    H2 = dgemm(1, 1, n, n, k, 1.0, G, G, 0.5, C0, 2, handle)
    assert np.array_equal(H, H2)
    A = np.asfortranarray(G.T)            # n x k, M contiguous
    out = dgemm(0, 0, n, n, k, -1.0, A, A, 1.0, C0, 1, handle)
    iu = np.triu_indices(n, 1)
    assert np.array_equal(out[iu], C0[iu])
    il = np.tril_indices(n)
    ref2 = C0 - A @ A.T
    assert np.max(np.abs(out[il] - ref2[il])) <= 1e-13 * np.max(np.abs(ref2))

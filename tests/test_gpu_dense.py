"""Parity of the dense FP64 device kernels (DMMA GEMM, blocked Cholesky, right-side solves) against
numpy/LAPACK on the same inputs.  Floating point: tolerance 1e-12 relative (FP64 summation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fort(a):
    """C-contiguous buffer holding the column-major image of a (so a.T view is the matrix)."""
    return np.ascontiguousarray(a.T)


@pytest.mark.parametrize("m,n,k,transb", [(128, 128, 16, "t"), (300, 200, 77, "n"), (384, 256, 512, "t"),
                                           (130, 129, 131, "n"), (1024, 640, 256, "t")])
def test_dgemm(mc, m, n, k, transb):
    rng = np.random.default_rng(m + n + k)
    ctx = mc.default_context()
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((k, n))
    C0 = rng.standard_normal((m, n))
    a, c = fort(A), fort(C0)
    b = fort(B.T) if transb == "t" else fort(B)
    ldb = n if transb == "t" else k
    ctx.dgemm(transb, m, n, k, -0.5, a, m, b, ldb, 1.5, c, m)
    ref = -0.5 * A @ B + 1.5 * C0
    assert np.abs(c.T - ref).max() < 1e-12 * max(1.0, np.abs(ref).max()) * k ** 0.5


@pytest.mark.parametrize("n", [5, 128, 200, 384, 1000, 2048])
def test_dpotrf(mc, n):
    rng = np.random.default_rng(n)
    ctx = mc.default_context()
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    a = fort(S)
    info = ctx.dpotrf(a)
    assert info == 0
    L = np.tril(a.T)
    Lref = np.linalg.cholesky(S)
    assert np.abs(L - Lref).max() < 1e-12 * np.abs(Lref).max()
    # strict upper triangle is untouched
    assert np.array_equal(np.triu(a.T, 1), np.triu(S, 1))


def test_dpotrf_reports_first_bad_pivot(mc):
    ctx = mc.default_context()
    n = 300
    S = np.eye(n) * 2.0
    S[170, 170] = -1.0
    a = fort(S)
    assert ctx.dpotrf(a) == 171


@pytest.mark.parametrize("m,n", [(9, 5), (300, 200), (777, 512), (128, 1000)])
def test_dpotrs_right(mc, m, n):
    rng = np.random.default_rng(m * n)
    ctx = mc.default_context()
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    L = np.linalg.cholesky(S)
    X0 = rng.standard_normal((m, n))
    x, l = fort(X0), fort(L)
    ctx.dpotrs_right(m, n, l, n, x, m)
    ref = np.linalg.solve(S, X0.T).T
    assert np.abs(x.T - ref).max() < 1e-11 * np.abs(ref).max()

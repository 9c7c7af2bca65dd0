"""GPU: the DMMA GEMM / SYRK kernel, the recursive Cholesky, triangular solves and SPD inverse
through the C ABI (gpl_dgemm_dev, gpl_dpotrf_dev, gpl_dtrsm_right_dev, gpl_dpotri_dev) against
torch fp64 on the same inputs.  Tolerances are fp64 rounding levels scaled by the inner dimension."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from gplite_b200 import _lib           # noqa: E402
from gpu_util import dev_colmajor, to_np, ptr, sync   # noqa: E402


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0)])
@pytest.mark.parametrize("m,n,k", [(128, 128, 16), (1, 1, 1), (37, 53, 29), (129, 257, 130),
                                   (512, 384, 1000), (300, 5, 4096)])
def test_dgemm(ctx, ta, tb, m, n, k):
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    A = rng.normal(size=(k, m) if ta else (m, k))
    B = rng.normal(size=(n, k) if tb else (k, n))
    Cm = rng.normal(size=(m, n))
    dA, lda = dev_colmajor(A)
    dB, ldb = dev_colmajor(B)
    dC, ldc = dev_colmajor(Cm)
    alpha, beta = -0.7, 1.3
    torch.cuda.synchronize()
    _lib.check(ctx.lib.gpl_dgemm_dev(ctx.h, ta, tb, m, n, k, alpha, ptr(dA), lda, ptr(dB), ldb, beta,
                                     ptr(dC), ldc, 0))
    sync(ctx)
    want = alpha * ((A.T if ta else A) @ (B.T if tb else B)) + beta * Cm
    got = to_np(dC, m)
    assert np.max(np.abs(got - want)) < 1e-13 * (k + 10) * 4


def test_dgemm_lower_only_and_beta_zero(ctx):
    rng = np.random.default_rng(5)
    n, k = 300, 77
    A = rng.normal(size=(n, k))
    Cm = rng.normal(size=(n, n))
    dA, lda = dev_colmajor(A)
    dC, ldc = dev_colmajor(Cm)
    torch.cuda.synchronize()
    _lib.check(ctx.lib.gpl_dgemm_dev(ctx.h, 0, 1, n, n, k, -1.0, ptr(dA), lda, ptr(dA), lda, 1.0,
                                     ptr(dC), ldc, 1))
    sync(ctx)
    got = to_np(dC, n)
    want = Cm - A @ A.T
    assert np.max(np.abs(np.tril(got) - np.tril(want))) < 1e-12
    assert np.array_equal(np.triu(got, 1), np.triu(Cm, 1))          # strict upper untouched
    # beta == 0 must not read C (NaN-safe)
    dC2 = torch.full_like(dC, float("nan"))
    _lib.check(ctx.lib.gpl_dgemm_dev(ctx.h, 0, 1, n, n, k, 1.0, ptr(dA), lda, ptr(dA), lda, 0.0,
                                     ptr(dC2), ldc, 0))
    sync(ctx)
    assert np.max(np.abs(to_np(dC2, n) - A @ A.T)) < 1e-12


def spd(n, seed, cond=1e3):
    rng = np.random.default_rng(seed)
    Q = rng.normal(size=(n, max(n // 2, 1)))
    return Q @ Q.T / Q.shape[1] + np.eye(n) / cond * 10


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 256, 300, 1000, 2049])
def test_dpotrf_trsm_potri(ctx, n):
    A = spd(n, n)
    dA, lda = dev_colmajor(A)
    info = C.c_int32(-1)
    logdet = C.c_double()
    torch.cuda.synchronize()
    _lib.check(ctx.lib.gpl_dpotrf_dev(ctx.h, n, ptr(dA), lda, C.byref(info), C.byref(logdet)))
    sync(ctx)
    L = np.tril(to_np(dA, n))
    Lw = np.linalg.cholesky(A)
    scale = np.max(np.abs(Lw))
    assert info.value == 0
    assert np.max(np.abs(L - Lw)) < 1e-11 * scale * max(1, n / 100)
    assert abs(logdet.value - 2 * np.sum(np.log(np.diag(Lw)))) < 1e-10 * max(1.0, abs(logdet.value))
    # B L^-T and B L^-1
    rng = np.random.default_rng(n + 1)
    for trans in (1, 0):
        m = 70
        B = rng.normal(size=(m, n))
        dB, ldb = dev_colmajor(B)
        torch.cuda.synchronize()
        _lib.check(ctx.lib.gpl_dtrsm_right_dev(ctx.h, trans, m, n, ptr(dA), lda, ptr(dB), ldb))
        sync(ctx)
        X = to_np(dB, m)
        resid = X @ (Lw.T if trans else Lw) - B
        assert np.max(np.abs(resid)) < 1e-10 * max(1, n / 100) * np.max(np.abs(X))
    # inverse
    torch.cuda.synchronize()
    _lib.check(ctx.lib.gpl_dpotri_dev(ctx.h, n, ptr(dA), lda))
    sync(ctx)
    Ainv = np.tril(to_np(dA, n))
    want = np.tril(np.linalg.inv(A))
    assert np.max(np.abs(Ainv - want)) < 1e-9 * np.max(np.abs(want)) * max(1, n / 100)


def test_dpotrf_reports_non_pd(ctx):
    n = 300
    A = spd(n, 3)
    A[200, 200] = -1.0
    dA, lda = dev_colmajor(A)
    info = C.c_int32(0)
    torch.cuda.synchronize()
    st = ctx.lib.gpl_dpotrf_dev(ctx.h, n, ptr(dA), lda, C.byref(info), None)
    sync(ctx)
    assert st == _lib.GPL_ERR_NOT_PD
    assert info.value == 201
    assert b"201" in ctx.lib.gpl_last_error()


def test_dpotrf_reports_non_pd_on_the_lookahead_path(ctx):
    """n >= 4 panels: the factorisation runs on two streams (look-ahead); the first failing pivot must
    still be reported exactly."""
    n, bad = 4608, 3333
    gen = torch.Generator(device="cuda").manual_seed(1)
    Q = torch.randn(n, 64, dtype=torch.float64, device="cuda", generator=gen)
    A = Q @ Q.T / 64 + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    A[bad, bad] = -1.0
    ld = (n + 15) // 16 * 16
    dA = torch.zeros(n, ld, dtype=torch.float64, device="cuda")
    dA[:, :n] = A
    info = C.c_int32(0)
    torch.cuda.synchronize()
    st = ctx.lib.gpl_dpotrf_dev(ctx.h, n, ptr(dA), ld, C.byref(info), None)
    sync(ctx)
    assert st == _lib.GPL_ERR_NOT_PD and info.value == bad + 1


@pytest.mark.parametrize("n", [8192, 8500])
def test_dpotrf_large(ctx, n):
    """Large n (all three GEMM tile configurations, deep recursion, the blocked look-ahead path): check against cuSOLVER through torch on the device, plus the inverse
    built on the same factor."""
    gen = torch.Generator(device="cuda").manual_seed(n)
    Q = torch.randn(n, 96, dtype=torch.float64, device="cuda", generator=gen)
    A = Q @ Q.T / 96 + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    ld = (n + 15) // 16 * 16
    dA = torch.zeros(n, ld, dtype=torch.float64, device="cuda")
    dA[:, :n] = A                                   # symmetric: row-major copy == column-major
    info = C.c_int32(-1)
    logdet = C.c_double()
    torch.cuda.synchronize()
    _lib.check(ctx.lib.gpl_dpotrf_dev(ctx.h, n, ptr(dA), ld, C.byref(info), C.byref(logdet)))
    sync(ctx)
    Lw = torch.linalg.cholesky(A)
    L = torch.tril(dA[:, :n].T)
    assert info.value == 0
    assert float((L - Lw).abs().max()) < 1e-11 * float(Lw.abs().max()) * 50
    assert abs(logdet.value - 2 * float(torch.log(torch.diagonal(Lw)).sum())) < 1e-9 * abs(logdet.value)
    torch.cuda.synchronize()
    _lib.check(ctx.lib.gpl_dpotri_dev(ctx.h, n, ptr(dA), ld))
    sync(ctx)
    Ainv = torch.tril(dA[:, :n].T)
    want = torch.tril(torch.cholesky_inverse(Lw))
    assert float((Ainv - want).abs().max()) < 1e-8 * float(want.abs().max())


def test_column_indexed_helpers_beyond_65535_columns(ctx):
    """gridDim.y is limited to 65535: the helpers that map columns to blockIdx.y (used on nt x n
    cross-covariances when predicting from more than 65535 training points) must stride."""
    m, n = 8, 70001
    ld = 16
    gen = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(n, ld, dtype=torch.float64, device="cuda", generator=gen)     # column-major m x n
    r = torch.rand(m, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    c = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    want = 1.5 * A[:, :m] * r[None, :] * c[:, None]
    torch.cuda.synchronize()
    _lib.check(ctx.lib.gpl_scale_block_dev(ctx.h, m, n, ptr(A), ld, ptr(r), ptr(c), 1.5, 0.0, 0))
    sync(ctx)
    assert float((A[:, :m] - want).abs().max()) < 1e-14

"""Development probe (run under gpurun): raw FP64 pipe rates, cuBLAS/cuSOLVER yardsticks via torch,
and timings of this repo's kernels through the C ABI.  Prints one JSON object per line.

    python tools/gpu_probe.py [section ...]     sections: pipes cublas gemm potrf kbuild egrad
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gplite_b200 as g                    # noqa: E402
from gplite_b200 import _lib               # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def ev_time(fn, ctx=None, reps=3, warm=1):
    """CUDA-event timing on the default stream around fn(); ctx work is fenced by ctx.sync()."""
    for _ in range(warm):
        fn()
    if ctx:
        ctx.sync()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        if ctx:
            ctx.sync()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), ts


def pipes():
    lib = C.CDLL(os.path.join(ROOT, "tools", "libprobe_fp64.so"))
    lib.probe_fp64.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    for mode, name in ((0, "DMMA.8x8x4"), (2, "DMMA 4x8 tile pattern"), (1, "DFMA")):
        for threads, bps in ((256, 1), (512, 1), (256, 2)):
            tf = C.c_double()
            rc = lib.probe_fp64(mode, threads, bps, 4000, C.byref(tf))
            emit(section="pipes", op=name, threads=threads, blocks_per_sm=bps, rc=rc, tflops=round(tf.value, 2))


def cublas():
    for n in (4096, 8192, 16384):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        t, _ = ev_time(lambda: torch.matmul(a, b), reps=3)
        emit(section="cublas", op="dgemm", n=n, ms=round(t * 1e3, 3), tflops=round(2 * n ** 3 / t / 1e12, 2))
        del a, b
    for n in (8192, 16384, 32768):
        q = torch.randn(n, 64, dtype=torch.float64, device="cuda")
        a = q @ q.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
        del q
        t, _ = ev_time(lambda: torch.linalg.cholesky(a), reps=2)
        emit(section="cusolver", op="dpotrf", n=n, ms=round(t * 1e3, 2), tflops=round(n ** 3 / 3 / t / 1e12, 2))
        del a
        torch.cuda.empty_cache()


def gemm(ctx):
    for (m, n, k, ta, tb, lower) in [(8192, 8192, 8192, 0, 1, 0), (8192, 8192, 8192, 0, 0, 0),
                                      (8192, 8192, 8192, 1, 0, 0), (16384, 16384, 16384, 0, 1, 0),
                                      (16384, 16384, 16384, 0, 1, 1), (16384, 16384, 512, 0, 1, 1),
                                      (16384, 16384, 128, 0, 1, 0), (16384, 128, 128, 0, 1, 0),
                                      (32768, 32768, 4096, 0, 1, 1)]:
        lda = k if ta else m
        ldb = n if tb else k
        A = torch.randn((m if ta else k), lda, dtype=torch.float64, device="cuda")
        B = torch.randn((k if tb else n), ldb, dtype=torch.float64, device="cuda")
        Cm = torch.zeros(n, m, dtype=torch.float64, device="cuda")
        f = lambda: _lib.check(ctx.lib.gpl_dgemm_dev(ctx.h, ta, tb, m, n, k, 1.0, C.c_void_p(A.data_ptr()), lda,
                                                     C.c_void_p(B.data_ptr()), ldb, 0.0,
                                                     C.c_void_p(Cm.data_ptr()), m, lower))
        t, _ = ev_time(f, ctx, reps=3)
        fl = 2.0 * m * n * k * (0.5 if lower else 1.0)
        emit(section="gemm", m=m, n=n, k=k, ta=ta, tb=tb, lower=lower, ms=round(t * 1e3, 3),
             tflops=round(fl / t / 1e12, 2))
        del A, B, Cm
        torch.cuda.empty_cache()


def potrf(ctx):
    for n in (4096, 8192, 16384, 32768):
        q = torch.randn(n, 64, dtype=torch.float64, device="cuda")
        a0 = q @ q.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
        del q
        a = a0.clone()
        info = C.c_int32()
        ld = C.c_double()

        def f():
            a.copy_(a0)
            torch.cuda.synchronize()
            _lib.check(ctx.lib.gpl_dpotrf_dev(ctx.h, n, C.c_void_p(a.data_ptr()), n, C.byref(info), C.byref(ld)))
        tcopy, _ = ev_time(lambda: (a.copy_(a0), torch.cuda.synchronize()), reps=2)
        t, _ = ev_time(f, ctx, reps=2)
        t -= tcopy
        emit(section="potrf", n=n, ms=round(t * 1e3, 2), tflops=round(n ** 3 / 3 / t / 1e12, 2),
             launches=ctx.launch_count())
        if n <= 16384:
            f()
            ctx.sync()
            t2, _ = ev_time(lambda: _lib.check(ctx.lib.gpl_dpotri_dev(ctx.h, n, C.c_void_p(a.data_ptr()), n)), ctx,
                            reps=1, warm=0)
            emit(section="potri", n=n, ms=round(t2 * 1e3, 2), tflops=round(2 * n ** 3 / 3 / t2 / 1e12, 2))
        del a, a0
        torch.cuda.empty_cache()


def kbuild(ctx):
    rng = np.random.default_rng(2)
    for kind, mk in (("cf_matern52", lambda: g.cf_matern52(lscale=1.3)), ("cf_sexp", lambda: g.cf_sexp(lscale=1.3)),
                     ("cf_matern32", lambda: g.cf_matern32(lscale=1.3)), ("cf_nn", lambda: g.cf_nn(normalize=False)),
                     ("sexp*lin", None)):
        n, d = 32768, 8
        x = torch.from_numpy(np.asfortranarray(rng.normal(size=(n, d))).T.copy()).cuda()   # (d, n) == col-major n x d
        K = torch.empty(n, n, dtype=torch.float64, device="cuda")
        if mk is None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                cfs = [g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed())]
        else:
            cfs = [mk()]
        factors, _ = g.model.flatten_cfs(cfs)
        pm = _lib.PackedModel(factors)
        for same in (1, 0):
            f = lambda: _lib.check(ctx.lib.gpl_cf_eval_dev(ctx.h, pm.ref(), C.c_void_p(x.data_ptr()), n,
                                                           C.c_void_p(x.data_ptr()), n, d, same, 0, 0, 1e-6,
                                                           C.c_void_p(K.data_ptr()), n))
            t, _ = ev_time(f, ctx, reps=3)
            emit(section="kbuild", kind=kind, n=n, d=d, same=same, ms=round(t * 1e3, 3),
                 gbs=round(8.0 * n * n / t / 1e9, 1))
        del K
        torch.cuda.empty_cache()


def egrad(ctx):
    for n in (4096, 8192, 16384, 32768):
        rng = np.random.default_rng(2)
        d = 8
        X = rng.normal(size=(n, d))
        y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
        gp = g.gp_init(cfs=g.cf_matern52(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        e, grad = g.gp_energy_grad(gp, X, y, ctx=ctx)
        t1 = time.perf_counter() - t0
        l1 = ctx.launch_count()
        t0 = time.perf_counter()
        e, grad = g.gp_energy_grad(gp, X, y, ctx=ctx)
        t2 = time.perf_counter() - t0
        emit(section="egrad", n=n, first_s=round(t1, 3), s=round(t2, 3), energy=e, grad=list(grad),
             launches=l1 - l0, tflops=round(n ** 3 / t2 / 1e12, 2))


if __name__ == "__main__":
    want = sys.argv[1:] or ["pipes", "cublas", "gemm", "potrf", "kbuild", "egrad"]
    ctx = g.default_context()
    emit(section="device", name=torch.cuda.get_device_name(0), sms=torch.cuda.get_device_properties(0).multi_processor_count)
    for s in want:
        try:
            if s == "pipes":
                pipes()
            elif s == "cublas":
                cublas()
            else:
                globals()[s](ctx)
        except Exception as ex:          # keep going: one call = many minutes
            emit(section=s, error=repr(ex))

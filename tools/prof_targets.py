"""Small single-purpose workloads for ncu captures (run under gpurun, one target per process).

    python tools/prof_targets.py gemm | syrk | kbuild | potrf N | egrad N
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                    # noqa: E402
from gplite_b200 import _lib               # noqa: E402

ctx = g.default_context()
what = sys.argv[1]
P = lambda t: C.c_void_p(t.data_ptr())

if what in ("gemm", "syrk"):
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    Cm = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        _lib.check(ctx.lib.gpl_dgemm_dev(ctx.h, 0, 1, n, n, n, -1.0, P(A), n, P(A if what == "syrk" else B), n,
                                         1.0, P(Cm), n, 1 if what == "syrk" else 0))
    ctx.sync()
elif what == "kbuild":
    n, d = 32768, 8
    rng = np.random.default_rng(2)
    x = torch.from_numpy(rng.normal(size=(d, n))).cuda()
    K = torch.empty(n, n, dtype=torch.float64, device="cuda")
    factors, _ = g.model.flatten_cfs([g.cf_matern52(lscale=1.3)])
    pm = _lib.PackedModel(factors)
    torch.cuda.synchronize()
    for _ in range(2):
        _lib.check(ctx.lib.gpl_cf_eval_dev(ctx.h, pm.ref(), P(x), n, P(x), n, d, 1, 0, 0, 1e-6, P(K), n))
    ctx.sync()
elif what == "potrf":
    n = int(sys.argv[2])
    q = torch.randn(n, 64, dtype=torch.float64, device="cuda")
    a = q @ q.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    info = C.c_int32()
    _lib.check(ctx.lib.gpl_dpotrf_dev(ctx.h, n, P(a), n, C.byref(info), None))
    ctx.sync()
elif what == "egrad":
    n = int(sys.argv[2])
    rng = np.random.default_rng(2)
    X = rng.normal(size=(n, 8))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    gp = g.gp_init(cfs=g.cf_matern52(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
    print(g.gp_energy_grad(gp, X, y, ctx=ctx))
elif what == "sweeps":
    # the HBM-bound sweeps of BASELINE config 3 (n = 50000, d = 10, cf_sexp * cf_lin + lik_bernoulli):
    # two Newton steps (pseudodata_la, build_B, potrf, trsv, gemv_t, loglik) and one EP sweep
    # (ep_sites + the same linear algebra) -- the ncu targets are the elementwise / GEMV kernels
    import warnings
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    rng = np.random.default_rng(3)
    X = rng.normal(size=(n, 10))
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(2 * X[:, 0] - X[:, 1])))).astype(float)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfs = g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed())
        gp = g.gp_init(cfs=cfs, lik=g.lik_bernoulli(), approx=g.approx_laplace(maxiter=2))
        g.gp_fit(gp, X, y, ctx=ctx)
        print("laplace", gp.fit["log_evidence"])
        gp = g.gp_init(cfs=cfs, lik=g.lik_bernoulli(), approx=g.approx_ep(maxiter=1))
        g.gp_fit(gp, X, y, ctx=ctx)
        print("ep", gp.fit["log_evidence"])
elif what == "gemv":
    # f = K alpha at config 3's size: one pass over a 50000 x 50000 matrix (20 GB)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        _lib.check(ctx.lib.gpl_dgemv_dev(ctx.h, 1, n, n, P(A), n, P(x), P(y)))
        _lib.check(ctx.lib.gpl_dgemv_dev(ctx.h, 0, n, n, P(A), n, P(x), P(y)))
    ctx.sync()

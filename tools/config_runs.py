"""BASELINE.json configs 3 and 4 at full size on one B200 (run under gpurun): wall time, iterations
and achieved rates through the public API.  One JSON object per line.

    python tools/config_runs.py cfg3_laplace | cfg3_ep | cfg3_grad | cfg4_fitc [scale]

scale < 1 shrinks n (e.g. 0.5) for quick checks.
"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g          # noqa: E402

what = sys.argv[1]
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ctx = g.default_context()


def emit(**kw):
    print(json.dumps(kw), flush=True)


if what.startswith("cfg3"):
    n, d = int(50000 * scale), 10
    rng = np.random.default_rng(3)
    X = rng.normal(size=(n, d))
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(2 * X[:, 0] - X[:, 1])))).astype(float)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cf = g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed())
    if what == "cfg3_grad":          # energy + analytic gradient at the Newton mode (new capability)
        gp = g.gp_init(cfs=cf, lik=g.lik_bernoulli(), approx=g.approx_laplace(maxiter=100, tol=1e-9))
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e, gr = g.gp_energy_grad(gp, X, y, ctx=ctx)
        t = time.perf_counter() - t0
        emit(config="3: energy + analytic gradient, cf_sexp*cf_lin + lik_bernoulli, laplace (tol 1e-9)", n=n, d=d,
             seconds=t, energy=e, grad=[float(v) for v in gr], launches=ctx.launch_count() - l0)
        sys.exit(0)
    approx = g.approx_laplace() if what == "cfg3_laplace" else g.approx_ep(damping=0.9, quad_order=11, maxiter=100)
    gp = g.gp_init(cfs=cf, lik=g.lik_bernoulli(), approx=approx)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.gp_fit(gp, X, y, ctx=ctx)
    t = time.perf_counter() - t0
    it = gp.fit["iters"]
    per_iter_flops = n ** 3 / 3 if what == "cfg3_laplace" else 2 * n ** 3 / 3
    emit(config="3: binary classification, cf_sexp*cf_lin + lik_bernoulli, " + approx.name, n=n, d=d, seconds=t,
         iters=it, log_evidence=gp.fit["log_evidence"], converged=gp.fit["converged"],
         algorithmic_tflops=per_iter_flops * it / t / 1e12, launches=ctx.launch_count() - l0)
    t0 = time.perf_counter()
    p = g.gp_pred(gp, X[:4096], var=True)
    emit(config="3: gp_pred mean+var", nt=4096, seconds=time.perf_counter() - t0,
         tflops=n * n * 4096.0 / (time.perf_counter() - t0) / 1e12, mean0=float(p["mean"][0]), var0=float(p["var"][0]))
elif what == "cfg4_fitc":
    n, d, m = int(2000000 * scale), 6, 4096
    rng = np.random.default_rng(4)
    X = rng.uniform(-3, 3, size=(n, d))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.5 * rng.normal(size=n)
    Zi = X[rng.permutation(n)[:m]].copy()
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_gaussian(sigma=0.5), method=g.method_fitc(inducing=Zi))
    flops = 2.0 * n * m * m + 2 * m ** 3 / 3      # TRSM for D + SYRK for A + two m^3/3 factorisations

    def used_gb():
        try:
            import pynvml
            pynvml.nvmlInit()
            return pynvml.nvmlDeviceGetMemoryInfo(pynvml.nvmlDeviceGetHandleByIndex(0)).used / 1e9
        except Exception:
            return None
    for run in ("cold", "warm"):
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        g.gp_fit(gp, X, y, ctx=ctx)
        t = time.perf_counter() - t0
        emit(config="4: method_fitc, cf_sexp + lik_gaussian", run=run, n=n, d=d, m=m, seconds=t,
             log_evidence=gp.fit["log_evidence"], algorithmic_tflops=flops / t / 1e12,
             kxz_gb_if_materialised=8.0 * n * m / 1e9, device_memory_used_gb=used_gb(),
             streamed=os.environ.get("GPL_FITC_STREAM_GB", "8 (default)"), launches=ctx.launch_count() - l0)
    t0 = time.perf_counter()
    p = g.gp_pred(gp, X[:100000], var=True)
    emit(config="4: gp_pred mean+var (FITC)", nt=100000, seconds=time.perf_counter() - t0,
         mean0=float(p["mean"][0]), var0=float(p["var"][0]))

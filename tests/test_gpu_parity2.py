"""GPU parity, round 2: the control-flow branches and config-shaped models the first round's tests
never reached.  Everything goes through the C ABI and is compared with the oracle restatement on
the same seeded inputs (tolerance 1e-8 relative unless stated).

* EP sweeps that are REJECTED (max|delta fmean| grew: damp *= damping, sweep discarded,
  R/ep.R:306-310 and :355-360) and EP runs that hit maxiter (log_evidence = -Inf, R/ep.R:322-325,
  :366-369), method_full and method_fitc.
* BASELINE config 3's model (cf_sexp * cf_lin, d = 10, Bernoulli) under Laplace and EP at n = 4096.
* BASELINE config 4's shape (FITC, d = 6, m = 512) at n = 20000.
"""
import copy
import math
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gplite_b200 as g                      # noqa: E402
from oracle import gplite_oracle as o        # noqa: E402
from helpers import scaled_err               # noqa: E402

TOL = 1e-8


def _separable(seed, n=150):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, 2))
    y = (X[:, 0] > 0).astype(float)
    Z = X[rng.choice(n, 30, replace=False)].copy()
    return X, y, Z


def _decisive(history, rel=1e-4):
    """every accept / reject decision of the oracle run was far from a tie, so that 1e-10-level
    differences cannot flip the control flow"""
    return all(abs(a - b) > rel * b for a, b in history if np.isfinite(b))


@pytest.mark.parametrize("method", ["full", "fitc"])
def test_ep_rejected_sweeps(ctx, method):
    """Probit EP on separable data with a large magnitude: several sweeps increase max|delta fmean|
    and are discarded with the damping shrunk (R/ep.R:306-310 / :355-360).  iters, the number of
    accepted sweeps and the final damp value must equal the oracle's."""
    X, y, Z = _separable(7)
    meth = g.method_full() if method == "full" else g.method_fitc(inducing=Z)
    gp = g.gp_init(cfs=g.cf_sexp(lscale=0.5, magn=5.0), lik=g.lik_bernoulli(link="probit"), method=meth,
                   approx=g.approx_ep(damping=0.9, maxiter=100))
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    fo = gpo.fit
    assert fo["iters"] - fo["accepted"] >= 5 and _decisive(fo["history"])       # the scenario does reject
    assert gp.fit["iters"] == fo["iters"]
    assert gp.fit["accepted"] == fo["accepted"]
    assert abs(gp.fit["damp"] - fo["damp"]) < 1e-14
    assert gp.fit["converged"] and math.isfinite(gp.fit["log_evidence"])
    assert abs(gp.fit["log_evidence"] - fo["log_evidence"]) < TOL * abs(fo["log_evidence"])
    vt = TOL if method == "full" else 1e-7
    assert scaled_err(gp.fit["fmean"], fo["fmean"]) < TOL
    assert scaled_err(gp.fit["fvar"], fo["fvar"]) < vt
    assert scaled_err(gp.fit["cavity_mean"], fo["cavity_mean"]) < vt
    assert scaled_err(gp.fit["cavity_var"], fo["cavity_var"]) < vt
    Xt = np.random.default_rng(70).normal(size=(40, 2))
    p, po = g.gp_pred(gp, Xt, var=True), o.gp_pred(gpo, Xt, var=True)
    assert scaled_err(p["mean"], po["mean"]) < TOL
    assert scaled_err(p["var"], po["var"]) < vt


@pytest.mark.parametrize("method", ["full", "fitc"])
def test_ep_maxiter_is_minus_inf(ctx, method):
    """EP that has not converged after maxiter sweeps (rejections included in the count) returns
    log_evidence = -Inf (R/ep.R:322-325 / :366-369) with the last sweep's site state."""
    X, y, Z = _separable(8 if method == "fitc" else 7)
    meth = g.method_full() if method == "full" else g.method_fitc(inducing=Z)
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0, magn=20.0), lik=g.lik_bernoulli(link="probit"), method=meth,
                   approx=g.approx_ep(damping=0.95, maxiter=40))
    gpo = copy.deepcopy(gp)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.gp_fit(gp, X, y, ctx=ctx)
        o.gp_fit(gpo, X, y)
    fo = gpo.fit
    assert fo["iters"] == 40 and fo["log_evidence"] == -np.inf and _decisive(fo["history"])
    assert gp.fit["iters"] == 40 and not gp.fit["converged"]
    assert gp.fit["log_evidence"] == -math.inf
    assert gp.fit["accepted"] == fo["accepted"]
    assert abs(gp.fit["damp"] - fo["damp"]) < 1e-14
    assert scaled_err(gp.fit["fmean"], fo["fmean"]) < 1e-7      # state of the last (possibly rejected) sweep
    assert scaled_err(gp.fit["z"], fo["z"]) < 1e-6


def test_ep_maxiter_one_keeps_evidence(ctx):
    """maxiter = 1 is exempt from the -Inf rule (`maxiter > 1 && iter == maxiter`, R/ep.R:322)."""
    X, y, _ = _separable(7)
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_bernoulli(), approx=g.approx_ep(maxiter=1))
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    assert gp.fit["iters"] == 1 and math.isfinite(gp.fit["log_evidence"])
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])


def _config3(n, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, 10))
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(2 * X[:, 0] - X[:, 1])))).astype(float)
    return X, y


@pytest.mark.parametrize("approx", ["laplace", "ep"])
def test_config3_model_at_4096(ctx, approx):
    """BASELINE config 3's model -- cf_sexp(lscale=1) * cf_lin(prior_magn=prior_fixed()), d = 10,
    lik_bernoulli -- under approx_laplace and approx_ep at n = 4096 (SURVEY 8d's generator)."""
    X, y = _config3(4096)
    ap = g.approx_laplace() if approx == "laplace" else g.approx_ep()
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed()), lik=g.lik_bernoulli(),
                   approx=ap)
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    assert gp.fit["iters"] == gpo.fit["iters"]
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
    assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
    if approx == "ep":
        assert gp.fit["accepted"] == gpo.fit["accepted"]
        assert scaled_err(gp.fit["fvar"], gpo.fit["fvar"]) < TOL
        assert scaled_err(gp.fit["cavity_var"], gpo.fit["cavity_var"]) < TOL
    Xt, _ = _config3(64, seed=33)
    p, po = g.gp_pred(gp, Xt, var=True), o.gp_pred(gpo, Xt, var=True)
    assert scaled_err(p["mean"], po["mean"]) < TOL
    assert scaled_err(p["var"], po["var"]) < TOL


@pytest.mark.parametrize("streamed", [False, True])
def test_config4_shape_fitc_20000(ctx, streamed, monkeypatch):
    """BASELINE config 4's shape -- method_fitc, d = 6, X ~ U(-3, 3), inducing points = rows of a seeded
    permutation, cf_sexp(lscale=1) + lik_gaussian(sigma=0.5) -- at n = 20000, m = 512; with Kxz resident
    and with Kxz streamed (never materialised: what the n = 2M fit does)."""
    monkeypatch.setenv("GPL_FITC_STREAM_GB", "0" if streamed else "1000")
    monkeypatch.setenv("GPL_FITC_CHUNK", "4096")              # several row chunks
    rng = np.random.default_rng(4)
    n, m, d = 20000, 512, 6
    X = rng.uniform(-3, 3, size=(n, d))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.5 * rng.normal(size=n)
    Z = X[rng.permutation(n)[:m]].copy()
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_gaussian(sigma=0.5), method=g.method_fitc(inducing=Z))
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
    assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
    assert scaled_err(gp.fit["alpha"], gpo.fit["alpha"]) < 1e-6
    Xt = rng.uniform(-3.2, 3.2, size=(200, d))
    p, po = g.gp_pred(gp, Xt, var=True), o.gp_pred(gpo, Xt, var=True)
    assert scaled_err(p["mean"], po["mean"]) < TOL
    assert scaled_err(p["var"], po["var"]) < 1e-7          # the reference's own formula cancels at ~1e-9


@pytest.mark.parametrize("approx", ["laplace", "ep"])
def test_fitc_streamed_non_gaussian(ctx, approx, monkeypatch):
    """Newton / EP loops of method_fitc with Kxz regenerated chunk by chunk in every pass (several chunks:
    GPL_FITC_STREAM_GB=0 and a small workspace are forced), vs the oracle."""
    monkeypatch.setenv("GPL_FITC_STREAM_GB", "0")
    monkeypatch.setenv("GPL_FITC_CHUNK", "1024")
    rng = np.random.default_rng(12)
    n, m = 5000, 96
    X = rng.normal(size=(n, 2))
    f = 2 * X[:, 0] - X[:, 1]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(float)
    Zi = X[rng.permutation(n)[:m]].copy()
    ap = g.approx_laplace() if approx == "laplace" else g.approx_ep()
    gp = g.gp_init(cfs=g.cf_matern52(lscale=1.2), lik=g.lik_bernoulli(), method=g.method_fitc(inducing=Zi), approx=ap)
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    assert gp.fit["iters"] == gpo.fit["iters"]
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
    assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
    if approx == "ep":
        assert scaled_err(gp.fit["fvar"], gpo.fit["fvar"]) < 1e-7
    Xt = rng.normal(size=(30, 2))
    p, po = g.gp_pred(gp, Xt, var=True), o.gp_pred(gpo, Xt, var=True)
    assert scaled_err(p["mean"], po["mean"]) < TOL and scaled_err(p["var"], po["var"]) < 1e-7
    with pytest.raises(g.GplError):
        g.fit_member(gp, "Kxz")               # never materialised

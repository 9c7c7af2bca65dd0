"""GPU: gp_fit / gp_energy / gp_pred / gp_energy_grad through the C ABI against the oracle on
the same seeded inputs.  North-star tolerance: log marginal likelihood, gradients, predictive
mean/variance within 1e-8 relative (fp64)."""
import copy
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gplite_b200 as g                      # noqa: E402
from oracle import gplite_oracle as o        # noqa: E402
from helpers import model_grid, basic_data, rel_err, scaled_err, outlier_counts   # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-8


def test_gaussian_grid_energy_and_pred(ctx):
    """tests/testthat/test_basic.R:64-132 on the golden fixture: our fit vs the textbook KAT
    (reference tolerance 1e-3) and vs the oracle (1e-8)."""
    z = np.load(os.path.join(GOLD, "gaussian_kat.npz"))
    x, xt, y = z["x"], z["xt"], z["y"]
    grid = model_grid()
    for mi, j in z["index"]:
        if j != 0:
            continue
        gp = g.gp_init(cfs=copy.deepcopy(grid[mi]), lik=g.lik_gaussian())
        g.set_param(gp, z[f"m{mi}_{j}_param"])
        g.gp_fit(gp, x, y, ctx=ctx)
        le = z[f"m{mi}_{j}_le"]
        assert abs(gp.fit["log_evidence"] - le) <= 1e-3 * abs(le)
        gpo = g.gp_init(cfs=copy.deepcopy(grid[mi]), lik=g.lik_gaussian())
        g.set_param(gpo, z[f"m{mi}_{j}_param"])
        o.gp_fit(gpo, x, y)
        assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) <= TOL * abs(gpo.fit["log_evidence"]), grid[mi]
        assert scaled_err(gp.fit["alpha"], gpo.fit["alpha"]) < 1e-6     # alpha is cond(C)-sensitive
        pred = g.gp_pred(gp, xt, var=True, jitter=1e-6)
        po = o.gp_pred(gpo, xt, var=True, jitter=1e-6)
        assert scaled_err(pred["mean"], po["mean"]) < TOL, grid[mi]
        assert np.max(np.abs(pred["var"] - po["var"])) < TOL * max(np.max(np.abs(po["var"])), 1e-3), grid[mi]
        assert scaled_err(pred["mean"], z[f"m{mi}_{j}_mean"]) < 1e-3


def test_readme_toy_regression(ctx):
    """BASELINE config 0: n=200 1-D, cf_sexp + lik_gaussian, fit, energy, predict on a 300-grid."""
    rng = np.random.default_rng(32004)
    n = 200
    x = rng.normal(size=n)
    y = np.sin(3 * x) * np.exp(-np.abs(x)) + 0.1 * rng.normal(size=n)
    xt = np.linspace(-4, 4, 300)
    gp = g.gp_init(cfs=g.cf_sexp(), lik=g.lik_gaussian())
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, x, y, ctx=ctx)
    o.gp_fit(gpo, x, y)
    assert abs(g.gp_energy(gp) - o.gp_energy(gpo)) < TOL * abs(o.gp_energy(gpo))
    p = g.gp_pred(gp, xt, var=True)
    po = o.gp_pred(gpo, xt, var=True)
    assert scaled_err(p["mean"], po["mean"]) < TOL
    assert scaled_err(p["var"], po["var"]) < TOL
    Cc = g.fit_member(gp, "C_chol")
    assert scaled_err(Cc, gpo.fit["C_chol"]) < 1e-9
    # stale-model guard (tests/testthat/test_pred.R:102-110)
    g.set_param(gp, g.get_param_vector(gp) + 0.1)
    with pytest.raises(RuntimeError):
        g.gp_pred(gp, xt)


def test_unfitted_model_predicts_prior(ctx):
    gp = g.gp_init(cfs=[g.cf_matern32(magn=1.7), g.cf_const(magn=0.4)])
    xt = np.linspace(-1, 1, 9)
    p = g.gp_pred(gp, xt, var=True, ctx=ctx)
    assert np.all(p["mean"] == 0)
    assert np.max(np.abs(p["var"] - (1.7 ** 2 + 0.4 ** 2))) < 1e-14


def _grad_models():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return [
            [g.cf_sexp(lscale=0.8)], [g.cf_matern32(lscale=0.9)], [g.cf_matern52(lscale=1.1)],
            [g.cf_nn()], [g.cf_lin(), g.cf_const(magn=0.5)],
            [g.cf_sexp(vars=[0], lscale=0.7) * g.cf_lin(vars=[1], prior_magn=g.prior_fixed())],
            [g.cf_periodic(vars=[0], period=1.7), g.cf_matern52(prior_lscale=g.prior_lognormal())],
            [g.cf_sexp(normalize=True) * g.cf_matern32() * g.cf_const(prior_magn=g.prior_fixed())],
        ]


def test_energy_and_gradient_vs_oracle(ctx):
    rng = np.random.default_rng(5)
    n = 333
    X = rng.normal(size=(n, 2))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + 0.3 * rng.normal(size=n)
    for cfs in _grad_models():
        gp = g.gp_init(cfs=cfs, lik=g.lik_gaussian(sigma=0.4))
        gpo = copy.deepcopy(gp)
        e, grad = g.gp_energy_grad(gp, X, y, ctx=ctx)
        eo, go = o.energy_grad_gaussian_full(gpo, X, y)
        assert abs(e - eo) < TOL * abs(eo), cfs
        # the period entry of the oracle gradient is itself a 1e-6 central difference
        tol = 1e-6 if any(c.name == "cf_periodic" for c in cfs) else TOL
        assert np.max(np.abs(grad - go)) < tol * max(1.0, np.max(np.abs(go))), (cfs, grad, go)


def _classification(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    f = 2 * X[:, 0] - X[:, 1]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(float)
    return X, y


@pytest.mark.parametrize("link", ["logit", "probit"])
def test_laplace_bernoulli(ctx, link):
    X, y = _classification(400, 3, 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfs = g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed())
    gp = g.gp_init(cfs=[cfs, g.cf_const(magn=0.5)], lik=g.lik_bernoulli(link), approx=g.approx_laplace())
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    assert gp.fit["iters"] == gpo.fit["iters"]
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
    assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
    Xt, _ = _classification(50, 3, 4)
    p = g.gp_pred(gp, Xt, var=True, transform=True)
    po = o.gp_pred(gpo, Xt, var=True, transform=True)
    assert scaled_err(p["mean"], po["mean"]) < TOL
    assert scaled_err(p["var"], po["var"]) < TOL


def test_laplace_other_likelihoods(ctx):
    rng = np.random.default_rng(8)
    n = 250
    X = rng.normal(size=(n, 2))
    f = X[:, 0] - 0.5 * X[:, 1]
    trials = rng.integers(1, 12, size=n).astype(float)
    yb = rng.binomial(trials.astype(int), 1 / (1 + np.exp(-f))).astype(float)
    yp = rng.poisson(np.exp(0.5 * f + 1.0)).astype(float)
    off = rng.normal(size=n) * 0.2
    cases = [(g.lik_binomial("logit"), yb, dict(trials=trials)),
             (g.lik_binomial("probit"), yb, dict(trials=trials, offset=off)),
             (g.lik_poisson(), yp, dict(offset=off)),
             (g.lik_poisson(), yp, dict())]
    for lik, y, extra in cases:
        gp = g.gp_init(cfs=g.cf_matern52(lscale=1.2), lik=lik)
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, y, ctx=ctx, **extra)
        o.gp_fit(gpo, X, y, **extra)
        assert gp.fit["iters"] == gpo.fit["iters"], lik
        assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"]), lik
        assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL, lik


def test_ep_bernoulli(ctx):
    X, y = _classification(300, 3, 5)
    for link, damping in (("logit", 0.9), ("probit", 0.8)):
        gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_bernoulli(link),
                       approx=g.approx_ep(damping=damping))
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, y, ctx=ctx)
        o.gp_fit(gpo, X, y)
        assert gp.fit["iters"] == gpo.fit["iters"]
        assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
        assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
        assert scaled_err(gp.fit["fvar"], gpo.fit["fvar"]) < TOL
        assert scaled_err(gp.fit["cavity_mean"], gpo.fit["cavity_mean"]) < TOL
        Xt, _ = _classification(40, 3, 6)
        p = g.gp_pred(gp, Xt, var=True)
        po = o.gp_pred(gpo, Xt, var=True)
        assert scaled_err(p["mean"], po["mean"]) < TOL
        assert scaled_err(p["var"], po["var"]) < TOL


def test_non_pd_aborts_full_gp(ctx):
    # duplicated inputs + zero jitter + tiny noise: B stays PD (it is I + ...), so force failure
    # through a NaN input instead; the library must report, not crash
    X = np.array([[0.0], [1.0], [np.nan]])
    y = np.zeros(3)
    gp = g.gp_init(cfs=g.cf_sexp(), lik=g.lik_gaussian())
    with pytest.raises(g.GplError):
        g.gp_fit(gp, X, y, ctx=ctx)


# ------------------------------------------------------------------------------ method_fitc
def _fitc_data(n, m, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, size=(n, d))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.5 * rng.normal(size=n)
    Zi = X[rng.permutation(n)[:m]].copy()
    return X, y, Zi


@pytest.mark.parametrize("n,m", [(500, 37), (1500, 200)])
def test_fitc_gaussian(ctx, n, m):
    """FITC fit + prediction vs the oracle restatement of
    R/gp_fit.R:101-123, R/laplace.R:122-152, R/linear_algebra.R:3-126, R/gp_pred.R:234-296."""
    X, y, Zi = _fitc_data(n, m, 2, 4)
    for cfs in ([g.cf_sexp(lscale=1.0)], [g.cf_matern32(lscale=1.5), g.cf_const(magn=0.3)]):
        gp = g.gp_init(cfs=cfs, lik=g.lik_gaussian(sigma=0.5), method=g.method_fitc(inducing=Zi))
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, y, ctx=ctx)
        o.gp_fit(gpo, X, y)
        assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
        assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
        assert scaled_err(gp.fit["alpha"], gpo.fit["alpha"]) < 1e-6
        assert scaled_err(g.fit_member(gp, "Kxz"), gpo.fit["Kxz"]) < 1e-12
        assert scaled_err(g.fit_member(gp, "Kz_chol"), gpo.fit["Kz_chol"]) < 1e-9
        Xt = np.random.default_rng(9).uniform(-3.5, 3.5, size=(60, 2))
        p = g.gp_pred(gp, Xt, var=True)
        po = o.gp_pred(gpo, Xt, var=True)
        assert scaled_err(p["mean"], po["mean"]) < TOL
        assert scaled_err(p["var"], po["var"]) < 1e-7     # reference formula cancels ~1e-9 itself


def test_fitc_laplace_and_ep_bernoulli(ctx):
    rng = np.random.default_rng(12)
    n, m = 600, 50
    X = rng.normal(size=(n, 2))
    f = 2 * X[:, 0] - X[:, 1]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(float)
    Zi = X[rng.permutation(n)[:m]].copy()
    for approx in (g.approx_laplace(), g.approx_ep()):
        gp = g.gp_init(cfs=g.cf_sexp(lscale=1.2), lik=g.lik_bernoulli(), method=g.method_fitc(inducing=Zi),
                       approx=approx)
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, y, ctx=ctx)
        o.gp_fit(gpo, X, y)
        assert gp.fit["iters"] == gpo.fit["iters"], approx
        assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"]), approx
        assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL, approx
        if approx.name == "ep":
            assert scaled_err(gp.fit["fvar"], gpo.fit["fvar"]) < 1e-7
        Xt = rng.normal(size=(30, 2))
        p = g.gp_pred(gp, Xt, var=True)
        po = o.gp_pred(gpo, Xt, var=True)
        assert scaled_err(p["mean"], po["mean"]) < TOL
        assert scaled_err(p["var"], po["var"]) < 1e-7


# ------------------------------------------------------------------------------ wider coverage
def test_pred_with_cfind_subset(ctx):
    """gp_pred(..., cfind=) predicts with a subset of the summed cfs but the full-model alpha / C_chol
    (R/gp_pred.R:206-207, R/cf.R:402-411)."""
    rng = np.random.default_rng(21)
    n = 300
    X = rng.normal(size=(n, 2))
    y = np.sin(2 * X[:, 0]) + 0.5 * X[:, 1] + 0.2 * rng.normal(size=n)
    cfs = [g.cf_sexp(vars=[0], lscale=0.7), g.cf_lin(vars=[1], magn=0.8), g.cf_const(magn=0.5)]
    gp = g.gp_init(cfs=cfs, lik=g.lik_gaussian(sigma=0.3))
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    Xt = rng.normal(size=(40, 2))
    for cfind in ([0], [1, 2], [0, 1, 2]):
        p = g.gp_pred(gp, Xt, var=True, cfind=cfind)
        po = o.gp_pred(gpo, Xt, var=True, cfind=cfind)
        assert scaled_err(p["mean"], po["mean"]) < TOL
        assert np.max(np.abs(p["var"] - po["var"])) < TOL * max(1e-2, np.max(np.abs(po["var"])))
    q = g.gp_pred(gp, Xt, quantiles=[0.05, 0.95])
    qo = o.gp_pred(gpo, Xt, quantiles=[0.05, 0.95])
    assert scaled_err(q["quantiles"], qo["quantiles"]) < 1e-7


def test_laplace_betabinom_and_ep_counts(ctx):
    rng = np.random.default_rng(31)
    n = 200
    X = rng.normal(size=(n, 2))
    f = X[:, 0] - 0.5 * X[:, 1]
    trials = rng.integers(2, 15, size=n).astype(float)
    yb = rng.binomial(trials.astype(int), 1 / (1 + np.exp(-f))).astype(float)
    yp = rng.poisson(np.exp(0.5 * f + 0.5)).astype(float)
    # beta-binomial is not log-concave everywhere: either both sides fit, or both refuse the bad-site branch
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_betabinom(phi=0.05))
    gpo = copy.deepcopy(gp)
    try:
        o.gp_fit(gpo, X, yb, trials=trials)
        ok = True
    except o.NonLogConcave:
        ok = False
    if ok:
        g.gp_fit(gp, X, yb, trials=trials, ctx=ctx)
        assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < 1e-7 * abs(gpo.fit["log_evidence"])
        assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < 1e-7
    else:
        with pytest.raises(g.GplError):
            g.gp_fit(gp, X, yb, trials=trials, ctx=ctx)
    for lik, y, extra in ((g.lik_binomial(), yb, dict(trials=trials)), (g.lik_poisson(), yp, dict())):
        gp = g.gp_init(cfs=g.cf_matern32(lscale=1.2), lik=lik, approx=g.approx_ep(damping=0.8))
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, y, ctx=ctx, **extra)
        o.gp_fit(gpo, X, y, **extra)
        assert gp.fit["iters"] == gpo.fit["iters"], lik
        assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"]), lik
        assert scaled_err(gp.fit["fvar"], gpo.fit["fvar"]) < TOL, lik


def test_gp_optim_runs_both_optimisers(ctx):
    """gp_optim (R/gp_optim.R:63-125): derivative-free search as in the reference, and the
    gradient-based option built on gpl_energy_grad; both must lower the energy to a similar optimum."""
    rng = np.random.default_rng(41)
    n = 150
    x = rng.normal(size=n)
    y = np.sin(3 * x) * np.exp(-np.abs(x)) + 0.1 * rng.normal(size=n)
    res = {}
    for method in ("Nelder-Mead", "L-BFGS-B"):
        gp = g.gp_init(cfs=g.cf_sexp(), lik=g.lik_gaussian())
        g.gp_fit(gp, x, y, ctx=ctx)
        e0 = g.gp_energy(gp)
        g.gp_optim(gp, x, y, verbose=False, warn=False, method=method, maxiter=300, ctx=ctx)
        res[method] = g.gp_energy(gp)
        assert res[method] < e0 - 1.0
    assert abs(res["Nelder-Mead"] - res["L-BFGS-B"]) < 0.05 * abs(res["L-BFGS-B"]) + 0.5


def test_energy_grad_at_scale_property(ctx):
    """Full-size style check without an O(n^3) CPU oracle: at n = 4096 the analytic gradient must
    match central differences of the library's own energy (size-independent property)."""
    rng = np.random.default_rng(2)
    n, d = 4096, 8
    X = rng.normal(size=(n, d))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    gp = g.gp_init(cfs=g.cf_matern52(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
    e, grad = g.gp_energy_grad(gp, X, y, ctx=ctx)
    p0 = g.get_param_vector(gp)
    h = 1e-4
    for k in range(len(p0)):
        es = []
        for sgn in (1, -1):
            p = p0.copy()
            p[k] += sgn * h
            g.set_param(gp, p)
            es.append(g.gp_energy_grad(gp, X, y, ctx=ctx)[0])
        num = (es[0] - es[1]) / (2 * h)
        assert abs(num - grad[k]) < 1e-5 * max(1.0, abs(grad[k])), (k, num, grad[k])
    g.set_param(gp, p0)


def test_gp_loo_laplace_and_ep(ctx):
    """gp_loo (R/gp_loo.R:73-133) for method_full: Laplace needs the posterior variances at the
    training inputs (GPU prediction path), EP reuses the cavity moments."""
    X, y = _classification(250, 2, 9)
    res = {}
    for approx in (g.approx_laplace(), g.approx_ep()):
        gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_bernoulli(), approx=approx)
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, y, ctx=ctx)
        o.gp_fit(gpo, X, y)
        l = g.gp_loo(gp, X, y)
        lo = o.gp_loo(gpo, X, y)
        assert abs(l["loo"] - lo["loo"]) < 1e-7 * abs(lo["loo"])
        assert scaled_err(l["loos"], lo["loos"]) < 1e-6
        res[approx.name] = l
    cmp = g.gp_compare(res["laplace"], res["ep"])
    assert cmp.shape == (2, 2) and np.min(np.abs(cmp[:, 0])) == 0.0


def test_full_size_config2_properties(ctx):
    """BASELINE config 2 at FULL size (n = 32768, d = 8, cf_matern52 + lik_gaussian): no O(n^3) CPU
    oracle can run here, so check size-independent properties -- the analytic gradient against a
    central difference of the library's own energy along a random direction, alpha against the
    defining equation (K + (jitter + sigma^2) I) alpha = y on sampled rows (K rows from the oracle),
    and predictions at training points reproducing y - sigma^2 alpha."""
    rng = np.random.default_rng(2)
    n, d = 32768, 8
    X = rng.normal(size=(n, d))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    gp = g.gp_init(cfs=g.cf_matern52(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
    e0, grad = g.gp_energy_grad(gp, X, y, ctx=ctx)
    p0 = g.get_param_vector(gp)
    u = rng.normal(size=len(p0))
    u /= np.linalg.norm(u)
    h = 1e-4
    es = []
    for sgn in (1, -1):
        g.set_param(gp, p0 + sgn * h * u)
        es.append(g.gp_energy_grad(gp, X, y, ctx=ctx)[0])
    g.set_param(gp, p0)
    num = (es[0] - es[1]) / (2 * h)
    assert abs(num - grad @ u) < 1e-5 * max(1.0, abs(grad @ u)), (num, grad @ u)
    g.gp_fit(gp, X, y, ctx=ctx)
    assert abs(-gp.fit["log_evidence"] - e0 + g.lpdf_prior(gp)) < 1e-9 * abs(e0)
    alpha = gp.fit["alpha"]
    idx = rng.choice(n, size=64, replace=False)
    Krows = o.cf_native_np("cf_matern52", X[idx], X, 1.3, 1.0)
    lhs = Krows @ alpha + (1e-6 + 0.25) * alpha[idx]
    assert np.max(np.abs(lhs - y[idx])) < 1e-8 * np.max(np.abs(y))
    p = g.gp_pred(gp, X[idx], var=True)
    assert np.max(np.abs(p["mean"] - (y[idx] - (0.25 + 1e-6) * alpha[idx]))) < 1e-8
    assert np.all(p["var"] > 0) and np.all(p["var"] < 1.0)


@pytest.mark.parametrize("seed,phi", [(44, 0.3), (41, 0.3), (42, 0.05), (37, 0.05), (7, 0.3)])
@pytest.mark.parametrize("approx", ["laplace", "ep"])
def test_non_log_concave_sites(ctx, approx, seed, phi):
    """pseudo-variances <= 0 (beta-binomial with gross outliers).  The reference
    refits with a pivoted LU (R/laplace.R:40-117, R/ep.R:50-138, R/gp_pred.R:220-223); the CUDA path
    eliminates the bad block instead.  Same evidence, alpha, fmean, fvar and predictions."""
    if seed == 37:
        rng = np.random.default_rng(seed)
        n = 120
        X = rng.normal(size=(n, 2))
        f = X[:, 0] - 0.5 * X[:, 1]
        trials = rng.integers(2, 15, size=n).astype(float)
        y = rng.binomial(trials.astype(int), 1 / (1 + np.exp(-f))).astype(float)
        idx = rng.choice(n, 6, replace=False)
        y[idx] = np.where(f[idx] > 0, 0.0, trials[idx])
    elif seed == 7:      # 133 bad sites: the small factors span more than one 128-block
        X, y, trials = outlier_counts(seed, n=1200, n_out=600)
    else:
        X, y, trials = outlier_counts(seed)
    mk = g.approx_laplace if approx == "laplace" else g.approx_ep
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_betabinom(phi=phi), approx=mk())
    gpo = copy.deepcopy(gp)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o.gp_fit(gpo, X, y, trials=trials)
        g.gp_fit(gp, X, y, trials=trials, ctx=ctx)
    assert gpo.fit.get("n_bad", 0) >= 1, "scenario no longer exercises the bad-site branch"
    assert gp.fit["n_bad"] == gpo.fit["n_bad"]
    assert gp.fit["iters"] == gpo.fit["iters"]
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
    assert scaled_err(gp.fit["alpha"], gpo.fit["alpha"]) < TOL
    assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
    if approx == "ep":
        assert scaled_err(gp.fit["fvar"], gpo.fit["fvar"]) < TOL
    Xt = np.random.default_rng(seed + 1).normal(size=(37, 2))
    p = g.gp_pred(gp, Xt, var=True)
    po = o.gp_pred(gpo, Xt, var=True)
    assert scaled_err(p["mean"], po["mean"]) < TOL
    assert scaled_err(p["var"], po["var"]) < TOL
    with pytest.raises(g.GplError):
        g.fit_member(gp, "C_chol")


@pytest.mark.parametrize("n,m,d", [(3000, 64, 3), (5000, 300, 6), (777, 5, 1), (2000, 130, 11)])
def test_kmeans_inducing_vs_oracle(ctx, n, m, d):
    """gpl_kmeans (device Lloyd) against the numpy restatement from the same initial rows."""
    from gplite_b200.gp import kmeans_centers
    rng = np.random.default_rng(n + m)
    X = np.concatenate([rng.normal(size=(n // 2, d)), rng.normal(size=(n - n // 2, d)) * 0.5 + 2.0])
    cen, idx = kmeans_centers(X, m, np.random.default_rng(5), ctx=ctx)
    ceno, labo, ito = o.kmeans_lloyd(X, idx, maxiter=10)
    assert scaled_err(cen, ceno) < 1e-10
    # every centre is the mean of a non-empty cluster after at least one update
    assert ito >= 2 and np.all(np.isfinite(cen))


def test_method_fitc_default_inducing(ctx):
    """method_fitc() without explicit inducing points: k-means selection (R/gp_fit.R:190-234), kept on
    the model for later fits (R/gp_fit.R:112), also per bin along one input."""
    rng = np.random.default_rng(3)
    n = 1500
    X = rng.uniform(-3, 3, size=(n, 2))
    y = np.sin(X[:, 0]) * np.cos(0.5 * X[:, 1]) + 0.2 * rng.normal(size=n)
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_gaussian(sigma=0.3), method=g.method_fitc(num_inducing=40))
    g.gp_fit(gp, X, y, ctx=ctx)
    z1 = gp.method.inducing.copy()
    assert z1.shape == (40, 2) and np.isfinite(gp.fit["log_evidence"])
    assert np.all(z1.min(axis=0) >= X.min(axis=0)) and np.all(z1.max(axis=0) <= X.max(axis=0))
    # the oracle, given the same inducing points, agrees on the fit
    gpo = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_gaussian(sigma=0.3), method=g.method_fitc(inducing=z1))
    o.gp_fit(gpo, X, y)
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
    g.gp_fit(gp, X, y + 0.1, ctx=ctx)
    assert np.array_equal(gp.method.inducing, z1)                 # chosen once
    p = g.gp_pred(gp, X[:50], var=True)
    assert np.sqrt(np.mean((p["mean"] - 0.1 - np.sin(X[:50, 0]) * np.cos(0.5 * X[:50, 1])) ** 2)) < 0.2
    gpb = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_gaussian(sigma=0.3),
                    method=g.method_fitc(num_inducing=23, bin_along=0, bin_count=5))
    g.gp_fit(gpb, X, y, ctx=ctx)
    zb = gpb.method.inducing
    assert zb.shape == (23, 2) and np.isfinite(gpb.fit["log_evidence"])
    cut = np.linspace(X[:, 0].min(), X[:, 0].max(), 6)
    per_bin = np.histogram(zb[:, 0], bins=cut)[0]
    assert sorted(per_bin.tolist()) == [4, 4, 5, 5, 5]


def test_gp_draw_and_joint_covariance(ctx):
    """gp_draw (prior, posterior joint and marginal) and gp_pred_post(cov = TRUE) for
    method_full against the oracle with the same standard normal variates; also on a fit with
    non-log-concave sites (the reference's C_lu prediction branch)."""
    rng = np.random.default_rng(12)
    n, nt, draws = 300, 57, 4
    X = rng.normal(size=(n, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.2 * rng.normal(size=n)
    Xt = rng.normal(size=(nt, 2))
    Zn = rng.standard_normal((nt, draws))
    cfs = [g.cf_matern52(lscale=0.9), g.cf_lin(magn=0.4)]
    gp = g.gp_init(cfs=cfs, lik=g.lik_gaussian(sigma=0.3))
    gpo = copy.deepcopy(gp)
    # prior draw (unfitted model)
    s0 = g.gp_draw(gp, Xt, draws=draws, normals=Zn, transform=False, ctx=ctx)
    assert scaled_err(s0, o.gp_draw_latent(gpo, Xt, Zn)) < 1e-7         # chol of a jitter-regularised matrix
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    pc = g.gp_pred_cov(gp, Xt)
    pco = o.gp_pred_post_full(gpo, Xt, cov=True)
    assert scaled_err(pc["mean"], pco["mean"]) < TOL and scaled_err(pc["cov"], pco["cov"]) < TOL
    s1 = g.gp_draw(gp, Xt, draws=draws, normals=Zn, transform=False)
    assert scaled_err(s1, o.gp_draw_latent(gpo, Xt, Zn)) < 1e-6
    s2 = g.gp_draw(gp, Xt, draws=draws, normals=Zn, marginal=True, cfind=[0])
    assert scaled_err(s2, o.gp_draw_latent(gpo, Xt, Zn, marginal=True, cfind=[0])) < TOL
    # seeded draws are reproducible, targets have the likelihood's noise
    a = g.gp_draw(gp, Xt, draws=3, seed=5)
    b = g.gp_draw(gp, Xt, draws=3, seed=5)
    assert np.array_equal(a, b) and a.shape == (nt, 3)
    t = g.gp_draw(gp, Xt, draws=200, seed=6, target=True)
    f = g.gp_draw(gp, Xt, draws=200, seed=6)
    assert abs(np.std(t - f) - 0.3) < 0.03
    # classification + bad sites
    Xc, yc, trials = outlier_counts(44)
    gpc = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_betabinom(phi=0.3))
    gpco = copy.deepcopy(gpc)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.gp_fit(gpc, Xc, yc, trials=trials, ctx=ctx)
        o.gp_fit(gpco, Xc, yc, trials=trials)
    assert gpc.fit["n_bad"] > 0
    pc = g.gp_pred_cov(gpc, Xt)
    pco = o.gp_pred_post_full(gpco, Xt, cov=True)
    assert scaled_err(pc["cov"], pco["cov"]) < TOL
    pr = g.gp_draw(gpc, Xt, draws=2, seed=1)
    assert np.all((pr > 0) & (pr < 1))
    tr = g.gp_draw(gpc, Xt, draws=2, seed=1, target=True, trials=np.full(nt, 7.0))
    assert np.all((tr >= 0) & (tr <= 7) & (tr == np.round(tr)))


def _rf_models():
    return [
        ("sexp", lambda: [g.cf_sexp(lscale=0.8, magn=1.2)], 60),
        ("sexp+lin+const", lambda: [g.cf_sexp(lscale=1.1, vars=[0, 2]), g.cf_lin(magn=0.5), g.cf_const(magn=0.7)], 40),
        ("nn", lambda: [g.cf_nn(sigma0=0.8, sigma=1.5, magn=1.1)], 50),
        ("periodic+sexp*nn", lambda: [g.cf_periodic(period=2.0, cf_base=g.cf_sexp(lscale=0.9), vars=[1]),
                                      g.cf_sexp(lscale=1.4) * g.cf_nn(sigma=1.0, prior_magn=g.prior_fixed())], [30, 44]),
    ]


@pytest.mark.parametrize("name", [m[0] for m in _rf_models()])
@pytest.mark.parametrize("lik", ["gaussian", "bernoulli"])
def test_method_rf(ctx, name, lik):
    """method_rf (random-feature GP): feature maps, weight-space Laplace, prediction with
    cfind, against the oracle with the same random draws (numpy's generator in place of R's)."""
    _, mk, nb = next(m for m in _rf_models() if m[0] == name)
    rng = np.random.default_rng(17)
    n, nt = 700, 45
    X = rng.normal(size=(n, 3))
    f = np.sin(X[:, 0]) + 0.5 * X[:, 1] - 0.3 * X[:, 2]
    Xt = rng.normal(size=(nt, 3))
    if lik == "gaussian":
        y, L = f + 0.2 * rng.normal(size=n), g.lik_gaussian(sigma=0.3)
    else:
        y, L = (rng.uniform(size=n) < 1 / (1 + np.exp(-2 * f))).astype(float), g.lik_bernoulli()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gp = g.gp_init(cfs=mk(), lik=L, method=g.method_rf(num_basis=nb, seed=7))
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    assert gp.fit["iters"] == gpo.fit["iters"]
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < TOL * abs(gpo.fit["log_evidence"])
    assert scaled_err(gp.fit["wmean"], gpo.fit["wmean"]) < 1e-7
    assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < TOL
    assert scaled_err(g.fit_member(gp, "wcov"), gpo.fit["wcov"]) < 1e-7
    assert abs(g.gp_energy(gp) - o.gp_energy(gpo)) < TOL * abs(o.gp_energy(gpo))
    for cfind in (None, [0]) if len(gp.cfs) > 1 else (None,):
        p = g.gp_pred(gp, Xt, var=True, cfind=cfind)
        po = o.gp_pred(gpo, Xt, var=True, cfind=cfind)
        assert scaled_err(p["mean"], po["mean"]) < 1e-7, cfind
        assert scaled_err(p["var"], po["var"]) < 1e-7, cfind


def test_method_rf_rejections(ctx):
    X = np.random.default_rng(0).normal(size=(50, 2))
    y = X[:, 0]
    with pytest.raises(NotImplementedError):
        g.gp_fit(g.gp_init(cfs=g.cf_matern32(), method=g.method_rf(num_basis=10)), X, y, ctx=ctx)
    with pytest.raises(ValueError):
        g.gp_fit(g.gp_init(cfs=g.cf_sexp(), method=g.method_rf(num_basis=11)), X, y, ctx=ctx)
    with pytest.raises(g.GplError):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g.gp_fit(g.gp_init(cfs=g.cf_sexp() * g.cf_lin(), method=g.method_rf(num_basis=10)), X, y, ctx=ctx)


@pytest.mark.parametrize("seed", [44, 41, 42])
@pytest.mark.parametrize("approx", ["laplace", "ep"])
def test_non_log_concave_sites_fitc(ctx, approx, seed):
    """SURVEY 8f-2, FITC variants (R/laplace.R:153-232, R/ep.R:186-272): evidence split over good /
    bad sites in the m-dimensional space of Kxz Lz^-T, final fit by the inversion lemma with S of
    mixed sign."""
    X, y, trials = outlier_counts(seed)
    Zi = X[np.random.default_rng(seed).permutation(len(X))[:30]].copy()
    mk = g.approx_laplace if approx == "laplace" else g.approx_ep
    gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_betabinom(phi=0.3), approx=mk(),
                   method=g.method_fitc(inducing=Zi))
    gpo = copy.deepcopy(gp)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o.gp_fit(gpo, X, y, trials=trials)
        g.gp_fit(gp, X, y, trials=trials, ctx=ctx)
    assert gpo.fit.get("n_bad", 0) >= 1, "scenario no longer exercises the bad-site branch"
    assert gp.fit["n_bad"] == gpo.fit["n_bad"]
    assert gp.fit["iters"] == gpo.fit["iters"]
    assert abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) < 1e-7 * abs(gpo.fit["log_evidence"])
    assert scaled_err(gp.fit["alpha"], gpo.fit["alpha"]) < 1e-6
    assert scaled_err(gp.fit["fmean"], gpo.fit["fmean"]) < 1e-7
    if approx == "ep":
        assert scaled_err(gp.fit["fvar"], gpo.fit["fvar"]) < 1e-7
    Xt = np.random.default_rng(seed + 1).normal(size=(37, 2))
    p = g.gp_pred(gp, Xt, var=True)
    po = o.gp_pred(gpo, Xt, var=True)
    assert scaled_err(p["mean"], po["mean"]) < 1e-7
    assert scaled_err(p["var"], po["var"]) < 1e-7


@pytest.mark.parametrize("method", ["fitc", "rf"])
def test_gp_loo_other_methods(ctx, method):
    """gp_loo dispatches on the approximation only (R/gp_loo.R:111-133): the Laplace branch predicts at
    the training points with train = TRUE, which for FITC carries the diagonal correction
    (R/gp_pred.R:253-256,280-286) -- here the fit's own marginal variances."""
    X, y = _classification(300, 2, 10)
    Zi = X[np.random.default_rng(1).permutation(300)[:25]].copy()
    meth = g.method_fitc(inducing=Zi) if method == "fitc" else g.method_rf(num_basis=40, seed=2)
    for approx in ((g.approx_laplace(), g.approx_ep()) if method == "fitc" else (g.approx_laplace(),)):
        gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_bernoulli(), approx=approx, method=meth)
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, y, ctx=ctx)
        o.gp_fit(gpo, X, y)
        l = g.gp_loo(gp, X, y)
        lo = o.gp_loo(gpo, X, y)
        assert abs(l["loo"] - lo["loo"]) < 1e-7 * abs(lo["loo"]), (method, approx)
        assert scaled_err(l["loos"], lo["loos"]) < 1e-6, (method, approx)


def test_gp_save_load_roundtrip(ctx, tmp_path):
    """R/saveload.R: a fitted model survives gp_save / gp_load and predicts the same."""
    rng = np.random.default_rng(4)
    X = rng.normal(size=(200, 2))
    y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=200)
    gp = g.gp_init(cfs=g.cf_matern32(lscale=0.8), lik=g.lik_gaussian(sigma=0.2), method=g.method_fitc(num_inducing=20))
    g.gp_fit(gp, X, y, ctx=ctx)
    p0 = g.gp_pred(gp, X[:30], var=True)
    fn = str(tmp_path / "model.pkl")
    g.gp_save(gp, fn)
    gp2 = g.gp_load(fn, ctx=ctx)
    assert gp2.fitted and np.array_equal(gp2.method.inducing, gp.method.inducing)
    assert abs(gp2.fit["log_evidence"] - gp.fit["log_evidence"]) < 1e-9 * abs(gp.fit["log_evidence"])
    p1 = g.gp_pred(gp2, X[:30], var=True)
    assert scaled_err(p1["mean"], p0["mean"]) < 1e-10 and scaled_err(p1["var"], p0["var"]) < 1e-10
    fn2 = str(tmp_path / "unfitted.pkl")
    g.gp_save(g.gp_init(cfs=g.cf_sexp()), fn2)
    assert not g.gp_load(fn2).fitted


def test_fitc_joint_covariance_and_draws(ctx):
    """gp_pred_post.method_fitc(cov = TRUE) (R/gp_pred.R:262-273) and joint gp_draw for FITC: the device
    uses the collapsed form diag(Dt) + (Ktz L^-T)(Ktz L^-T)^T, the oracle the reference's expression."""
    X, y = _classification(400, 2, 12)
    Zi = X[np.random.default_rng(3).permutation(400)[:35]].copy()
    Xt = np.random.default_rng(4).normal(size=(41, 2))
    Zn = np.random.default_rng(5).standard_normal((41, 3))
    for lik, yy in ((g.lik_gaussian(sigma=0.4), X[:, 0] + 0.3 * y), (g.lik_bernoulli(), y)):
        gp = g.gp_init(cfs=[g.cf_sexp(lscale=1.0), g.cf_lin(magn=0.3)], lik=lik, method=g.method_fitc(inducing=Zi))
        gpo = copy.deepcopy(gp)
        g.gp_fit(gp, X, yy, ctx=ctx)
        o.gp_fit(gpo, X, yy)
        pc = g.gp_pred_cov(gp, Xt)
        pco = o.gp_pred_post_fitc(gpo, Xt, cov=True)
        assert scaled_err(pc["mean"], pco["mean"]) < 1e-7 and scaled_err(pc["cov"], pco["cov"]) < 1e-7
        s1 = g.gp_draw(gp, Xt, draws=3, normals=Zn, transform=False)
        assert scaled_err(s1, o.gp_draw_latent(gpo, Xt, Zn)) < 1e-5


def test_gp_draw_method_rf(ctx):
    """gp_draw_analytic.method_rf (R/gp_draw.R:145-163): weights drawn from N(wmean, wcov), pushed
    through the test-point features (gpl_rf_features)."""
    rng = np.random.default_rng(21)
    X = rng.normal(size=(400, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.1 * rng.normal(size=400)
    Xt = rng.normal(size=(33, 2))
    gp = g.gp_init(cfs=[g.cf_sexp(lscale=0.9), g.cf_lin(magn=0.4)], lik=g.lik_gaussian(sigma=0.2),
                   method=g.method_rf(num_basis=30, seed=3))
    gpo = copy.deepcopy(gp)
    g.gp_fit(gp, X, y, ctx=ctx)
    o.gp_fit(gpo, X, y)
    for cfind, msel in ((None, 32), ([0], 30)):
        Zn = rng.standard_normal((msel, 4))
        s1 = g.gp_draw(gp, Xt, draws=4, normals=Zn, transform=False, cfind=cfind)
        s0 = o.gp_draw_latent(gpo, Xt, Zn, cfind=cfind)
        assert s1.shape == (33, 4) and scaled_err(s1, s0) < 1e-6, cfind
    assert g.gp_draw(gp, Xt, draws=2, seed=1).shape == (33, 2)


@pytest.mark.parametrize("lik", ["bernoulli", "poisson", "binomial"])
def test_laplace_energy_gradient(ctx, lik):
    """New capability: analytic energy gradient for approx_laplace with non-Gaussian likelihoods
    (gpl_energy_grad_laplace) against the oracle formula, and L-BFGS gp_optim on top of it."""
    rng = np.random.default_rng(5)
    n = 500
    X = rng.normal(size=(n, 3))
    f0 = 1.5 * X[:, 0] - X[:, 1]
    kw = {}
    if lik == "bernoulli":
        y, L = (rng.uniform(size=n) < 1 / (1 + np.exp(-f0))).astype(float), g.lik_bernoulli()
    elif lik == "poisson":
        y, L = rng.poisson(np.exp(0.5 * f0)).astype(float), g.lik_poisson()
    else:
        kw["trials"] = rng.integers(3, 9, size=n).astype(float)
        y, L = rng.binomial(kw["trials"].astype(int), 1 / (1 + np.exp(-f0))).astype(float), g.lik_binomial()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mk = lambda: g.gp_init(cfs=[g.cf_matern52(lscale=1.1, magn=1.2), g.cf_sexp(lscale=0.7, vars=[0]) *
                                    g.cf_lin(prior_magn=g.prior_fixed())], lik=L,
                               approx=g.approx_laplace(maxiter=100, tol=1e-11))
        e, gr = g.gp_energy_grad(mk(), X, y, ctx=ctx, **kw)
        eo, go = o.energy_grad_laplace_full(mk(), X, y, **kw)
    assert abs(e - eo) < TOL * abs(eo)
    # the third derivative is a central difference (h = 1e-4) of the reference's own forward-difference
    # curvature (eps = 1e-6, R/lik.R:449-454): its rounding noise (~1e-10 / 2h) is the floor of this comparison
    assert np.max(np.abs(gr - go)) < 1e-5 * max(1.0, np.max(np.abs(go))), (gr, go)
    if lik == "bernoulli":
        gp = g.gp_init(cfs=g.cf_sexp(lscale=0.4, magn=0.8), lik=L, approx=g.approx_laplace(maxiter=100, tol=1e-10))
        e0 = g.gp_energy_grad(copy.deepcopy(gp), X, y, ctx=ctx)[0]
        gp = g.gp_optim(gp, X, y, method="L-BFGS-B", verbose=False, warn=False, maxiter=30)
        assert g.gp_energy(gp) < e0 - 1.0


def test_analytic_prediction_matches_sampling(ctx):
    """tests/testthat/test_pred.R:125-142 and test_sparse.R:153-170: the analytic latent prediction
    agrees with the moments of gp_draw (the reference uses 2e5 draws and tol 1e-2) -- for method_full,
    method_fitc and method_rf, and the staleness guard of test_pred.R:85-118 on gp_draw."""
    rng = np.random.default_rng(1)
    n, nt, draws = 25, 30, 100000
    x = rng.normal(size=(n, 1))
    xt = np.linspace(-3, 3, nt)[:, None]
    f = 2 * x[:, 0] - 4
    trials = np.full(n, 6.0)
    cases = [(g.lik_gaussian(), f + 0.3 * rng.normal(size=n), {}),
             (g.lik_bernoulli(), (rng.uniform(size=n) < 0.5).astype(float), {}),
             (g.lik_binomial(link="probit"), rng.binomial(6, 0.4, size=n).astype(float), dict(trials=trials)),
             (g.lik_poisson(), rng.poisson(2.0, size=n).astype(float), {})]
    methods = [g.method_full(), g.method_fitc(inducing=x[:10].copy()), g.method_rf(num_basis=20, seed=4)]
    for lik, y, kw in cases:
        for meth in methods:
            gp = g.gp_init(cfs=g.cf_sexp(lscale=0.8), lik=lik, method=copy.deepcopy(meth))
            g.gp_fit(gp, x, y, ctx=ctx, **kw)
            pred = g.gp_pred(gp, xt, var=True)
            s = g.gp_draw(gp, xt, draws=draws, transform=False, seed=4321)
            scale = max(1.0, np.max(np.abs(pred["mean"])))
            assert np.max(np.abs(s.mean(axis=1) - pred["mean"])) < 2e-2 * scale, (lik, meth)
            assert np.max(np.abs(s.std(axis=1) - np.sqrt(pred["var"]))) < 2e-2 * max(1.0, np.sqrt(pred["var"]).max()), (lik, meth)
    p = g.get_param_vector(gp)
    p[0] = 0.0
    g.set_param(gp, p)
    with pytest.raises(RuntimeError):
        g.gp_draw(gp, xt, draws=1)
    with pytest.raises(RuntimeError):
        g.gp_pred(gp, xt, var=True)

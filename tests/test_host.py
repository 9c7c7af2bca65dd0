"""CPU: host-side API plumbing (constructors, parameter packing, priors, guards) against the
reference's semantics (tests/testthat/test_param.R, test_warnings.R)."""
import math
import warnings

import numpy as np
import pytest
import scipy.stats as sst

import gplite_b200 as g
from gplite_b200 import model as M
from helpers import model_grid


def test_param_roundtrip_and_names():
    """tests/testthat/test_param.R:73-85."""
    rng = np.random.default_rng(0)
    for cfs in model_grid():
        gp = g.gp_init(cfs=cfs, lik=g.lik_gaussian())
        p = g.get_param(gp)
        v = np.array(list(p.values())) + rng.normal(size=len(p))
        g.set_param(gp, v)
        assert np.allclose(g.get_param_vector(gp), v, rtol=0, atol=1e-14)
        assert not gp.fitted
    gp = g.gp_init(cfs=[g.cf_sexp(), g.cf_periodic(cf_base=g.cf_matern32())], lik=g.lik_gaussian())
    assert list(g.get_param(gp).keys()) == [
        "cf_sexp.lscale", "cf_sexp.magn", "cf_periodic.period", "cf_periodic.lscale",
        "cf_periodic.magn", "lik_gaussian.sigma"]


def test_fixed_params_excluded_and_defaults():
    cf = g.cf_sexp(prior_magn=g.prior_fixed())
    gp = g.gp_init(cfs=cf, lik=g.lik_bernoulli())
    assert list(g.get_param(gp).keys()) == ["cf_sexp.lscale"]
    assert g.cf_sexp().lscale == 0.3 and g.cf_nn().sigma == 3.0 and g.cf_nn().normalize
    assert g.lik_gaussian().sigma == 0.5
    a = g.approx_laplace()
    assert (a.maxiter, a.tol) == (30, 1e-4)
    e = g.approx_ep()
    assert (e.damping, e.quad_order, e.maxiter) == (0.9, 11, 100)
    assert g.method_fitc().num_inducing == 100 and g.method_fitc().seed == 12345


def test_product_flattening_and_warning():
    """tests/testthat/test_warnings.R: >1 free magnitude in a product warns."""
    with pytest.warns(UserWarning):
        p = g.cf_sexp() * g.cf_lin()
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        p2 = g.cf_sexp() * g.cf_lin(prior_magn=g.prior_fixed())
        p3 = p2 * g.cf_const(prior_magn=g.prior_fixed())
    assert [c.name for c in p3.cfs] == ["cf_sexp", "cf_lin", "cf_const"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        q = g.cf_matern32() * (g.cf_sexp() * g.cf_lin())
    assert [c.name for c in q.cfs] == ["cf_sexp", "cf_lin", "cf_matern32"]   # cf2 * cf1, R/cf.R:208-210


def test_ep_with_gaussian_rejected():
    with pytest.raises(ValueError):
        g.gp_init(lik=g.lik_gaussian(), approx=g.approx_ep())
    with pytest.raises(ValueError):
        g.approx_ep(damping=1.0)


def test_priors_match_r_formulas():
    for lp in (-1.3, 0.0, 0.7):
        ht = g.prior_half_t(df=3.0, scale=2.0)
        want = math.log(2) - math.log(2.0) + sst.t.logpdf(math.exp(lp) / 2.0, df=3.0) + lp
        assert abs(M.lpdf_prior_one(ht, lp) - want) < 1e-13
        ln = g.prior_lognormal(loc=0.2, scale=0.5)
        assert abs(M.lpdf_prior_one(ln, lp) - sst.norm.logpdf(lp, 0.2, 0.5)) < 1e-13
        for pr in (ht, ln):
            h = 1e-6
            num = (M.lpdf_prior_one(pr, lp + h) - M.lpdf_prior_one(pr, lp - h)) / (2 * h)
            assert abs(M.dlpdf_prior_one(pr, lp) - num) < 1e-7
    assert M.lpdf_prior_one(g.prior_logunif(), 3.0) == 0.0


def test_flatten_descriptor():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfs = [g.cf_sexp(vars=[1]) * g.cf_lin(vars=[0, 2]), g.cf_periodic(period=2.0), g.cf_const()]
    factors, where = M.flatten_cfs(cfs)
    assert [f["term"] for f in factors] == [0, 0, 1, 2]
    assert [f["kind"] for f in factors] == ["cf_sexp", "cf_lin", "cf_sexp", "cf_const"]
    assert factors[2]["periodic"] and factors[2]["period"] == 2.0 and not factors[2]["normalize"]
    per = cfs[1]
    assert where[(id(per), "period")] == (2, 3)
    assert where[(id(per.base), "lscale")] == (2, 0)


def test_unfitted_and_unsupported_paths_need_no_gpu():
    gp = g.gp_init(method=g.method_rf(), lik=g.lik_bernoulli(), approx=g.approx_ep())
    with pytest.raises(NotImplementedError):            # "No EP implementation for method_rf yet."
        g.gp_fit(gp, np.zeros((3, 1)), np.zeros(3), ctx=object())
    gp = g.gp_init(cfs=g.cf_matern52(), method=g.method_rf())
    with pytest.raises(NotImplementedError):            # no random features for Matern kernels
        g.gp_fit(gp, np.zeros((3, 1)), np.zeros(3), ctx=object())
    assert g.model.check_num_basis([g.cf_sexp(), g.cf_lin(), g.cf_const(), g.cf_lin(vars=[1, 2])], 50, 4) == [50, 4, 1, 2]
    w = g.model.rf_weights([g.cf_sexp(lscale=0.5), g.cf_nn(sigma0=2.0, sigma=3.0)], [6, 4], 2, seed=1)
    assert w.shape == (3 * 2 + 4 * 2 + 4,)
    nrm = np.random.default_rng(1).standard_normal(12)
    assert np.allclose(w[:6], nrm[:6] / 0.5) and np.allclose(w[6:14], nrm[:8]) and np.allclose(w[14:], nrm[8:12] * 6.0)
    gp = g.gp_init()
    with pytest.raises(RuntimeError):
        g.gp_energy(gp)


def test_models_own_copies_of_their_components():
    """R value semantics: a cf / lik object reused in two models, or twice in one model, is never
    aliased by set_param; products own copies of their members."""
    import gplite_b200 as g
    from gplite_b200 import model as M
    cf = g.cf_sexp(lscale=0.5)
    lik = g.lik_gaussian(sigma=0.3)
    gp1 = g.gp_init(cfs=cf, lik=lik)
    gp2 = g.gp_init(cfs=cf, lik=lik)
    g.set_param(gp1, g.get_param_vector(gp1) + 1.0)
    assert gp2.cfs[0].lscale == 0.5 and gp2.lik.sigma == 0.3 and cf.lscale == 0.5
    gp3 = g.gp_init(cfs=[cf, cf])                      # the same object twice: two independent summands
    assert len(g.get_param(gp3)) == 5
    p = g.get_param_vector(gp3)
    p[0] += 1.0
    g.set_param(gp3, p)
    assert gp3.cfs[0].lscale != gp3.cfs[1].lscale
    factors, where = M.flatten_cfs(gp3.cfs)
    assert len(set(where.values())) == 4               # every free cf parameter maps to its own (factor, slot)
    prod = cf * g.cf_lin(prior_magn=g.prior_fixed())
    prod.cfs[0].lscale = 9.0
    assert cf.lscale == 0.5


def test_cfind_validation_and_lost_handle():
    import copy
    import pytest
    import gplite_b200 as g
    from gplite_b200.gp import _mask_of
    assert _mask_of(None, 3) == 0 and _mask_of([0, 2], 3) == 5
    with pytest.raises(ValueError):
        _mask_of([], 3)                                # "no terms" must not silently mean "all terms"
    with pytest.raises(ValueError):
        _mask_of([3], 3)
    gp = g.gp_init()
    gp.fitted, gp.fit = True, dict(type="analytic", alpha=np.zeros(2), log_evidence=0.0)
    gp.x = np.zeros((2, 1))
    gp2 = copy.deepcopy(gp)                            # device handles do not travel
    with pytest.raises(RuntimeError, match="gp_fit again"):
        g.gp_pred(gp2, np.zeros((1, 1)))

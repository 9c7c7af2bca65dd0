"""Host-side model objects: the cf_*, lik_*, prior_*, method_*, approx_* constructors and the
hyper-parameter plumbing of gplite, kept name-for-name (R/cf.R:88-396, R/lik.R:42-182,
R/priors.R, R/method.R, R/approx.R, R/misc.R:33-81).  Pure host logic -- no numerics on n x n
objects happen here.

Differences forced by the host language: `vars` are 0-based column indices; objects are small
Python classes instead of S3 lists (attribute names equal the list names).
"""
from __future__ import annotations

import copy
import math
import warnings

import numpy as np


# ----------------------------------------------------------------------------- priors
class Prior:
    def __init__(self, name, **kw):
        self.name = name
        self.__dict__.update(kw)

    def __repr__(self):
        return self.name


def prior_fixed():
    return Prior("prior_fixed")


def prior_logunif():
    return Prior("prior_logunif")


def prior_lognormal(loc=0.0, scale=1.0):
    return Prior("prior_lognormal", loc=loc, scale=scale)


def prior_half_t(df=1.0, scale=1.0):
    return Prior("prior_half_t", df=df, scale=scale)


def _dt_log(x, df):
    return (math.lgamma((df + 1) / 2) - math.lgamma(df / 2) - 0.5 * math.log(df * math.pi)
            - (df + 1) / 2 * math.log1p(x * x / df))


def lpdf_prior_one(prior, param):
    """R/priors.R:96-112; `param` is log(theta)."""
    if prior.name in ("prior_fixed", "prior_logunif"):
        return 0.0
    if prior.name == "prior_half_t":
        theta = math.exp(param)
        return math.log(2) - math.log(prior.scale) + _dt_log(theta / prior.scale, prior.df) + param
    if prior.name == "prior_lognormal":
        r = (param - prior.loc) / prior.scale
        return -0.5 * math.log(2 * math.pi) - math.log(prior.scale) - 0.5 * r * r
    raise ValueError(prior.name)


def dlpdf_prior_one(prior, param):
    """d lpdf / d log(theta) (used only by the gradient path, which the reference lacks)."""
    if prior.name in ("prior_fixed", "prior_logunif"):
        return 0.0
    if prior.name == "prior_half_t":
        theta = math.exp(param)
        x = theta / prior.scale
        return -(prior.df + 1) * (x * x / prior.df) / (1 + x * x / prior.df) + 1.0
    if prior.name == "prior_lognormal":
        return -(param - prior.loc) / prior.scale ** 2
    raise ValueError(prior.name)


# ----------------------------------------------------------------------------- covariance functions
_CF_PARAMS = {"cf_const": ["magn"], "cf_lin": ["magn"], "cf_sexp": ["lscale", "magn"],
              "cf_matern32": ["lscale", "magn"], "cf_matern52": ["lscale", "magn"],
              "cf_nn": ["sigma0", "sigma", "magn"], "cf_periodic": ["period"]}
# slot of each parameter in gpl_cf_factor.p (include/gplite_b200.h); period is slot 3
_CF_SLOTS = {"cf_const": {"magn": 0}, "cf_lin": {"magn": 0}, "cf_sexp": {"lscale": 0, "magn": 1},
             "cf_matern32": {"lscale": 0, "magn": 1}, "cf_matern52": {"lscale": 0, "magn": 1},
             "cf_nn": {"sigma0": 0, "sigma": 1, "magn": 2}}


class CF:
    def __init__(self, name, **kw):
        self.name = name
        self.means = None
        self.scales = None
        self.__dict__.update(kw)

    def __mul__(self, other):                               # "*.cf", R/cf.R:196-213
        if self.name == "cf_prod":
            # R value semantics: the product owns copies of its members
            out = copy.copy(self)
            out.cfs = [copy.deepcopy(c) for c in self.cfs] + \
                [copy.deepcopy(c) for c in (other.cfs if other.name == "cf_prod" else [other])]
            check_if_overparametrized_magnitude(out)
            return out
        if other.name == "cf_prod":
            return other * self
        return cf_prod(self, other)

    def __repr__(self):
        if self.name == "cf_prod":
            return "cf_prod(" + ", ".join(repr(c) for c in self.cfs) + ")"
        names = _CF_PARAMS[self.name]
        return f"{self.name}(" + "; ".join(f"{n} = {getattr(self, n):.3g}" for n in names) + ")"


def _vars(v):
    if v is None:
        return None
    v = [int(i) for i in np.atleast_1d(v)]
    return v


def cf_const(magn=1.0, prior_magn=None):
    return CF("cf_const", vars=None, magn=magn, priors={"magn": prior_magn or prior_logunif()},
              normalize=False)


def cf_lin(vars=None, magn=1.0, prior_magn=None, normalize=False):
    return CF("cf_lin", vars=_vars(vars), magn=magn, priors={"magn": prior_magn or prior_logunif()},
              normalize=normalize)


def _stationary(name, vars, lscale, magn, prior_lscale, prior_magn, normalize):
    return CF(name, vars=_vars(vars), lscale=lscale, magn=magn,
              priors={"lscale": prior_lscale or prior_logunif(), "magn": prior_magn or prior_logunif()},
              normalize=normalize)


def cf_sexp(vars=None, lscale=0.3, magn=1.0, prior_lscale=None, prior_magn=None, normalize=False):
    return _stationary("cf_sexp", vars, lscale, magn, prior_lscale, prior_magn, normalize)


def cf_matern32(vars=None, lscale=0.3, magn=1.0, prior_lscale=None, prior_magn=None, normalize=False):
    return _stationary("cf_matern32", vars, lscale, magn, prior_lscale, prior_magn, normalize)


def cf_matern52(vars=None, lscale=0.3, magn=1.0, prior_lscale=None, prior_magn=None, normalize=False):
    return _stationary("cf_matern52", vars, lscale, magn, prior_lscale, prior_magn, normalize)


def cf_nn(vars=None, sigma0=1.0, sigma=3.0, magn=1.0, prior_sigma0=None, prior_sigma=None,
          prior_magn=None, normalize=True):
    return CF("cf_nn", vars=_vars(vars), sigma0=sigma0, sigma=sigma, magn=magn,
              priors={"sigma0": prior_sigma0 or prior_half_t(), "sigma": prior_sigma or prior_half_t(),
                      "magn": prior_magn or prior_logunif()}, normalize=normalize)


def cf_periodic(vars=None, period=1.0, cf_base=None, prior_period=None):
    base = copy.deepcopy(cf_base) if cf_base is not None else cf_sexp()
    base.normalize = False            # R/cf.R:180
    base.vars = None                  # R/cf.R:181
    return CF("cf_periodic", vars=_vars(vars), period=period, base=base,
              priors={"period": prior_period or prior_logunif()})


def cf_prod(*cfs):
    # members are copied (R lists have value semantics: reusing a cf object elsewhere must not alias)
    out = CF("cf_prod", cfs=[copy.deepcopy(c) for c in cfs])
    check_if_overparametrized_magnitude(out)
    return out


def check_if_overparametrized_magnitude(cf):
    """R/misc.R:58-81."""
    names = [n for n, _, _ in _named_free_params([cf])]
    if sum("magn" in n for n in names) > 1:
        warnings.warn(
            "Detected a product of covariance functions with more than one non-fixed magnitude "
            "parameters. This means the magnitude of the product kernel is unnecessarily "
            "overparametrized, which is likely to lead to to problems in hyperparameter "
            "optimization. Consider setting prior_magn = prior_fixed() for all except one of the "
            "kernels in the product.")


# ----------------------------------------------------------------------------- likelihoods
class LikObj:
    def __init__(self, name, **kw):
        self.name = name
        self.__dict__.update(kw)

    def __repr__(self):
        return self.name


def lik_gaussian(sigma=0.5, prior_sigma=None):
    return LikObj("lik_gaussian", sigma=sigma, link="identity",
                  priors={"sigma": prior_sigma or prior_logunif()})


def _check_link(link, allowed):
    if link not in allowed:
        raise ValueError(f"Unknown link: {link}")


def lik_bernoulli(link="logit"):
    _check_link(link, ("logit", "probit"))
    return LikObj("lik_bernoulli", link=link, priors={})


def lik_binomial(link="logit"):
    _check_link(link, ("logit", "probit"))
    return LikObj("lik_binomial", link=link, priors={})


def lik_betabinom(link="logit", phi=0.1, prior_phi=None):
    _check_link(link, ("logit", "probit"))
    return LikObj("lik_betabinom", link=link, phi=phi, priors={"phi": prior_phi or prior_logunif()})


def lik_poisson(link="log"):
    _check_link(link, ("log",))
    return LikObj("lik_poisson", link=link, priors={})


_LIK_PARAMS = {"lik_gaussian": ["sigma"], "lik_bernoulli": [], "lik_binomial": [],
               "lik_betabinom": ["phi"], "lik_poisson": []}


# ----------------------------------------------------------------------------- methods / approximations
class Method:
    def __init__(self, name, **kw):
        self.name = name
        self.__dict__.update(kw)

    def __repr__(self):
        return "method_" + self.name


def method_full():
    return Method("full")


def method_fitc(inducing=None, num_inducing=100, bin_along=None, bin_count=10, seed=12345):
    """R/method.R:105-127.  bin_along is a 0-based input column, like `vars` (1-based in R)."""
    if inducing is not None:
        inducing = np.asarray(inducing, dtype=np.float64)
        if inducing.ndim == 1:
            inducing = inducing[:, None]
        num_inducing = inducing.shape[0]
    return Method("fitc", seed=seed, inducing=inducing, num_inducing=num_inducing,
                  bin_along=bin_along, bin_count=min(bin_count, num_inducing))


def method_rf(num_basis=400, seed=12345):
    """R/method.R:129-133: random features (linearised GP).  num_basis: one number or one per cf."""
    return Method("rf", seed=seed, num_basis=num_basis)


def check_num_basis(cfs, num_basis, num_inputs):
    """R/gp.R:205-228."""
    nb = [int(num_basis)] * len(cfs) if np.isscalar(num_basis) else [int(v) for v in num_basis]
    if len(nb) != len(cfs):
        raise ValueError("The length of num_basis must match to the number of covariance functions.")
    for k, cf in enumerate(cfs):
        if cf.name == "cf_const":
            nb[k] = 1
        elif cf.name == "cf_lin":
            nb[k] = len(cf.vars) if cf.vars is not None else num_inputs
    return nb


def rf_weights(cfs, num_basis, num_inputs, seed):
    """The random draws of rf_featmap.* (R/cf.R:571-657), one block per factor in flatten_cfs order,
    laid out as gpl_rf_spec.weights expects.  Every covariance function re-seeds with the same seed
    in the reference (set.seed(seed) inside each rf_featmap), so every factor restarts the stream;
    numpy's default_rng replaces R's generator."""
    out = []
    for k, cf in enumerate(cfs):
        members = cf.cfs if cf.name == "cf_prod" else [cf]
        for m in members:
            base = m.base if m.name == "cf_periodic" else m
            nsel = len(m.vars) if m.vars is not None else num_inputs
            nf = 2 * nsel if m.name == "cf_periodic" else nsel
            rng = np.random.default_rng(seed)
            if base.name == "cf_sexp":
                if num_basis[k] % 2:
                    raise ValueError("number of features must be an even number.")
                mw = num_basis[k] // 2
                w = rng.standard_normal(mw * nf).reshape(mw, nf).T / base.lscale     # nf x mw
                out.append(np.asfortranarray(w).ravel(order="F"))
            elif base.name == "cf_nn":
                mw = num_basis[k]
                w = rng.standard_normal(mw * nf).reshape(mw, nf).T                   # nf x mw, * sigma via the inputs
                w0 = rng.standard_normal(mw) * base.sigma0 * base.sigma
                out.append(np.asfortranarray(w).ravel(order="F"))
                out.append(w0)
            elif base.name in ("cf_matern32", "cf_matern52"):
                raise NotImplementedError("Random Fourier features for Matern kernels not implemented yet.")
    return np.concatenate(out) if out else np.zeros(0)


class ApproxObj:
    def __init__(self, name, **kw):
        self.name = name
        self.__dict__.update(kw)

    def __repr__(self):
        return "approx_" + self.name


def approx_laplace(maxiter=30, tol=1e-4):
    return ApproxObj("laplace", maxiter=maxiter, tol=tol)


def approx_ep(damping=0.9, quad_order=11, maxiter=100):
    if damping <= 0 or damping >= 1:
        raise ValueError("Damping should be between 0 and 1 (exclusive)")
    return ApproxObj("ep", damping=damping, quad_order=quad_order, maxiter=maxiter)


# ----------------------------------------------------------------------------- parameters
def _named_free_params(objs):
    """[(name, owner, attr)] in get_param.gp order (R/gp.R:150-155, R/cf.R:319-351): cfs in list
    order, products flattened, cf_periodic = period then its base (renamed cf_periodic.*);
    prior_fixed parameters excluded (R/misc.R:33-44)."""
    out = []
    for o in objs:
        if o.name == "cf_prod":
            out += _named_free_params(o.cfs)
            continue
        names = _CF_PARAMS.get(o.name)
        if names is None:
            names = _LIK_PARAMS[o.name]
        for nm in names:
            if o.priors[nm].name != "prior_fixed":
                out.append((f"{o.name}.{nm}", o, nm))
        if o.name == "cf_periodic":
            for (_, owner, nm) in _named_free_params([o.base]):
                out.append((f"cf_periodic.{nm}", owner, nm))
    return out


def get_param_objs(cfs, lik):
    return _named_free_params(list(cfs) + [lik])


def learn_scales(cf, x):
    """R/cf.R:822-855."""
    if cf.name == "cf_prod":
        for m in cf.cfs:
            learn_scales(m, x)
        return
    if cf.name == "cf_periodic":
        return
    xs = x if cf.vars is None else x[:, cf.vars]
    cf.means = xs.mean(axis=0)
    cf.scales = xs.std(axis=0, ddof=1) if xs.shape[0] > 1 else np.full(xs.shape[1], np.nan)


def flatten_cfs(cfs):
    """gp$cfs -> list of factor dicts for the C ABI (gpl_cf_factor) plus, for every free
    parameter, its (factor index, slot)."""
    factors = []
    where = {}          # (id(owner), attr) -> (factor index, slot)
    for term, cf in enumerate(cfs):
        members = cf.cfs if cf.name == "cf_prod" else [cf]
        for m in members:
            base = m.base if m.name == "cf_periodic" else m
            if base.name in ("cf_prod", "cf_periodic"):
                raise ValueError("nested cf_prod / cf_periodic bases are not supported")
            names = _CF_PARAMS[base.name]
            f = dict(kind=base.name, term=term, periodic=(m.name == "cf_periodic"),
                     normalize=bool(getattr(m, "normalize", False)) and m.name != "cf_periodic",
                     vars=m.vars, means=m.means, scales=m.scales,
                     p=[getattr(base, nm) for nm in names],
                     period=getattr(m, "period", 1.0))
            idx = len(factors)
            factors.append(f)
            for nm in names:
                where[(id(base), nm)] = (idx, _CF_SLOTS[base.name][nm])
            if m.name == "cf_periodic":
                where[(id(m), "period")] = (idx, 3)
    return factors, where

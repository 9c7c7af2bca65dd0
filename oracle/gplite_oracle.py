"""CPU oracle for the gplite dense-GP hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*: a numpy/scipy fp64 restatement of the reference's
algorithm (jpiironen/gplite v0.13.0, R + Rcpp).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``gplite_b200/`` (the product) imports it, and the
product has no CPU fallback.

Pinning status
--------------
* K entries (``cf_native``): PINNED.  ``oracle/_ref/libgplite_ref_cf.so`` is the
  reference's own ``src/cf.cpp`` compiled verbatim (oracle/Makefile); the numpy
  restatement below is checked against it in ``tests/test_oracle.py``.
* Gaussian log-evidence / predictions: pinned to the reference's *own test formula*
  (tests/testthat/test_basic.R:64-132, textbook R&W expressions, tol 1e-3 there).
* Laplace (non-Gaussian), EP, FITC numerics and everything at the 1e-8 level:
  **parity unpinned** by the reference's tests (they hold no golden vectors and R is not
  installed here, so the R side cannot be run).  The restatement below follows the R
  sources line by line and is the pin.
* Gradient: does not exist in the reference (gp_optim is Nelder-Mead,
  R/gp_optim.R:94).  ``energy_grad`` is a derived analytic formula cross-checked by
  central differences of the restated energy.

Every function cites the reference file:line it follows.  Model objects are duck-typed:
anything with the attribute names of the R lists (``name``, ``vars``, ``lscale``,
``magn``, ``normalize``, ``means``, ``scales`` ...) works; the product's host-side
classes in ``gplite_b200`` use the same names.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import scipy.linalg as sla
import scipy.special as sps
import scipy.stats as sst

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "libgplite_ref_cf.so")
_ref_lib = None

_KIND_NATIVE = {"cf_sexp": 0, "cf_matern32": 1, "cf_matern52": 2, "cf_nn": 3}


def ref_available() -> bool:
    return os.path.exists(_REF_SO)


def _load_ref():
    global _ref_lib
    if _ref_lib is None:
        lib = ctypes.CDLL(_REF_SO)
        lib.gplite_ref_cf.restype = ctypes.c_int
        lib.gplite_ref_cf.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
            ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            ctypes.c_void_p]
        _ref_lib = lib
    return _ref_lib


# --------------------------------------------------------------------------------------
# native covariance loops  (src/cf.cpp)
# --------------------------------------------------------------------------------------

def cf_native_ref(kind: str, x1, x2, p1, p2, p3=0.0, diag_only=False):
    """The reference's compiled src/cf.cpp (cf_sexp_c :9, cf_matern32_c :48,
    cf_matern52_c :86, cf_nn_c :125) through oracle/ref_driver.cpp."""
    lib = _load_ref()
    x1 = np.asfortranarray(np.atleast_2d(np.asarray(x1, dtype=np.float64)))
    x2 = np.asfortranarray(np.atleast_2d(np.asarray(x2, dtype=np.float64)))
    n1, d = x1.shape
    n2 = x2.shape[0]
    if diag_only:
        out = np.empty(min(n1, n2), dtype=np.float64)
    else:
        out = np.empty((n1, n2), dtype=np.float64, order="F")
    rc = lib.gplite_ref_cf(_KIND_NATIVE[kind], x1.ctypes.data, n1, x2.ctypes.data, n2, d,
                           float(p1), float(p2), float(p3), int(bool(diag_only)),
                           out.ctypes.data)
    if rc != 0:
        raise RuntimeError("gplite_ref_cf failed")
    return out


def cf_native_np(kind: str, x1, x2, p1, p2, p3=0.0, diag_only=False):
    """numpy restatement of the four loops of src/cf.cpp.

    sexp   :37-38   magn^2 exp(-0.5 r2 / lscale^2)
    m32    :75-76   magn^2 (1 + sqrt3 r/l) exp(-sqrt3 r/l)
    m52    :114-115 magn^2 (1 + sqrt5 r/l + 5 r^2/(3 l^2)) exp(-sqrt5 r/l)
    nn     :149-158 magn^2 (2/pi) asin(2a / sqrt((1+2b)(1+2c)))
    diag_only (:17-25 etc.) pairs row i of x1 with row i of x2, i < min(n1, n2).
    The squared distance is accumulated feature by feature, left to right, like the
    reference's sum(square(row_i - row_j)).
    """
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = np.atleast_2d(np.asarray(x2, dtype=np.float64))
    if diag_only:
        n = min(x1.shape[0], x2.shape[0])
        a, b = x1[:n], x2[:n]
        diff_sq = lambda k: (a[:, k] - b[:, k]) ** 2
        dot = lambda k: (a[:, k] * p2) * (b[:, k] * p2)
        nb = lambda k: (a[:, k] * p2) ** 2
        nc = lambda k: (b[:, k] * p2) ** 2
        shape = (n,)
    else:
        a, b = x1, x2
        diff_sq = lambda k: (a[:, k][:, None] - b[:, k][None, :]) ** 2
        dot = lambda k: (a[:, k] * p2)[:, None] * (b[:, k] * p2)[None, :]
        nb = lambda k: ((a[:, k] * p2) ** 2)[:, None]
        nc = lambda k: ((b[:, k] * p2) ** 2)[None, :]
        shape = (a.shape[0], b.shape[0])
    d = x1.shape[1]
    if kind == "cf_nn":
        sigma0, sigma, magn = p1, p2, p3
        s0sq = sigma0 * sigma0 * sigma * sigma
        sa = np.zeros(shape)
        sb = np.zeros(shape if diag_only else (shape[0], 1))
        sc = np.zeros(shape if diag_only else (1, shape[1]))
        for k in range(d):
            sa = sa + dot(k)
            sb = sb + nb(k)
            sc = sc + nc(k)
        aa = s0sq + sa
        bb = s0sq + sb
        cc = s0sq + sc
        return (magn * magn) * (2.0 / math.pi) * np.arcsin(
            2 * aa / np.sqrt((1 + 2 * bb) * (1 + 2 * cc)))
    lscale, magn = p1, p2
    r2 = np.zeros(shape)
    for k in range(d):
        r2 = r2 + diff_sq(k)
    magn2 = magn * magn
    if kind == "cf_sexp":
        return magn2 * np.exp(-0.5 * r2 / (lscale * lscale))
    r = np.sqrt(r2)
    if kind == "cf_matern32":
        return magn2 * (1.0 + math.sqrt(3.0) * r / lscale) * np.exp(-math.sqrt(3.0) * r / lscale)
    if kind == "cf_matern52":
        return magn2 * (1.0 + math.sqrt(5.0) * r / lscale + 5.0 * r * r / (3.0 * lscale * lscale)) \
            * np.exp(-math.sqrt(5.0) * r / lscale)
    raise ValueError(kind)


def cf_native(kind, x1, x2, p1, p2, p3=0.0, diag_only=False, use_ref=None):
    """Reference binary when built and the problem is small enough for its scalar loop,
    numpy restatement otherwise (the two are checked against each other in tests)."""
    if use_ref is None:
        n1 = np.atleast_2d(x1).shape[0]
        n2 = np.atleast_2d(x2).shape[0]
        use_ref = ref_available() and (n1 * n2 <= 4_000_000 or diag_only and min(n1, n2) <= 4_000_000)
    f = cf_native_ref if use_ref else cf_native_np
    return f(kind, x1, x2, p1, p2, p3, diag_only)


# --------------------------------------------------------------------------------------
# input preparation + eval_cf composition  (R/misc.R, R/cf.R)
# --------------------------------------------------------------------------------------

def _as_matrix(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]          # as.matrix(vector) is a column  (R/misc.R:21)
    return x


def learn_scales_cf(cf, x):
    """R/cf.R:836-845: column means and stats::sd (n-1) of the selected columns.
    cf_prod recurses into members (:847-850); cf_periodic learns nothing (:852-855)."""
    if cf.name == "cf_prod":
        for m in cf.cfs:
            learn_scales_cf(m, x)
        return cf
    if cf.name == "cf_periodic":
        return cf
    x = _as_matrix(x)
    if getattr(cf, "vars", None) is not None:
        x = x[:, list(cf.vars)]
    cf.means = x.mean(axis=0)
    cf.scales = x.std(axis=0, ddof=1) if x.shape[0] > 1 else np.full(x.shape[1], np.nan)
    return cf


def prepare_inputmat(cf, x):
    """R/misc.R:17-31: column subset `vars`, then (x - means) / scales iff normalize and
    the scales have been learned."""
    x = _as_matrix(x)
    if getattr(cf, "vars", None) is not None:
        x = x[:, list(cf.vars)]
    if getattr(cf, "normalize", False):
        means, scales = getattr(cf, "means", None), getattr(cf, "scales", None)
        if means is not None and scales is not None:
            x = (x - np.asarray(means)[None, :]) / np.asarray(scales)[None, :]
    return x


def eval_cf(obj, x1, x2, cfind=None, diag_only=False):
    """R/cf.R:402-497.  `obj` is a list of cfs (summed, :402-411, with optional cfind
    subset, 0-based here) or one cf."""
    if isinstance(obj, (list, tuple)):
        idx = range(len(obj)) if cfind is None else cfind
        K = 0.0
        for k in idx:
            K = K + eval_cf(obj[k], x1, x2, diag_only=diag_only)
        return K
    cf = obj
    name = cf.name
    if name == "cf_prod":                                         # :491-497
        K = 1.0
        for m in cf.cfs:
            K = K * eval_cf(m, x1, x2, diag_only=diag_only)
        return K
    a = prepare_inputmat(cf, x1)
    b = prepare_inputmat(cf, x2)
    n1, n2 = a.shape[0], b.shape[0]
    if name == "cf_const":                                        # :413-424
        return np.full(min(n1, n2), cf.magn ** 2) if diag_only else np.full((n1, n2), cf.magn ** 2)
    if name == "cf_lin":                                          # :426-435
        if diag_only:
            n = min(n1, n2)
            return cf.magn ** 2 * np.einsum("ij,ij->i", a[:n], b[:n])
        return cf.magn ** 2 * (a @ b.T)
    if name in ("cf_sexp", "cf_matern32", "cf_matern52"):         # :437-465
        return cf_native(name, a, b, cf.lscale, cf.magn, diag_only=diag_only)
    if name == "cf_nn":                                           # :467-476
        return cf_native(name, a, b, cf.sigma0, cf.sigma, cf.magn, diag_only=diag_only)
    if name == "cf_periodic":                                     # :478-489
        w = 2 * math.pi / cf.period
        at = np.concatenate([np.sin(w * a), np.cos(w * a)], axis=1)
        bt = np.concatenate([np.sin(w * b), np.cos(w * b)], axis=1)
        return eval_cf(cf.base, at, bt, diag_only=diag_only)
    raise ValueError(f"unknown cf {name}")


# --------------------------------------------------------------------------------------
# likelihoods  (R/lik.R)
# --------------------------------------------------------------------------------------

def _offset(f, offset):
    return f if offset is None else f + offset                     # R/misc.R:6-15


def get_response(lik, f):
    """R/lik.R:458-498."""
    if lik.name == "lik_gaussian":
        return f
    if lik.name in ("lik_bernoulli", "lik_binomial", "lik_betabinom"):
        if lik.link == "probit":
            return sst.norm.cdf(f)
        if lik.link == "logit":
            return 1.0 / (1.0 + np.exp(-f))
        raise ValueError(lik.link)
    if lik.name == "lik_poisson":
        return np.exp(f)
    raise ValueError(lik.name)


def _trials(lik, y, trials):
    if lik.name == "lik_bernoulli":
        return np.ones_like(np.asarray(y, dtype=np.float64))      # R/lik.R:311
    if trials is None:
        raise ValueError("trials must be provided")               # R/lik.R:316-318
    return np.asarray(trials, dtype=np.float64)


def _dbinom_log(y, t, mu):
    # stats::dbinom(y, t, mu, log=TRUE); same value as R's saddle-point form up to rounding
    with np.errstate(divide="ignore", invalid="ignore"):
        lchoose = sps.gammaln(t + 1) - sps.gammaln(y + 1) - sps.gammaln(t - y + 1)
        return lchoose + sps.xlogy(y, mu) + sps.xlog1py(t - y, -mu)


def get_loglik(lik, f, y, sum=True, trials=None, offset=None):
    """R/lik.R:297-366.  `f` may be a matrix (n x q) with y broadcast along rows."""
    f = _offset(np.asarray(f, dtype=np.float64), offset if np.ndim(f) == 1 or offset is None
                else np.asarray(offset)[:, None])
    y = np.asarray(y, dtype=np.float64)
    yb = y if f.ndim == 1 else y[:, None]
    if lik.name == "lik_gaussian":
        s = lik.sigma
        ll = -0.5 * math.log(2 * math.pi) - math.log(s) - 0.5 * ((yb - f) / s) ** 2
    elif lik.name in ("lik_bernoulli", "lik_binomial"):
        t = _trials(lik, y, trials)
        tb = t if f.ndim == 1 else t[:, None]
        ll = _dbinom_log(yb, tb, get_response(lik, f))
    elif lik.name == "lik_betabinom":
        t = _trials(lik, y, trials)
        tb = t if f.ndim == 1 else t[:, None]
        mu = get_response(lik, f)
        a = mu / lik.phi
        b = (1 - mu) / lik.phi
        g = sps.gammaln
        ll = (g(tb + 1) - g(yb + 1) - g(tb - yb + 1)) + (g(yb + a) + g(tb - yb + b) - g(tb + a + b)) \
            + (g(a + b) - g(a) - g(b))
    elif lik.name == "lik_poisson":
        mu = get_response(lik, f)
        ll = sps.xlogy(yb, mu) - mu - sps.gammaln(yb + 1)
    else:
        raise ValueError(lik.name)
    return float(np.sum(ll)) if sum else ll


def get_loglik_d(lik, f, y, trials=None, offset=None):
    """R/lik.R:370-443."""
    f = _offset(np.asarray(f, dtype=np.float64), offset)
    y = np.asarray(y, dtype=np.float64)
    mu = get_response(lik, f)
    if lik.name in ("lik_bernoulli", "lik_binomial", "lik_betabinom"):
        t = _trials(lik, y, trials)
        dmu = sst.norm.pdf(f) if lik.link == "probit" else mu * (1 - mu)
        if lik.name == "lik_betabinom":
            phi = lik.phi
            a = mu / phi
            b = (1 - mu) / phi
            term2 = (dmu / phi) * (sps.digamma(y + a) - sps.digamma(t - y + b))
            term3 = (dmu / phi) * (-sps.digamma(a) + sps.digamma(b))
            return term2 + term3
        return dmu * (y / mu - (t - y) / (1 - mu))               # :396
    if lik.name == "lik_poisson":
        return (y / mu - 1) * mu                                  # :441
    raise ValueError(lik.name)


def get_loglik_d2(lik, f, y, eps=1e-6, trials=None, offset=None):
    """R/lik.R:449-454: forward finite difference of the gradient, eps = 1e-6."""
    grad = get_loglik_d(lik, f, y, trials=trials, offset=offset)
    grad_eps = get_loglik_d(lik, f + eps, y, trials=trials, offset=offset)
    return grad, (grad_eps - grad) / eps


def get_pseudodata_la(lik, f, y, trials=None, offset=None, min_abs_curvature=1e-12):
    """R/lik.R:188-201."""
    y = np.asarray(y, dtype=np.float64)
    if lik.name == "lik_gaussian":
        return y.copy(), np.full(len(y), lik.sigma ** 2)
    f = np.asarray(f, dtype=np.float64).ravel()
    grad, grad2 = get_loglik_d2(lik, f, y, trials=trials, offset=offset)
    grad2 = np.sign(grad2) * np.maximum(np.abs(grad2), min_abs_curvature)
    return f - grad / grad2, -1.0 / grad2


# --------------------------------------------------------------------------------------
# quadrature  (R/quadrature.R) and logsumexp (R/gp_loo.R:136-144)
# --------------------------------------------------------------------------------------

def gauss_hermite_points(p: int):
    """R/quadrature.R:28-34.  Nodes = roots of probabilists' He_p (R: polyroot; here the
    Golub-Welsch nodes, ascending), weights = p! / (p^2 He_{p-1}(x)^2) -- they sum to 1."""
    x, _ = np.polynomial.hermite_e.hermegauss(p)
    hm2, hm1 = np.ones_like(x), x.copy()                          # He_0, He_1
    for k in range(1, p - 1):
        hm2, hm1 = hm1, x * hm1 - k * hm2                         # He_{k+1} = x He_k - k He_{k-1}
    he_pm1 = hm1 if p >= 2 else hm2
    w = math.factorial(p) / (p ** 2 * he_pm1 ** 2)
    return x, w


def logsumexp_w(x, weights):
    c = np.max(x, axis=-1, keepdims=True)
    return (c + np.log(np.sum(weights * np.exp(x - c), axis=-1, keepdims=True)))[..., 0]


def get_tilted_moments(lik, post_mean, post_prec, z_old, P_old, y, quad_order=11,
                       trials=None, offset=None):
    """R/lik.R:237-270."""
    cavity_prec = post_prec - P_old
    cavity_mean = 1 / cavity_prec * (post_mean * post_prec - z_old * P_old)
    xq, wq = gauss_hermite_points(quad_order)
    fgrid = xq[None, :] * np.sqrt(1 / cavity_prec)[:, None] + cavity_mean[:, None]
    loglik = get_loglik(lik, fgrid, y, sum=False, trials=trials, offset=offset)
    lik_v = np.exp(loglik)
    quad_ok = True
    if np.any(np.isnan(lik_v)):
        bad = np.isnan(lik_v)
        lik_v[bad] = 0.0
        loglik[np.isnan(loglik)] = -np.inf
        quad_ok = False
    log_C = logsumexp_w(loglik, wq)
    C = np.exp(log_C)
    mean_tilted = (1 / C) * ((fgrid * lik_v) @ wq)
    s2_tilted = (1 / C) * (((fgrid - mean_tilted[:, None]) ** 2 * lik_v) @ wq)
    return dict(mean_tilted=mean_tilted, var_tilted=s2_tilted, log_normalizing_const=log_C,
                mean_cavity=cavity_mean, var_cavity=1 / cavity_prec, quad_ok=quad_ok)


def get_pseudodata_ep(lik, post_mean, post_prec, z_old, P_old, y, quad_order=11, damping=0.5,
                      trials=None, offset=None):
    """R/lik.R:207-232."""
    tm = get_tilted_moments(lik, post_mean, post_prec, z_old, P_old, y, quad_order,
                            trials=trials, offset=offset)
    P_new = 1 / tm["var_tilted"] - 1 / tm["var_cavity"]
    nu_new = tm["mean_tilted"] / tm["var_tilted"] - tm["mean_cavity"] / tm["var_cavity"]
    nu_old = z_old * P_old
    nu_new = damping * nu_new + (1 - damping) * nu_old
    P_new = damping * P_new + (1 - damping) * P_old
    return dict(z=(1 / P_new) * nu_new, var=1 / P_new,
                log_normalizing_const=tm["log_normalizing_const"],
                mean_cavity=tm["mean_cavity"], var_cavity=tm["var_cavity"], quad_ok=tm["quad_ok"])


# --------------------------------------------------------------------------------------
# linear algebra helpers  (R/linear_algebra.R)
# --------------------------------------------------------------------------------------

def _chol_lower(A):
    return sla.cholesky(A, lower=True, check_finite=False)        # t(chol(A))


def _fsolve(L, b):
    return sla.solve_triangular(L, b, lower=True, check_finite=False)     # forwardsolve(L, b)


def _bsolve_t(L, b):
    return sla.solve_triangular(L, b, lower=True, trans="T", check_finite=False)  # backsolve(t(L), b)


def inv_lemma_get(K, U, D, logdet=True):
    """R/linear_algebra.R:3-43 with V = t(U)."""
    D = np.asarray(D, dtype=np.float64).ravel()
    inv_D_U = U / D[:, None]
    V_inv_D_U = U.T @ inv_D_U
    V_inv_D_U = 0.5 * (V_inv_D_U + V_inv_D_U.T)
    L = _chol_lower(K + V_inv_D_U)
    ld = None
    if logdet:
        ld = np.sum(np.log(D)) - 2 * np.sum(np.log(np.diag(_chol_lower(K)))) \
            + 2 * np.sum(np.log(np.diag(L)))
    return dict(chol=L, D=D, inv_D_U=inv_D_U, logdet=ld)


def inv_lemma_solve(obj, rhs, lhs=None, rhs_diag=False, lhs_diag=False, diag_only=False):
    """R/linear_algebra.R:45-126 (all implemented variants)."""
    D, inv_D_U, L = obj["D"], obj["inv_D_U"], obj["chol"]
    mid = lambda aux: _bsolve_t(L, _fsolve(L, aux))
    if lhs is not None:
        if rhs_diag:
            aux_right = inv_D_U.T * rhs[None, :]
            if lhs_diag:
                if diag_only:
                    return lhs / D * rhs - np.sum((lhs[:, None] * inv_D_U).T * mid(aux_right), axis=0)
                return np.diag(lhs / D * rhs) - (lhs[:, None] * inv_D_U) @ mid(aux_right)
            return lhs * (rhs / D)[None, :] - (lhs @ inv_D_U) @ mid(aux_right)
        aux_right = inv_D_U.T @ rhs
        if lhs_diag:
            return (lhs / D)[:, None] * rhs - (lhs[:, None] * inv_D_U) @ mid(aux_right)
        return lhs @ (rhs / D[:, None]) - (lhs @ inv_D_U) @ mid(aux_right)
    if rhs_diag:
        aux_right = inv_D_U.T * rhs[None, :]
        return np.diag(rhs / D) - inv_D_U @ mid(aux_right)
    aux_right = inv_D_U.T @ rhs
    aux_left = rhs / D[:, None] if rhs.ndim == 2 else rhs / D
    return aux_left - inv_D_U @ mid(aux_right)


# --------------------------------------------------------------------------------------
# Laplace  (R/laplace.R)   -- V > 0 branch only (log-concave likelihoods)
# --------------------------------------------------------------------------------------

class NonLogConcave(RuntimeError):
    pass


def laplace_iter_full(gp, K, y, fhat_old, trials=None, offset=None):
    """R/laplace.R:13-38,119."""
    z, V = get_pseudodata_la(gp.lik, fhat_old, y, trials=trials, offset=offset)
    n = len(z)
    if not np.all(V > 0):
        return _laplace_iter_full_bad(gp, K, y, fhat_old, z, V, trials, offset)
    return _laplace_iter_full_good(gp, K, y, z, V, trials, offset)


def _sub(v, idx):
    return None if v is None else np.asarray(v, dtype=np.float64)[idx]


def _lu_solve_full(K, V, rhs):
    """alpha = solve(U, solve(L, solve(P, rhs))) with Matrix::lu(K + diag(V)) (partial pivoting),
    R/laplace.R:113-114, R/ep.R:128-129."""
    lu, piv = sla.lu_factor(K + np.diag(V))
    return sla.lu_solve((lu, piv), rhs), (lu, piv)


def _bad_site_common(K, z, V, good, bad, fit1):
    """Shared linear algebra of R/laplace.R:63-83 and R/ep.R:81-101: predictive 'prior' of the bad
    sites given the good ones, and the factor of A = I + Lk^T diag(1/V_bad) Lk."""
    K21 = K[np.ix_(bad, good)]
    K22 = K[np.ix_(bad, bad)]
    pred_mean = K21 @ fit1["alpha"]
    aux = _fsolve(fit1["C_chol"], K21.T)
    pred_cov = K22 - aux.T @ aux
    Lk = _chol_lower(pred_cov)                       # failure = improper posterior: error propagates
    A = np.eye(len(bad)) + Lk.T @ ((1 / V[bad])[:, None] * Lk)
    try:
        A_chol = _chol_lower(A)
    except (np.linalg.LinAlgError, sla.LinAlgError):
        A_chol = None
    if A_chol is not None and not np.all(np.isfinite(A_chol)):
        A_chol = None
    return pred_mean, pred_cov, Lk, A_chol


def _laplace_iter_full_bad(gp, K, y, fhat_old, z, V, trials, offset):
    """R/laplace.R:40-117 (non-log-concave sites: p(y) = p(y_good) p(y_bad | y_good), LU for the fit)."""
    good = np.where(V > 0)[0]
    bad = np.where(~(V > 0))[0]
    fit1 = _laplace_iter_full_good(gp, K[np.ix_(good, good)], y[good], z[good], V[good],
                                   _sub(trials, good), _sub(offset, good))
    pred_mean, pred_cov, Lk, A_chol = _bad_site_common(K, z, V, good, bad, fit1)
    if A_chol is None:
        return dict(fmean=fhat_old, log_evidence=-np.inf)                      # :78-80
    Q = lambda v: _bsolve_t(Lk, _bsolve_t(A_chol, _fsolve(A_chol, Lk.T @ v)))
    alpha2 = Q(z[bad] / V[bad])                                                # :81
    aux = _bsolve_t(Lk, _fsolve(Lk, pred_mean))                                # :83
    aux = Q(aux)                                                               # :84
    f2hat = pred_cov @ (alpha2 + aux)                                          # :85
    invC2_V2_logdet = 2 * np.sum(np.log(np.diag(A_chol)))                      # :88
    df2 = np.sum(_fsolve(Lk, f2hat - pred_mean) ** 2)                          # :90
    log_lik2 = get_loglik(gp.lik, f2hat, y[bad], trials=_sub(trials, bad), offset=_sub(offset, bad))
    log_evidence = fit1["log_evidence"] + float(log_lik2 - 0.5 * df2 + 0.5 * invC2_V2_logdet)
    alpha, C_lu = _lu_solve_full(K, V, z)                                      # :113-114
    fhat_new = K @ alpha
    return dict(fmean=fhat_new, z=z, V=V, C_lu=C_lu, alpha=alpha, log_evidence=log_evidence,
                n_bad=len(bad))


def _laplace_iter_full_good(gp, K, y, z, V, trials, offset):
    """R/laplace.R:23-38 (all pseudo-variances positive)."""
    n = len(z)
    s = np.sqrt(1 / V)
    B = np.eye(n) + s[:, None] * (K * s[None, :])                 # :27
    C_chol = np.sqrt(V)[:, None] * _chol_lower(B)                 # :28
    alpha = _bsolve_t(C_chol, _fsolve(C_chol, z))                 # :29
    fhat_new = K @ alpha                                          # :30
    C_logdet = 2 * np.sum(np.log(np.diag(C_chol)))
    V_logdet = np.sum(np.log(V))
    f_invK_f = fhat_new @ alpha
    log_lik = get_loglik(gp.lik, fhat_new, y, trials=trials, offset=offset)
    log_evidence = log_lik - 0.5 * f_invK_f - 0.5 * (C_logdet - V_logdet)   # :37
    return dict(fmean=fhat_new, z=z, V=V, C_chol=C_chol, alpha=alpha,
                log_evidence=float(log_evidence))


def laplace_full(gp, K, y, maxiter=100, tol=1e-4, trials=None, offset=None):
    """R/laplace.R:274-297."""
    n = len(y)
    fhat = np.zeros(n)
    if gp.lik.name == "lik_gaussian":
        maxiter = 1
    it = 0
    for it in range(1, maxiter + 1):
        fit = laplace_iter_full(gp, K, y, fhat, trials=trials, offset=offset)
        diff = np.max(np.abs(fhat - fit["fmean"]))
        fhat = fit["fmean"]
        if diff < tol:
            break
    fit["iters"] = it
    fit["converged"] = not (maxiter > 1 and it == maxiter)
    return fit


def kmeans_lloyd(x, init_idx, maxiter=10):
    """Deterministic restatement of what gpl_kmeans computes (the reference calls stats::kmeans,
    R/gp_fit.R:221,228, whose Hartigan-Wong variant and RNG stream are R-only): Lloyd iterations
    from the given initial rows, ties to the lowest centre index, empty clusters keep their centre,
    stop when no assignment changes."""
    x = _as_matrix(x)
    cen = x[np.asarray(init_idx)].copy()
    lab = np.full(len(x), -1)
    it = 0
    for it in range(1, maxiter + 1):
        d2 = ((x[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2)
        new = np.argmin(d2, axis=1)
        changed = np.any(new != lab)
        lab = new
        if not changed:
            break
        for c in range(len(cen)):
            sel = lab == c
            if sel.any():
                cen[c] = x[sel].mean(axis=0)
    return cen, lab, it


def fitc_prepare(gp, x, z_ind, jitter):
    """R/gp_fit.R:105-110 (identical in :167-172)."""
    m = z_ind.shape[0]
    Kz = eval_cf(gp.cfs, z_ind, z_ind) + jitter * np.eye(m)
    Kxz = eval_cf(gp.cfs, x, z_ind)
    Kz_chol = _chol_lower(Kz)
    K_diag = eval_cf(gp.cfs, x, x, diag_only=True) + jitter
    D = K_diag - np.sum(_fsolve(Kz_chol, Kxz.T) ** 2, axis=0)
    return Kz, Kxz, Kz_chol, K_diag, D


def laplace_iter_fitc(gp, Kz, Kz_chol, Kxz, D, y, fhat_old, trials=None, offset=None):
    """R/laplace.R:122-152,234-237."""
    z, V = get_pseudodata_la(gp.lik, fhat_old, y, trials=trials, offset=offset)
    if not np.all(V > 0):
        return _laplace_iter_fitc_bad(gp, Kz, Kz_chol, Kxz, D, y, fhat_old, z, V, trials, offset)
    return _laplace_iter_fitc_good(gp, Kz, Kz_chol, Kxz, D, y, z, V, trials, offset)


def _fitc_bad_common(Kz_chol, Kxz, D, z, V, good, bad, fit1):
    """R/laplace.R:183-203 == R/ep.R:226-246."""
    Lb = _fsolve(Kz_chol, Kxz[bad].T)
    Lg = _fsolve(Kz_chol, Kxz[good].T)
    K21 = Lb.T @ Lg
    prior_cov = Lb.T @ Lb + np.diag(D[bad])
    pred_mean = K21 @ fit1["alpha"]
    pred_cov = prior_cov - K21 @ inv_lemma_solve(fit1["C_inv"], K21.T)
    Lk = _chol_lower(pred_cov)
    A = np.eye(len(bad)) + Lk.T @ ((1 / V[bad])[:, None] * Lk)
    try:
        A_chol = _chol_lower(A)
        if not np.all(np.isfinite(A_chol)):
            A_chol = None
    except (np.linalg.LinAlgError, sla.LinAlgError):
        A_chol = None
    return pred_mean, pred_cov, Lk, A_chol


def _laplace_iter_fitc_bad(gp, Kz, Kz_chol, Kxz, D, y, fhat_old, z, V, trials, offset):
    """R/laplace.R:153-232."""
    good = np.where(V > 0)[0]
    bad = np.where(~(V > 0))[0]
    fit1 = _laplace_iter_fitc_good(gp, Kz, Kz_chol, Kxz[good], D[good], y[good], z[good], V[good],
                                   _sub(trials, good), _sub(offset, good))
    pred_mean, pred_cov, Lk, A_chol = _fitc_bad_common(Kz_chol, Kxz, D, z, V, good, bad, fit1)
    if A_chol is None:
        return dict(fmean=fhat_old, log_evidence=-np.inf, log_post_unnorm=-np.inf)
    Q = lambda v: _bsolve_t(Lk, _bsolve_t(A_chol, _fsolve(A_chol, Lk.T @ v)))
    alpha2 = Q(z[bad] / V[bad])
    aux = Q(_bsolve_t(Lk, _fsolve(Lk, pred_mean)))
    f2hat = pred_cov @ (alpha2 + aux)
    invC2_V2_logdet = 2 * np.sum(np.log(np.diag(A_chol)))
    df2 = np.sum(_fsolve(Lk, f2hat - pred_mean) ** 2)
    log_lik2 = get_loglik(gp.lik, f2hat, y[bad], trials=_sub(trials, bad), offset=_sub(offset, bad))
    lpu = fit1["log_post_unnorm"] + float(log_lik2 - 0.5 * df2)
    le = fit1["log_evidence"] + float(log_lik2 - 0.5 * df2 + 0.5 * invC2_V2_logdet)
    S = D + V
    C_inv = inv_lemma_get(Kz, Kxz, S, logdet=False)                             # :225
    alpha = inv_lemma_solve(C_inv, z)
    fhat_new = Kxz @ _bsolve_t(Kz_chol, _fsolve(Kz_chol, Kxz.T @ alpha)) + D * alpha
    return dict(fmean=fhat_new, z=z, V=V, Kz=Kz, Kxz=Kxz, Kz_chol=Kz_chol, C_inv=C_inv, diag=D,
                alpha=alpha, log_evidence=le, log_post_unnorm=lpu, n_bad=len(bad))


def _laplace_iter_fitc_good(gp, Kz, Kz_chol, Kxz, D, y, z, V, trials, offset):
    """R/laplace.R:134-152 (all pseudo-variances positive)."""
    S = D + V
    C_inv = inv_lemma_get(Kz, Kxz, S)
    alpha = inv_lemma_solve(C_inv, z)
    fhat_new = Kxz @ _bsolve_t(Kz_chol, _fsolve(Kz_chol, Kxz.T @ alpha)) + D * alpha   # :141
    C_logdet = C_inv["logdet"]
    V_logdet = np.sum(np.log(V))
    f_invK_f = fhat_new @ alpha
    log_lik = get_loglik(gp.lik, fhat_new, y, trials=trials, offset=offset)
    log_evidence = log_lik - 0.5 * f_invK_f + 0.5 * (V_logdet - C_logdet)
    return dict(fmean=fhat_new, z=z, V=V, Kz=Kz, Kxz=Kxz, Kz_chol=Kz_chol, C_inv=C_inv, diag=D,
                alpha=alpha, log_evidence=float(log_evidence),
                log_post_unnorm=float(log_lik - 0.5 * f_invK_f))


def laplace_fitc(gp, Kz, Kz_chol, Kxz, D, y, maxiter=100, tol=1e-4, trials=None, offset=None):
    """R/laplace.R:299-325."""
    fhat = np.zeros(len(y))
    if gp.lik.name == "lik_gaussian":
        maxiter = 1
    log_post_old = -np.inf
    it = 0
    for it in range(1, maxiter + 1):
        fit = laplace_iter_fitc(gp, Kz, Kz_chol, Kxz, D, y, fhat, trials=trials, offset=offset)
        fhat = fit["fmean"]
        diff = fit["log_post_unnorm"] - log_post_old
        log_post_old = fit["log_post_unnorm"]
        if abs(diff) < tol:
            break
    fit["iters"] = it
    return fit


# --------------------------------------------------------------------------------------
# EP  (R/ep.R)   -- V > 0 branch only
# --------------------------------------------------------------------------------------

def ep_iter_full(gp, K, y, fmean_old, fvar_old, z_old, P_old, damping, quad_order=11,
                 trials=None, offset=None):
    """R/ep.R:16-49,140-145."""
    pobs = get_pseudodata_ep(gp.lik, fmean_old, 1 / fvar_old, z_old, P_old, y,
                             quad_order=quad_order, damping=damping, trials=trials, offset=offset)
    z, V = pobs["z"], pobs["var"]
    n = len(z)
    if np.any(np.isnan(V) | np.isnan(z)):
        return dict(fmean=None, log_evidence=-np.inf)
    if not np.all(V > 0):
        return _ep_iter_full_bad(gp, K, pobs)
    return _ep_iter_full_good(K, pobs)


def _ep_iter_full_bad(gp, K, pobs):
    """R/ep.R:50-138."""
    z, V = pobs["z"], pobs["var"]
    good = np.where(V > 0)[0]
    bad = np.where(~(V > 0))[0]
    sub = {k: (v[good] if isinstance(v, np.ndarray) else v) for k, v in pobs.items()}
    fit1 = _ep_iter_full_good(K[np.ix_(good, good)], sub)
    pred_mean, pred_cov, Lk, A_chol = _bad_site_common(K, z, V, good, bad, fit1)
    if A_chol is None:
        return dict(fmean=None, log_evidence=-np.inf)                          # :95-97
    Q = lambda v: _bsolve_t(Lk, _bsolve_t(A_chol, _fsolve(A_chol, Lk.T @ v)))
    invC2_V2_logdet = 2 * np.sum(np.log(np.diag(A_chol)))
    dz = pred_mean - z[bad]
    df2 = dz @ Q(dz / V[bad])                                                  # :107-110
    mc, vc, logZ = pobs["mean_cavity"], pobs["var_cavity"], pobs["log_normalizing_const"]
    log_evidence2 = -0.5 * df2 + 0.5 * invC2_V2_logdet + np.sum(logZ[bad]) \
        + 0.5 * np.sum(np.log((vc[bad] + V[bad]) / V[bad])) \
        + 0.5 * np.sum((mc[bad] - z[bad]) ** 2 / (vc[bad] + V[bad]))          # :112-114
    alpha, C_lu = _lu_solve_full(K, V, z)                                      # :122-123
    fmean_new = K @ alpha
    fcov_new = V[:, None] * sla.lu_solve(C_lu, K)                              # :125
    return dict(fmean=fmean_new, fvar=np.diag(fcov_new).copy(), cavity_mean=mc, cavity_var=vc,
                z=z, P=1 / V, C_lu=C_lu, alpha=alpha, quad_ok=pobs["quad_ok"],
                log_evidence=float(fit1["log_evidence"] + log_evidence2), n_bad=len(bad))


def _ep_iter_full_good(K, pobs):
    """R/ep.R:34-49,140-145 (all pseudo-variances positive)."""
    z, V = pobs["z"], pobs["var"]
    n = len(z)
    s = np.sqrt(1 / V)
    B = np.eye(n) + s[:, None] * (K * s[None, :])
    C_chol = np.sqrt(V)[:, None] * _chol_lower(B)
    alpha = _bsolve_t(C_chol, _fsolve(C_chol, z))
    fmean_new = K @ alpha
    fcov_new = V[:, None] * _bsolve_t(C_chol, _fsolve(C_chol, K))   # :42 (full n x n!)
    C_logdet = 2 * np.sum(np.log(np.diag(C_chol)))
    z_invC_z = z @ alpha
    mc, vc = pobs["mean_cavity"], pobs["var_cavity"]
    log_evidence = -0.5 * C_logdet - 0.5 * z_invC_z + np.sum(pobs["log_normalizing_const"]) \
        + 0.5 * np.sum(np.log(vc + V)) + 0.5 * np.sum((mc - z) ** 2 / (vc + V))      # :45-49
    return dict(fmean=fmean_new, fvar=np.diag(fcov_new).copy(), cavity_mean=mc, cavity_var=vc,
                z=z, P=1 / V, C_chol=C_chol, alpha=alpha, log_evidence=float(log_evidence),
                quad_ok=pobs["quad_ok"])


def _ep_outer(step, n, fvar0, maxiter, tol, damping, damping_min):
    """Shared control flow of R/ep.R:286-330 and :333-374."""
    P_old = np.zeros(n)
    z_old = np.zeros(n)
    fmean_old = np.zeros(n)
    fvar_old = fvar0
    diff_old = np.inf
    damp = damping
    it = 0
    accepted = 0
    history = []          # (diff_new, diff_old) per sweep: lets the tests see how decisive each accept / reject was
    for it in range(1, maxiter + 1):
        fit = step(fmean_old, fvar_old, z_old, P_old, damp)
        if fit["fmean"] is None:
            fit["iters"] = it
            return fit
        diff_new = np.max(np.abs(fmean_old - fit["fmean"]))
        history.append((float(diff_new), float(diff_old)))
        if diff_new > diff_old:
            damp = damp * damping
            continue
        damp = max(damp, damping_min)
        fmean_old, fvar_old, z_old, P_old = fit["fmean"], fit["fvar"], fit["z"], fit["P"]
        diff_old = diff_new
        accepted += 1
        if diff_new < tol:
            break
    if maxiter > 1 and it == maxiter:
        fit["log_evidence"] = -np.inf
    if not fit.get("quad_ok", True):
        fit["log_evidence"] = -np.inf
    fit["iters"] = it
    fit["accepted"] = accepted
    fit["damp"] = damp
    fit["history"] = history
    return fit


def ep_full(gp, K, y, maxiter=300, tol=1e-4, damping=0.9, damping_min=0.1, quad_order=11,
            trials=None, offset=None):
    """R/ep.R:286-330."""
    step = lambda fm, fv, z, P, damp: ep_iter_full(gp, K, y, fm, fv, z, P, damp, quad_order,
                                                   trials=trials, offset=offset)
    return _ep_outer(step, len(y), np.diag(K).copy(), maxiter, tol, damping, damping_min)


def ep_iter_fitc(gp, Kz, Kz_chol, Kxz, D, y, fmean_old, fvar_old, z_old, P_old, damping,
                 quad_order=11, trials=None, offset=None):
    """R/ep.R:148-185,274-279."""
    pobs = get_pseudodata_ep(gp.lik, fmean_old, 1 / fvar_old, z_old, P_old, y,
                             quad_order=quad_order, damping=damping, trials=trials, offset=offset)
    z, V = pobs["z"], pobs["var"]
    if np.any(np.isnan(V) | np.isnan(z)):
        return dict(fmean=None, log_evidence=-np.inf)
    if not np.all(V > 0):
        return _ep_iter_fitc_bad(gp, Kz, Kz_chol, Kxz, D, pobs)
    return _ep_iter_fitc_good(Kz, Kz_chol, Kxz, D, pobs)


def _ep_iter_fitc_bad(gp, Kz, Kz_chol, Kxz, D, pobs):
    """R/ep.R:186-272."""
    z, V = pobs["z"], pobs["var"]
    good = np.where(V > 0)[0]
    bad = np.where(~(V > 0))[0]
    sub = {k: (v[good] if isinstance(v, np.ndarray) else v) for k, v in pobs.items()}
    fit1 = _ep_iter_fitc_good(Kz, Kz_chol, Kxz[good], D[good], sub)
    pred_mean, pred_cov, Lk, A_chol = _fitc_bad_common(Kz_chol, Kxz, D, z, V, good, bad, fit1)
    if A_chol is None:
        return dict(fmean=None, log_evidence=-np.inf)
    Q = lambda v: _bsolve_t(Lk, _bsolve_t(A_chol, _fsolve(A_chol, Lk.T @ v)))
    invC2_V2_logdet = 2 * np.sum(np.log(np.diag(A_chol)))
    dz = pred_mean - z[bad]
    df2 = dz @ Q(dz / V[bad])
    mc, vc, logZ = pobs["mean_cavity"], pobs["var_cavity"], pobs["log_normalizing_const"]
    le2 = -0.5 * df2 + 0.5 * invC2_V2_logdet + np.sum(logZ[bad]) \
        + 0.5 * np.sum(np.log((vc[bad] + V[bad]) / V[bad])) \
        + 0.5 * np.sum((mc[bad] - z[bad]) ** 2 / (vc[bad] + V[bad]))
    S = D + V
    C_inv = inv_lemma_get(Kz, Kxz, S, logdet=False)                             # :259
    alpha = inv_lemma_solve(C_inv, z)
    inv_Kz_Kzx = _bsolve_t(Kz_chol, _fsolve(Kz_chol, Kxz.T))
    aux = inv_lemma_solve(C_inv, V, inv_Kz_Kzx, rhs_diag=True)
    diag1 = np.sum(Kxz.T * aux, axis=0)
    diag2 = inv_lemma_solve(C_inv, V, D, rhs_diag=True, lhs_diag=True, diag_only=True)
    fmean_new = Kxz @ _bsolve_t(Kz_chol, _fsolve(Kz_chol, Kxz.T @ alpha)) + D * alpha
    return dict(fmean=fmean_new, fvar=diag1 + diag2, cavity_mean=mc, cavity_var=vc, z=z, P=1 / V,
                Kz=Kz, Kxz=Kxz, Kz_chol=Kz_chol, C_inv=C_inv, diag=D, alpha=alpha,
                log_evidence=float(fit1["log_evidence"] + le2), quad_ok=pobs["quad_ok"], n_bad=len(bad))


def _ep_iter_fitc_good(Kz, Kz_chol, Kxz, D, pobs):
    """R/ep.R:169-185 (all pseudo-variances positive)."""
    z, V = pobs["z"], pobs["var"]
    S = D + V
    C_inv = inv_lemma_get(Kz, Kxz, S)
    alpha = inv_lemma_solve(C_inv, z)
    inv_Kz_Kzx = _bsolve_t(Kz_chol, _fsolve(Kz_chol, Kxz.T))
    aux = inv_lemma_solve(C_inv, V, inv_Kz_Kzx, rhs_diag=True)
    diag1 = np.sum(Kxz.T * aux, axis=0)
    diag2 = inv_lemma_solve(C_inv, V, D, rhs_diag=True, lhs_diag=True, diag_only=True)
    fvar_new = diag1 + diag2
    fmean_new = Kxz @ _bsolve_t(Kz_chol, _fsolve(Kz_chol, Kxz.T @ alpha)) + D * alpha
    C_logdet = C_inv["logdet"]
    z_invC_z = z @ alpha
    mc, vc = pobs["mean_cavity"], pobs["var_cavity"]
    log_evidence = -0.5 * C_logdet - 0.5 * z_invC_z + np.sum(pobs["log_normalizing_const"]) \
        + 0.5 * np.sum(np.log(vc + V)) + 0.5 * np.sum((mc - z) ** 2 / (vc + V))
    return dict(fmean=fmean_new, fvar=fvar_new, cavity_mean=mc, cavity_var=vc, z=z, P=1 / V,
                Kz=Kz, Kxz=Kxz, Kz_chol=Kz_chol, C_inv=C_inv, diag=D, alpha=alpha,
                log_evidence=float(log_evidence), quad_ok=pobs["quad_ok"])


def ep_fitc(gp, Kz, Kz_chol, Kxz, K_diag, D, y, damping=0.9, damping_min=0.1, maxiter=100,
            tol=1e-3, quad_order=11, trials=None, offset=None):
    """R/ep.R:333-374."""
    step = lambda fm, fv, z, P, damp: ep_iter_fitc(gp, Kz, Kz_chol, Kxz, D, y, fm, fv, z, P, damp,
                                                   quad_order, trials=trials, offset=offset)
    return _ep_outer(step, len(y), K_diag, maxiter, tol, damping, damping_min)


# --------------------------------------------------------------------------------------
# fit dispatch  (R/gp_fit.R)
# --------------------------------------------------------------------------------------

# --------------------------------------------------------------------------------------
# method_rf  (R/cf.R:532-690, R/gp.R:205-262, R/laplace.R:241-262,327-350, R/gp_pred.R:298-315)
# --------------------------------------------------------------------------------------
def _erf(t):
    from scipy.special import erf
    return erf(t)                      # == pnorm(t, sd = sqrt(.5)) - pnorm(-t, sd = sqrt(.5)), R/cf.R:646


def rf_featmap_one(cf, num_feat, num_inputs, normals):
    """rf_featmap.<class> for one covariance function; `normals(count)` replaces stats::rnorm after
    set.seed(seed) (every call restarts the same stream, as every rf_featmap re-seeds)."""
    if cf.name == "cf_const":
        return lambda x: cf.magn * np.ones((len(_as_matrix(x)), 1))              # R/cf.R:553-559
    if cf.name == "cf_lin":
        return lambda x: cf.magn * prepare_inputmat(cf, x)                       # :561-569
    if cf.name == "cf_sexp":
        nin = len(cf.vars) if cf.vars is not None else num_inputs
        if num_feat % 2:
            raise ValueError("number of features must be an even number.")
        m = num_feat // 2
        w = normals(m * nin).reshape(m, nin).T * (1 / cf.lscale)                 # :598-600 (column-major fill)

        def fm(x):
            h = prepare_inputmat(cf, x) @ w
            return cf.magn * np.sqrt(1 / m) * np.hstack([np.cos(h), np.sin(h)])  # :602-606
        return fm
    if cf.name == "cf_nn":
        nin = len(cf.vars) if cf.vars is not None else num_inputs
        m = num_feat
        stream = normals(m * nin + m)
        w = stream[:m * nin].reshape(m, nin).T * cf.sigma                        # :643
        w0 = stream[m * nin:][None, :] * cf.sigma0 * cf.sigma                    # :644
        w_aug = np.vstack([w0, w])

        def fm(x):
            xp = prepare_inputmat(cf, x)
            return cf.magn / np.sqrt(m) * _erf(np.hstack([np.ones((len(xp), 1)), xp]) @ w_aug)
        return fm
    if cf.name == "cf_periodic":
        nin = len(cf.vars) if cf.vars is not None else num_inputs
        base = rf_featmap_one(cf.base, num_feat, 2 * nin, normals)               # :667

        def fm(x):
            xp = prepare_inputmat(cf, x)
            return base(np.hstack([np.sin(2 * np.pi / cf.period * xp), np.cos(2 * np.pi / cf.period * xp)]))
        return fm
    if cf.name == "cf_prod":
        if any(m.name in ("cf_lin", "cf_const") for m in cf.cfs):
            raise ValueError("Random features for product kernel containing constant or linear kernel "
                             "not implemented yet.")
        maps = [rf_featmap_one(m, num_feat, num_inputs, normals) for m in cf.cfs]

        def fm(x):
            z = 1.0
            for f in maps:
                z = z * f(x)
            return z
        return fm
    raise NotImplementedError("Random Fourier features for Matern kernels not implemented yet.")


def rf_check_num_basis(cfs, num_basis, num_inputs):
    """R/gp.R:205-228."""
    nb = [int(num_basis)] * len(cfs) if np.isscalar(num_basis) else [int(v) for v in num_basis]
    for k, cf in enumerate(cfs):
        if cf.name == "cf_const":
            nb[k] = 1
        elif cf.name == "cf_lin":
            nb[k] = len(cf.vars) if cf.vars is not None else num_inputs
    return nb


def rf_featuremap(gp, num_inputs):
    nb = rf_check_num_basis(gp.cfs, gp.method.num_basis, num_inputs)
    normals = lambda count: np.random.default_rng(gp.method.seed).standard_normal(count)
    maps = [rf_featmap_one(cf, nb[k], num_inputs, normals) for k, cf in enumerate(gp.cfs)]

    def featuremap(x, cfind=None):                                               # R/cf.R:534-551
        idx = range(len(maps)) if cfind is None else cfind
        return np.hstack([maps[k](x) for k in idx])
    return featuremap, nb


def laplace_iter_rf(gp, Z, y, fhat_old, trials=None, offset=None):
    """R/laplace.R:241-262."""
    z, V = get_pseudodata_la(gp.lik, fhat_old, y, trials=trials, offset=offset)
    n, d = Z.shape
    H = Z.T @ (Z / V[:, None])
    L = _chol_lower(H + np.eye(d))
    what = _bsolve_t(L, _fsolve(L, Z.T @ (z / V)))
    wcov = _bsolve_t(L, _fsolve(L, np.eye(d)))
    wcov_logdet = -2 * np.sum(np.log(np.diag(L)))
    fhat_new = Z @ what
    log_prior = -0.5 * d * np.log(2 * np.pi) - 0.5 * np.sum(what ** 2)
    log_lik = get_loglik(gp.lik, fhat_new, y, trials=trials, offset=offset)
    log_evidence = 0.5 * n * np.log(2 * np.pi) + 0.5 * wcov_logdet + log_prior + log_lik
    return dict(fmean=fhat_new, wmean=what, wcov=wcov, z=z, V=V, log_evidence=float(log_evidence))


def laplace_rf(gp, Z, y, maxiter=100, tol=1e-4, trials=None, offset=None):
    """R/laplace.R:327-350."""
    fhat = np.zeros(len(y))
    if gp.lik.name == "lik_gaussian":
        maxiter = 1
    it = 0
    for it in range(1, maxiter + 1):
        fit = laplace_iter_rf(gp, Z, y, fhat, trials=trials, offset=offset)
        fnew = Z @ fit["wmean"]
        diff = np.max(np.abs(fhat - fnew))
        fhat = fnew
        if diff < tol:
            break
    fit["iters"] = it
    return fit


def gp_pred_post_rf(gp, xt, var=False, cfind=None):
    """R/gp_pred.R:298-315 with get_w_mean / get_w_cov (R/gp.R:230-262)."""
    xt = _as_matrix(xt)
    featuremap, nb = rf_featuremap(gp, xt.shape[1])
    zt = featuremap(xt, cfind)
    ends = np.concatenate([[0], np.cumsum(nb)])
    idx = range(len(nb)) if cfind is None else cfind
    inds = np.concatenate([np.arange(ends[k], ends[k + 1]) for k in idx])
    mean = zt @ gp.fit["wmean"][inds]
    if var:
        return dict(mean=mean, var=np.diag(zt @ (gp.fit["wcov"][np.ix_(inds, inds)] @ zt.T)).copy())
    return dict(mean=mean)


def gp_fit(gp, x, y, trials=None, offset=None, jitter=None):
    """R/gp_fit.R:54-60 -> learn_scales -> fit_model -> fit_{laplace,ep}.method_{full,fitc}.
    Mutates and returns `gp` (sets gp.x, gp.fit, gp.fitted)."""
    x = _as_matrix(x)
    y = np.asarray(y, dtype=np.float64)
    for cf in gp.cfs:
        learn_scales_cf(cf, x)
    jit = 1e-6 if jitter is None else jitter                      # R/gp.R:276-281
    n = len(y)
    meth, appr = gp.method.name, gp.approx.name
    if meth == "full":
        K = eval_cf(gp.cfs, x, x) + jit * np.eye(n)               # R/gp_fit.R:95,154
        if appr == "laplace":
            fit = laplace_full(gp, K, y, maxiter=gp.approx.maxiter, tol=gp.approx.tol,
                               trials=trials, offset=offset)
        else:
            fit = ep_full(gp, K, y, quad_order=gp.approx.quad_order, damping=gp.approx.damping,
                          damping_min=0.1, maxiter=gp.approx.maxiter, trials=trials, offset=offset)
    elif meth == "fitc":
        z_ind = _as_matrix(gp.method.inducing)
        Kz, Kxz, Kz_chol, K_diag, D = fitc_prepare(gp, x, z_ind, jit)
        try:
            if appr == "laplace":
                fit = laplace_fitc(gp, Kz, Kz_chol, Kxz, D, y, maxiter=gp.approx.maxiter,
                                   tol=gp.approx.tol, trials=trials, offset=offset)
            else:
                fit = ep_fitc(gp, Kz, Kz_chol, Kxz, K_diag, D, y, quad_order=gp.approx.quad_order,
                              damping=gp.approx.damping, damping_min=0.1,
                              maxiter=gp.approx.maxiter, trials=trials, offset=offset)
        except (np.linalg.LinAlgError, sla.LinAlgError):
            fit = dict(log_evidence=-np.inf)                      # R/gp_fit.R:113-121
    elif meth == "rf":
        featuremap, nb = rf_featuremap(gp, x.shape[1])                 # R/gp_fit.R:125-132
        fit = laplace_rf(gp, featuremap(x), y, maxiter=gp.approx.maxiter, tol=gp.approx.tol,
                         trials=trials, offset=offset)
    else:
        raise ValueError(meth)
    gp.x = x
    gp.fit = fit
    gp.fit["type"] = "analytic"
    gp.fitted = True
    return gp


# --------------------------------------------------------------------------------------
# priors + energy  (R/priors.R, R/gp.R:138-147)
# --------------------------------------------------------------------------------------

def lpdf_prior_one(prior, param):
    """R/priors.R:96-112; `param` is the log of the parameter."""
    if prior.name in ("prior_fixed", "prior_logunif"):
        return 0.0
    if prior.name == "prior_half_t":
        theta = math.exp(param)
        return math.log(2) - math.log(prior.scale) + sst.t.logpdf(theta / prior.scale, df=prior.df) + param
    if prior.name == "prior_lognormal":
        return sst.norm.logpdf(param, loc=prior.loc, scale=prior.scale)
    raise ValueError(prior.name)


_CF_PARAMS = {"cf_const": ["magn"], "cf_lin": ["magn"], "cf_sexp": ["lscale", "magn"],
              "cf_matern32": ["lscale", "magn"], "cf_matern52": ["lscale", "magn"],
              "cf_nn": ["sigma0", "sigma", "magn"], "cf_periodic": ["period"]}
_LIK_PARAMS = {"lik_gaussian": ["sigma"], "lik_bernoulli": [], "lik_binomial": [],
               "lik_betabinom": ["phi"], "lik_poisson": []}


def free_params(obj):
    """(owner, attribute) pairs in get_param.gp order (R/gp.R:150-155, R/cf.R:319-351):
    cfs in list order with products flattened, periodic = period then base; then lik.
    prior_fixed parameters are excluded (R/misc.R:33-44)."""
    out = []
    if isinstance(obj, (list, tuple)):
        for o in obj:
            out += free_params(o)
        return out
    if obj.name == "cf_prod":
        return free_params(obj.cfs)
    names = _CF_PARAMS.get(obj.name, _LIK_PARAMS.get(obj.name))
    for nm in names:
        if obj.priors[nm].name != "prior_fixed":
            out.append((obj, nm))
    if obj.name == "cf_periodic":
        out += free_params(obj.base)
    return out


def lpdf_prior(gp):
    """R/gp.R:167-169 etc."""
    lp = 0.0
    for owner, nm in free_params(list(gp.cfs) + [gp.lik]):
        lp += lpdf_prior_one(owner.priors[nm], math.log(getattr(owner, nm)))
    return lp


def gp_energy(gp, include_prior=True):
    """R/gp.R:138-147."""
    e = -gp.fit["log_evidence"]
    if include_prior:
        e -= lpdf_prior(gp)
    return e


# --------------------------------------------------------------------------------------
# prediction  (R/gp_pred.R)
# --------------------------------------------------------------------------------------

def gp_pred_post_full(gp, xt, var=False, cov=False, cfind=None, jitter=None):
    """R/gp_pred.R:202-232."""
    Kt = eval_cf(gp.cfs, xt, gp.x, cfind)
    Ktt = eval_cf(gp.cfs, xt, xt, cfind)
    alpha = gp.fit["alpha"]
    mean = Kt @ alpha
    if var or cov:
        nt = len(mean)
        jit = 1e-6 if jitter is None else jitter
        if gp.fit.get("C_chol") is not None:
            aux = _fsolve(gp.fit["C_chol"], Kt.T)
            pred_cov = Ktt - aux.T @ aux
        else:                                                     # :220-223 (LU of K + diag(V))
            pred_cov = Ktt - Kt @ sla.lu_solve(gp.fit["C_lu"], Kt.T)
        if cov:
            return dict(mean=mean, cov=pred_cov + jit * np.eye(nt))
        return dict(mean=mean, var=np.diag(pred_cov).copy())
    return dict(mean=mean)


def gp_draw_latent(gp, xt, normals, marginal=False, cfind=None, jitter=None):
    """Latent part of gp_draw (R/gp_draw.R:66-78,105-123; mvnrnd :183-207) for method_full with
    the standard normal variates supplied: mean + t(chol(cov)) %*% normals, or the marginal form
    mean + sqrt(var) * normals; prior (mean 0, K(xt, xt) + jitter I) when the model is not fitted."""
    xt = _as_matrix(xt)
    nt = len(xt)
    jit = 1e-6 if jitter is None else jitter
    if not getattr(gp, "fitted", False):
        cov = eval_cf(gp.cfs, xt, xt, cfind) + jit * np.eye(nt)          # R/gp_pred.R:142-156
        return _chol_lower(cov) @ normals
    if gp.method.name == "rf":                                    # R/gp_draw.R:145-163; normals: M_sel x draws
        featuremap, nb = rf_featuremap(gp, xt.shape[1])
        ends = np.concatenate([[0], np.cumsum(nb)])
        idx = range(len(nb)) if cfind is None else cfind
        inds = np.concatenate([np.arange(ends[k], ends[k + 1]) for k in idx])
        w = gp.fit["wmean"][inds][:, None] + _chol_lower(gp.fit["wcov"][np.ix_(inds, inds)]) @ normals
        return featuremap(xt, cfind) @ w
    post = gp_pred_post_full if gp.method.name == "full" else gp_pred_post_fitc
    if marginal:
        pred = post(gp, xt, var=True, cfind=cfind, jitter=jitter)
        return pred["mean"][:, None] + np.sqrt(pred["var"])[:, None] * normals
    pred = post(gp, xt, cov=True, cfind=cfind, jitter=jitter)
    return pred["mean"][:, None] + _chol_lower(pred["cov"]) @ normals


def gp_pred_post_fitc(gp, xt, var=False, cfind=None, jitter=None, train=False, cov=False):
    """R/gp_pred.R:234-296."""
    z_ind = _as_matrix(gp.method.inducing)
    alpha = gp.fit["alpha"]
    Ktz = eval_cf(gp.cfs, xt, z_ind, cfind)
    Kxz = gp.fit["Kxz"]
    Kz_chol = gp.fit["Kz_chol"]
    Dt = None
    if var or train or cov:
        Dt = eval_cf(gp.cfs, xt, xt, cfind, diag_only=True)
        Dt = Dt - np.sum(_fsolve(Kz_chol, Ktz.T) ** 2, axis=0)
    Kt_aux1 = _fsolve(Kz_chol, Kxz.T)
    Kt_aux2 = _fsolve(Kz_chol, Ktz.T)
    mean = Kt_aux2.T @ (Kt_aux1 @ alpha)
    if train:
        mean = mean + Dt * alpha
    if not (var or cov):
        return dict(mean=mean)
    C_inv = gp.fit["C_inv"]
    if cov:                                                       # :262-273 (train = FALSE branch)
        jit = 1e-6 if jitter is None else jitter
        Linv_Ktz = _fsolve(Kz_chol, Ktz.T)
        prior_cov = Linv_Ktz.T @ Linv_Ktz + np.diag(Dt)
        pred_cov = prior_cov - Kt_aux2.T @ (Kt_aux1 @ inv_lemma_solve(C_inv, Kt_aux1.T)) @ Kt_aux2
        return dict(mean=mean, cov=pred_cov + jit * np.eye(len(mean)))
    Kz_inv_Ktz = _bsolve_t(Kz_chol, _fsolve(Kz_chol, Ktz.T))
    prior_var = np.sum(Ktz * Kz_inv_Ktz.T, axis=1) + Dt
    if train:
        aux = inv_lemma_solve(C_inv, Kt_aux1.T, Kt_aux1) @ Kt_aux2 \
            + inv_lemma_solve(C_inv, Dt, Kt_aux1, rhs_diag=True)
        diag1 = np.sum(Kt_aux2 * aux, axis=0)
        diag2 = np.sum(inv_lemma_solve(C_inv, Kt_aux1.T, Dt, lhs_diag=True) * Kt_aux2.T, axis=1)
        diag3 = inv_lemma_solve(C_inv, Dt, Dt, rhs_diag=True, lhs_diag=True, diag_only=True)
        var_red = diag1 + diag2 + diag3
    else:
        aux = inv_lemma_solve(C_inv, Kt_aux1.T, Kt_aux1) @ Kt_aux2
        var_red = np.sum(Kt_aux2 * aux, axis=0)
    return dict(mean=mean, var=prior_var - var_red)


def gp_pred(gp, xnew, var=False, cfind=None, jitter=None, offset=None, transform=False,
            quantiles=None, quad_order=15):
    """R/gp_pred.R:88-122 (front end: prior if unfitted, offset, quantiles, transform)."""
    if quantiles is not None or transform:
        var = True
    xnew = _as_matrix(xnew)
    if not getattr(gp, "fitted", False) or getattr(gp, "fit", None) is None:
        nt = xnew.shape[0]
        pred = dict(mean=np.zeros(nt))                            # :142-156
        if var:
            pred["var"] = np.diag(eval_cf(gp.cfs, xnew, xnew, cfind)).copy()
    elif gp.method.name == "full":
        pred = gp_pred_post_full(gp, xnew, var=var, cfind=cfind, jitter=jitter)
    elif gp.method.name == "rf":
        pred = gp_pred_post_rf(gp, xnew, var=var, cfind=cfind)
    else:
        pred = gp_pred_post_fitc(gp, xnew, var=var, cfind=cfind, jitter=jitter)
    pred["mean"] = _offset(pred["mean"], offset)
    if quantiles is not None:
        q = np.stack([sst.norm.ppf(qq, loc=pred["mean"], scale=np.sqrt(pred["var"]))
                      for qq in quantiles], axis=1)
        pred["quantiles"] = get_response(gp.lik, q) if transform else q
    if transform:
        xq, wq = gauss_hermite_points(quad_order)
        fgrid = xq[None, :] * np.sqrt(pred["var"])[:, None] + pred["mean"][:, None]
        ft = get_response(gp.lik, fgrid)
        mt = ft @ wq
        pred["var"] = ((ft - mt[:, None]) ** 2) @ wq
        pred["mean"] = mt
    return pred


def gp_loo(gp, x, y, quad_order=11, trials=None, offset=None):
    """R/gp_loo.R:73-133, quadrature = TRUE branch (any method: the Laplace branch predicts at the
    training points with train = TRUE, R/gp_loo.R:123)."""
    y = np.asarray(y, dtype=np.float64)
    if gp.approx.name == "ep":
        loo_mean, loo_var = gp.fit["cavity_mean"], gp.fit["cavity_var"]
    else:
        fhat, z, V = gp.fit["fmean"], gp.fit["z"], gp.fit["V"]
        grad = (z - fhat) / V
        if gp.method.name == "full":
            post_var = gp_pred_post_full(gp, _as_matrix(x), var=True)["var"]
        elif gp.method.name == "fitc":
            post_var = gp_pred_post_fitc(gp, _as_matrix(x), var=True, train=True)["var"]
        else:
            post_var = gp_pred_post_rf(gp, _as_matrix(x), var=True)["var"]
        loo_var = 1 / (1 / post_var - 1 / V)
        loo_mean = fhat - loo_var * grad
    xq, wq = gauss_hermite_points(quad_order)
    fgrid = xq[None, :] * np.sqrt(loo_var)[:, None] + loo_mean[:, None]
    loglik = get_loglik(gp.lik, fgrid, y, sum=False, trials=trials, offset=offset)
    loos = logsumexp_w(loglik, wq)
    n = len(y)
    return dict(loo=float(np.sum(loos)), sd=float(n * (np.std(loos, ddof=1) / math.sqrt(n))), loos=loos)


# --------------------------------------------------------------------------------------
# gradient oracle (NOT in the reference -- derived; see module docstring)
# --------------------------------------------------------------------------------------

def _dcf_dlogparam(cf, nm, x1, x2):
    """d k / d log(param) for one non-product cf, as a dense matrix."""
    a = prepare_inputmat(cf, x1)
    b = prepare_inputmat(cf, x2)
    name = cf.name
    if nm == "magn":
        return 2.0 * eval_cf(cf, x1, x2)
    if name in ("cf_sexp", "cf_matern32", "cf_matern52") and nm == "lscale":
        r2 = np.zeros((a.shape[0], b.shape[0]))
        for k in range(a.shape[1]):
            r2 += (a[:, k][:, None] - b[:, k][None, :]) ** 2
        l, m2 = cf.lscale, cf.magn ** 2
        if name == "cf_sexp":
            return m2 * np.exp(-0.5 * r2 / l ** 2) * r2 / l ** 2
        r = np.sqrt(r2)
        if name == "cf_matern32":
            return m2 * (3 * r2 / l ** 2) * np.exp(-math.sqrt(3) * r / l)
        return m2 * (5 * r2 / (3 * l ** 2)) * (1 + math.sqrt(5) * r / l) * np.exp(-math.sqrt(5) * r / l)
    if name == "cf_nn":
        s0, s, m2 = cf.sigma0, cf.sigma, cf.magn ** 2
        s0sq = s0 * s0 * s * s
        xa = a @ b.T
        na = np.sum(a * a, axis=1)[:, None]
        nb = np.sum(b * b, axis=1)[None, :]
        A = s0sq + s * s * xa
        B = s0sq + s * s * na
        Cc = s0sq + s * s * nb
        den = np.sqrt((1 + 2 * B) * (1 + 2 * Cc))
        u = 2 * A / den
        dasin = 1.0 / np.sqrt(1 - u * u)
        if nm == "sigma0":      # d/dlog s0 of s0sq = 2 s0sq ; A,B,C all shift by it
            dA = dB = dC = 2 * s0sq
        else:                   # sigma: everything scales with sigma^2
            dA, dB, dC = 2 * A, 2 * B, 2 * Cc
        du = 2 * dA / den - u * (dB / (1 + 2 * B) + dC / (1 + 2 * Cc))
        return m2 * (2 / math.pi) * dasin * du
    raise ValueError((name, nm))


def _member_deriv(m, owner, nm, x1, x2):
    """d k_m / d log(param) for one product member m, or None when m does not own the parameter."""
    if m.name == "cf_periodic":
        if m is owner and nm == "period":
            # central difference of the transform only (oracle-side shortcut; the whole gradient is
            # cross-checked against central differences of the energy in tests/test_oracle.py)
            h = 1e-6
            mp = _clone_with(m, period=m.period * math.exp(h))
            mm = _clone_with(m, period=m.period * math.exp(-h))
            return (eval_cf(mp, x1, x2) - eval_cf(mm, x1, x2)) / (2 * h)
        if m.base is owner:
            a = prepare_inputmat(m, x1)
            b = prepare_inputmat(m, x2)
            w = 2 * math.pi / m.period
            at = np.concatenate([np.sin(w * a), np.cos(w * a)], axis=1)
            bt = np.concatenate([np.sin(w * b), np.cos(w * b)], axis=1)
            return _dcf_dlogparam(m.base, nm, at, bt)
        return None
    if m is owner:
        return _dcf_dlogparam(m, nm, x1, x2)
    return None


def _dK_dlogparam(cfs, owner, nm, x1, x2):
    """d (sum_k prod_f k_f) / d log(param of `owner`)."""
    total = 0.0
    for cf in cfs:
        members = cf.cfs if cf.name == "cf_prod" else [cf]
        for i, m in enumerate(members):
            hit = _member_deriv(m, owner, nm, x1, x2)
            if hit is None:
                continue
            for j, other in enumerate(members):
                if j != i:
                    hit = hit * eval_cf(other, x1, x2)
            total = total + hit
    return total


def _clone_with(obj, **kw):
    import copy
    c = copy.copy(obj)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def energy_grad_gaussian_full(gp, x, y, jitter=None, include_prior=True):
    """Energy and its gradient w.r.t. the log-parameters in get_param.gp order, for
    lik_gaussian + method_full.  E = -log_evidence - log_prior  (R/gp.R:138-147);
    dE/dphi_j = -0.5 tr((alpha alpha^T - C^-1) dC/dphi_j) - dlogprior/dphi_j with
    C = K + (jitter + sigma^2) I."""
    x = _as_matrix(x)
    gp_fit(gp, x, y, jitter=jitter)
    fit = gp.fit
    n = len(y)
    Cc = fit["C_chol"]
    Cinv = sla.cho_solve((Cc, True), np.eye(n), check_finite=False)
    alpha = fit["alpha"]
    W = np.outer(alpha, alpha) - Cinv
    params = free_params(list(gp.cfs) + [gp.lik])
    g = np.zeros(len(params))
    for j, (owner, nm) in enumerate(params):
        if owner is gp.lik:
            dC = 2 * gp.lik.sigma ** 2 * np.eye(n)
        else:
            dC = _dK_dlogparam(gp.cfs, owner, nm, x, x)
        g[j] = -0.5 * np.sum(W * dC)
        pr = owner.priors[nm]
        if include_prior:
            h = 1e-6
            lp = math.log(getattr(owner, nm))
            g[j] -= (lpdf_prior_one(pr, lp + h) - lpdf_prior_one(pr, lp - h)) / (2 * h)
    return gp_energy(gp, include_prior), g


def energy_grad_laplace_full(gp, x, y, trials=None, offset=None, jitter=None, include_prior=True, h=1e-4):
    """Energy and gradient w.r.t. the covariance-function log-parameters for approx_laplace with a
    non-Gaussian likelihood, method_full (NOT in the reference; Rasmussen & Williams Alg. 5.1 applied
    to the reference's objective).  With a = alpha, R = (K + V)^-1, pv = diag((K^-1 + V^-1)^-1),
    d3 = d/df of get_loglik_d2 (central difference of the reference's own finite-difference
    curvature): s2 = pv d3 / 2, u = (I - R K) s2,
    dE/dphi_j = -1/2 sum [(a+u)(a+u)^T - u u^T - R] o dK/dphi_j  - dlogprior/dphi_j."""
    x = _as_matrix(x)
    gp_fit(gp, x, y, trials=trials, offset=offset, jitter=jitter)
    fit = gp.fit
    n = len(y)
    jit = 1e-6 if jitter is None else jitter
    K = eval_cf(gp.cfs, x, x) + jit * np.eye(n)
    f, a, V = fit["fmean"], fit["alpha"], fit["V"]
    s = 1 / np.sqrt(V)
    L = _chol_lower(np.eye(n) + s[:, None] * K * s[None, :])
    Binv = sla.cho_solve((L, True), np.eye(n), check_finite=False)
    R = s[:, None] * Binv * s[None, :]
    pv = V * (1 - np.diag(Binv))
    _, Vp = get_pseudodata_la(gp.lik, f + h, y, trials=trials, offset=offset)
    _, Vm = get_pseudodata_la(gp.lik, f - h, y, trials=trials, offset=offset)
    d3 = (1 / Vm - 1 / Vp) / (2 * h)
    s2 = 0.5 * pv * d3
    u = s2 - R @ (K @ s2)
    p = a + u
    W = np.outer(p, p) - np.outer(u, u) - R
    params = free_params(list(gp.cfs))
    g = np.zeros(len(params))
    for j, (owner, nm) in enumerate(params):
        g[j] = -0.5 * np.sum(W * _dK_dlogparam(gp.cfs, owner, nm, x, x))
        if include_prior:
            hh = 1e-6
            lp = math.log(getattr(owner, nm))
            pr = owner.priors[nm]
            g[j] -= (lpdf_prior_one(pr, lp + hh) - lpdf_prior_one(pr, lp - hh)) / (2 * hh)
    return gp_energy(gp, include_prior), g

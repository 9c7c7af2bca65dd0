"""gp_init / gp_fit / gp_energy / gp_optim / gp_pred: gplite's public API
(NAMESPACE:13-46) over the CUDA library.  All n x n work happens inside libgplite_b200.so;
this file mirrors the R control flow that stays on the host (R/gp.R, R/gp_fit.R:54-88,
R/gp_optim.R:63-160, R/gp_pred.R:88-122).
"""
from __future__ import annotations

import ctypes as C
import math
import warnings

import numpy as np

from . import _lib
from . import model as M
from .quadrature import gauss_hermite_points


class GP:
    def __init__(self, cfs, lik, method, approx):
        self.cfs = cfs
        self.lik = lik
        self.method = method
        self.approx = approx
        self.fitted = False
        self.fit = None
        self.x = None
        self._handle = None      # device-resident gp$fit members (C_chol, Kxz ...)
        self._ctx = None
        self._train = None       # (y, trials, offset, jitter) of the last gp_fit, for gp_save / gp_load

    def __getstate__(self):      # device handles do not travel (gp_save)
        st = dict(self.__dict__)
        st["_handle"] = None
        st["_ctx"] = None
        return st

    def __setstate__(self, st):  # copy.deepcopy / unpickling: the device-resident factor is gone
        self.__dict__.update(st)
        self._handle = None
        self._ctx = None

    def __repr__(self):
        return f"GP(cfs={self.cfs}, lik={self.lik}, {self.method}, {self.approx})"

    def _release(self):
        if self._handle is not None and self._ctx is not None and self._ctx.h:
            self._ctx.lib.gpl_fit_free(self._ctx.h, self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


def gp_init(cfs=None, lik=None, method=None, approx=None):
    """R/gp.R:52-71."""
    cfs = M.cf_sexp() if cfs is None else cfs
    lik = M.lik_gaussian() if lik is None else lik
    method = M.method_full() if method is None else method
    approx = M.approx_laplace() if approx is None else approx
    if lik.name == "lik_gaussian" and approx.name == "ep":
        raise ValueError("approx_ep disabled for lik_gaussian: use approx_laplace instead, because "
                         "it gives exact results")
    if not isinstance(cfs, (list, tuple)):
        cfs = [cfs]
    # R value semantics: the model owns copies, so a cf / lik / method object reused in another model
    # (or twice in this one) is never aliased by set_param / gp_optim / gp_fit
    import copy
    return GP([copy.deepcopy(c) for c in cfs], copy.deepcopy(lik), copy.deepcopy(method), copy.deepcopy(approx))


# ------------------------------------------------------------------------------- parameters
def _param_names(items):
    """R's named vector may repeat a name (two cf_sexp summands); a dict cannot: later repeats get '#2', '#3' ..."""
    seen, out = {}, []
    for name, _, _ in items:
        seen[name] = seen.get(name, 0) + 1
        out.append(name if seen[name] == 1 else f"{name}#{seen[name]}")
    return out


def get_param(gp):
    """Named vector of log-parameters, R/gp.R:150-155."""
    items = M.get_param_objs(gp.cfs, gp.lik)
    return {nm: math.log(getattr(owner, attr)) for nm, (_, owner, attr) in zip(_param_names(items), items)}


def get_param_vector(gp):
    items = M.get_param_objs(gp.cfs, gp.lik)
    return np.array([math.log(getattr(owner, attr)) for _, owner, attr in items], dtype=np.float64)


def set_param(gp, param):
    """R/gp.R:158-165: unpack log-parameters in get_param order; marks the model stale."""
    items = M.get_param_objs(gp.cfs, gp.lik)
    if isinstance(param, dict):
        if list(param.keys()) != _param_names(items):
            raise ValueError("parameter names do not match get_param(gp)")
        values = list(param.values())
    else:
        values = list(np.asarray(param, dtype=np.float64).ravel())
        if len(values) != len(items):
            raise ValueError(f"expected {len(items)} parameters, got {len(values)}")
    for (_, owner, attr), v in zip(items, values):
        setattr(owner, attr, math.exp(float(v)))
    gp.fitted = False
    return gp


def lpdf_prior(gp):
    return sum(M.lpdf_prior_one(owner.priors[attr], math.log(getattr(owner, attr)))
               for _, owner, attr in M.get_param_objs(gp.cfs, gp.lik))


def is_fitted(gp, type="analytic"):
    """R/gp.R:191-205."""
    found = isinstance(gp.fit, dict) and gp.fit.get("type") == type
    if found and not gp.fitted:
        raise RuntimeError(
            "The GP object seems to contain a posterior fit, but is not refitted after setting new "
            "hyperparameter values. Please refit using gp_fit after calling set_param.")
    return found


# ------------------------------------------------------------------------------- packing
def _as_matrix(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    return np.asfortranarray(x)


def _pack_lik(lik):
    param = getattr(lik, "sigma", getattr(lik, "phi", 0.0))
    return _lib.Lik(_lib.LIK_KINDS[lik.name], _lib.LINKS[lik.link], float(param))


def _pack_approx(gp, fitc=False):
    a = gp.approx
    keep = []
    if a.name == "laplace":
        ap = _lib.Approx(0, int(a.maxiter), float(a.tol), 0.0, 0.0, 0, 0, None, None)
    else:
        xq, wq = gauss_hermite_points(int(a.quad_order))
        xq = np.ascontiguousarray(xq)
        wq = np.ascontiguousarray(wq)
        keep = [xq, wq]
        tol = 1e-3 if fitc else 1e-4                         # R/ep.R:286,335
        ap = _lib.Approx(1, int(a.maxiter), tol, float(a.damping), 0.1, int(a.quad_order), 0,
                         xq.ctypes.data_as(C.POINTER(C.c_double)),
                         wq.ctypes.data_as(C.POINTER(C.c_double)))
    return ap, keep


def _vec(v, n):
    if v is None:
        return None
    v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))
    return v


def _mask_of(cfind, nterms):
    """cfind (0-based indices into gp$cfs) -> the ABI's 64-bit term mask (0 means "all terms")."""
    if cfind is None:
        return 0
    idx = [int(k) for k in np.atleast_1d(cfind)]
    if not idx:
        raise ValueError("cfind must select at least one covariance function")
    lim = min(nterms, 64)
    mask = 0
    for k in idx:
        if k < 0 or k >= lim:
            raise ValueError(f"cfind entry {k} out of range (0 .. {lim - 1})")
        mask |= 1 << k
    return mask


def _term_mask(gp, cfind):
    return _mask_of(cfind, len(gp.cfs))


def _need_handle(gp):
    if gp._handle is None or gp._ctx is None:
        raise RuntimeError("the device-resident fit of this GP object is gone (the object was copied or "
                           "unpickled outside gp_load): call gp_fit again")


# ------------------------------------------------------------------------------- eval_cf
def eval_cf(cfs, x1, x2=None, cfind=None, diag_only=False, jitter=0.0, ctx=None):
    """eval_cf.list (R/cf.R:402-411) on the GPU.  x2=None means x2 is x1 (symmetric fast path)."""
    ctx = ctx or _lib.default_context()
    if not isinstance(cfs, (list, tuple)):
        cfs = [cfs]
    x1 = _as_matrix(x1)
    same = x2 is None
    x2m = x1 if same else _as_matrix(x2)
    if x1.shape[1] != x2m.shape[1]:
        raise ValueError("x1 and x2 must have the same number of columns")
    factors, _ = M.flatten_cfs(cfs)
    pm = _lib.PackedModel(factors)
    mask = _mask_of(cfind, len(cfs))
    n1, n2 = x1.shape[0], x2m.shape[0]
    if diag_only:
        out = np.empty(min(n1, n2), dtype=np.float64)
    else:
        out = np.empty((n1, n2), dtype=np.float64, order="F")
    _lib.check(ctx.lib.gpl_cf_eval(ctx.h, pm.ref(), _lib._ptr(x1), n1, _lib._ptr(x2m), n2,
                                   x1.shape[1], 1 if same else 0, mask, 1 if diag_only else 0,
                                   float(jitter), _lib._ptr(out)))
    return out


# ------------------------------------------------------------------------------- gp_fit
def gp_fit(gp, x, y, trials=None, offset=None, jitter=None, ctx=None):
    """R/gp_fit.R:54-60 -> fit_laplace / fit_ep .method_full / .method_fitc."""
    ctx = ctx or _lib.default_context()
    x = _as_matrix(x)
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
    n, d = x.shape
    if len(y) != n:
        raise ValueError("x and y have different numbers of observations")
    for cf in gp.cfs:
        M.learn_scales(cf, x)
    jit = 1e-6 if jitter is None else float(jitter)          # R/gp.R:276-281
    factors, _ = M.flatten_cfs(gp.cfs)
    pm = _lib.PackedModel(factors)
    lik = _pack_lik(gp.lik)
    meth = gp.method.name
    if meth not in ("full", "fitc", "rf"):
        raise NotImplementedError(f"method_{meth} is outside the GPU hot path (SURVEY section 2)")
    ap, keep = _pack_approx(gp, fitc=(meth == "fitc"))
    tr = _vec(trials, n)
    off = _vec(offset, n)
    out = _lib.FitOut()
    vec = {k: np.empty(n) for k in ("fmean", "z", "V", "alpha", "fvar", "cavity_mean", "cavity_var")}
    handle = C.c_void_p()
    gp._release()
    gp._ctx = ctx
    args_tail = [jit, C.byref(out)] + [_lib._ptr(vec[k]) for k in
                                       ("fmean", "z", "V", "alpha", "fvar", "cavity_mean", "cavity_var")] \
        + [C.byref(handle)]
    if meth == "rf":
        if gp.approx.name != "laplace":
            raise NotImplementedError("No EP implementation for method_rf yet.")      # R/gp_fit.R:146-148
        nbasis = M.check_num_basis(gp.cfs, gp.method.num_basis, d)
        gp.method.num_basis = nbasis                                                   # R/gp_fit.R:128
        w = np.ascontiguousarray(M.rf_weights(gp.cfs, nbasis, d, gp.method.seed))
        nb_arr = np.ascontiguousarray(nbasis, dtype=np.int32)
        spec = _lib.RfSpec(len(nbasis), nb_arr.ctypes.data_as(C.POINTER(C.c_int32)),
                           w.ctypes.data_as(C.POINTER(C.c_double)), len(w))
        wmean = np.empty(int(sum(nbasis)))
        _lib.check(ctx.lib.gpl_fit_rf(ctx.h, pm.ref(), C.byref(lik), C.byref(ap), C.byref(spec), _lib._ptr(x),
                                      n, d, _lib._ptr(y), _lib._ptr(tr), _lib._ptr(off), C.byref(out),
                                      _lib._ptr(vec["fmean"]), _lib._ptr(vec["z"]), _lib._ptr(vec["V"]),
                                      _lib._ptr(wmean), C.byref(handle)))
        status = _lib.GPL_OK
    elif meth == "full":
        status = ctx.lib.gpl_fit_full(ctx.h, pm.ref(), C.byref(lik), C.byref(ap), _lib._ptr(x), n, d,
                                      _lib._ptr(y), _lib._ptr(tr), _lib._ptr(off), *args_tail)
        if status == _lib.GPL_ERR_NONFINITE:
            gp.fit = dict(log_evidence=-math.inf)
        else:
            _lib.check(status)                                # non-PD aborts (R/gp_fit.R:95-97)
    else:
        if gp.method.inducing is None:
            gp.method.inducing = get_inducing(gp, x, ctx=ctx)      # kept for later fits (R/gp_fit.R:112)
            gp.method.num_inducing = gp.method.inducing.shape[0]
        zind = _as_matrix(gp.method.inducing)
        m = zind.shape[0]
        status = ctx.lib.gpl_fit_fitc(ctx.h, pm.ref(), C.byref(lik), C.byref(ap), _lib._ptr(x), n, d,
                                      _lib._ptr(zind), m, _lib._ptr(y), _lib._ptr(tr),
                                      _lib._ptr(off), *args_tail)
        if status in (_lib.GPL_ERR_NOT_PD, _lib.GPL_ERR_NONFINITE):
            # R/gp_fit.R:113-121: errors are swallowed into log_evidence = -Inf
            print(_lib.load().gpl_last_error().decode())
            gp.fit = dict(log_evidence=-math.inf)
        else:
            _lib.check(status)
    if status == _lib.GPL_OK:
        fit = dict(fmean=vec["fmean"], z=vec["z"], V=vec["V"], alpha=vec["alpha"],
                   log_evidence=out.log_evidence, iters=out.iters, converged=bool(out.converged),
                   n_bad=int(out.n_bad))
        if out.n_bad > 500:                                   # R/laplace.R:48-50, R/ep.R:59-61
            warnings.warn("More than 500 data points with pseudo-variance < 0 (outliers), computation "
                          "getting slow. Check that the model and hyperparameter values are sensible.")
        if meth == "rf":
            fit.pop("alpha")
            fit["wmean"] = wmean
        if gp.approx.name == "ep":
            fit.update(fvar=vec["fvar"], cavity_mean=vec["cavity_mean"], cavity_var=vec["cavity_var"],
                       P=1.0 / vec["V"], quad_ok=bool(out.quad_ok), accepted=int(out.accepted),
                       damp=float(out.damping_final))
        if gp.approx.name == "laplace" and gp.approx.maxiter > 1 and not out.converged \
                and gp.lik.name != "lik_gaussian":
            warnings.warn("Maximum number of iterations in Laplace reached")
        if gp.approx.name == "ep" and not out.converged:
            warnings.warn("Maximum number of iterations in EP reached")
        gp.fit = fit
        gp._handle = handle
    gp.x = x
    gp.fit["type"] = "analytic"
    gp.fitted = True
    gp._train = (y, None if trials is None else np.asarray(trials, dtype=np.float64),
                 None if offset is None else np.asarray(offset, dtype=np.float64), jitter)
    return gp


def gp_save(gp, filename):
    """R/saveload.R:32-34.  The model, its hyperparameters, the host-side members of gp$fit and the
    training data are pickled; the device-resident factor is not (see gp_load)."""
    import pickle
    with open(filename, "wb") as f:
        pickle.dump(gp, f)


def gp_load(filename, ctx=None):
    """R/saveload.R:38-41.  A model saved after gp_fit is refitted on load from the stored training
    data (deterministic: same hyperparameters, same inducing points / random features), which
    rebuilds the device-resident factor that prediction needs."""
    import pickle
    with open(filename, "rb") as f:
        gp = pickle.load(f)
    if gp.fitted and gp.fit is not None and gp._train is not None and gp.x is not None:
        y, trials, offset, jitter = gp._train
        gp_fit(gp, gp.x, y, trials=trials, offset=offset, jitter=jitter, ctx=ctx)
    return gp


def kmeans_centers(x, m, rng, ctx=None, maxiter=10):
    """stats::kmeans(x, m)$centers as used by get_inducing: Lloyd iterations on the device
    (gpl_kmeans) from m distinct rows drawn with `rng`; x is used as given (already standardised)."""
    ctx = ctx or _lib.default_context()
    x = _as_matrix(x)
    n, d = x.shape
    if m > n:
        raise ValueError("more cluster centers than distinct data points.")      # stats::kmeans' message
    idx = np.ascontiguousarray(np.sort(rng.choice(n, size=m, replace=False)), dtype=np.int64)
    cen = np.empty((m, d), dtype=np.float64, order="F")
    iters = C.c_int32()
    _lib.check(ctx.lib.gpl_kmeans(ctx.h, _lib._ptr(x), n, d, _lib._ptr(idx), m, int(maxiter), _lib._ptr(cen),
                                  None, C.byref(iters)))
    return cen, idx


def _scale(x):
    """R's scale(): centre, divide by the (n - 1) standard deviation."""
    mu = x.mean(axis=0)
    sd = x.std(axis=0, ddof=1)
    return (x - mu) / sd, mu, sd


def get_inducing(gp, x, ctx=None):
    """R/gp_fit.R:190-234: k-means centres of the standardised inputs, optionally per bin along one
    variable.  The reference seeds R's RNG with method$seed; here numpy's default_rng(seed) picks the
    initial centres, so the points agree with the reference in distribution, not draw by draw."""
    if gp.method.inducing is not None:
        return gp.method.inducing
    x = _as_matrix(x)
    rng = np.random.default_rng(gp.method.seed)
    m = int(gp.method.num_inducing)
    if gp.method.bin_along is None:
        xs, mu, sd = _scale(x)
        zs, _ = kmeans_centers(xs, m, rng, ctx)
        return np.asfortranarray(zs * sd + mu)
    var = int(gp.method.bin_along)             # 0-based column index, like `vars` (R: 1-based)
    nb = int(gp.method.bin_count)
    eps2 = 2 * np.finfo(np.float64).eps
    cut = np.linspace(x[:, var].min() - eps2, x[:, var].max() + eps2, nb + 1)
    bins = [x[(x[:, var] > cut[i]) & (x[:, var] < cut[i + 1])] for i in range(nb)]
    bins.sort(key=lambda b: -b.shape[0])                               # order(bin_sizes, decreasing = TRUE)
    per_bin = [m // nb] * nb
    for i in range(m - sum(per_bin)):
        per_bin[i] += 1
    out = []
    for xb, k in zip(bins, per_bin):
        if k == 0:
            continue
        xs, mu, sd = _scale(xb)
        zs, _ = kmeans_centers(xs, k, rng, ctx)
        out.append(zs * sd + mu)
    return np.asfortranarray(np.vstack(out))


def fit_member(gp, what):
    """Materialise a device-resident member of gp$fit on demand ("C_chol", "alpha")."""
    n = gp.x.shape[0]
    m = gp.method.inducing.shape[0] if gp.method.name == "fitc" else 0
    M_rf = int(sum(gp.method.num_basis)) if gp.method.name == "rf" else 0
    shape = {"C_chol": (n, n), "alpha": (n,), "Kxz": (n, m), "Kz_chol": (m, m), "var_train": (n,), "wmean": (M_rf,),
             "wcov": (M_rf, M_rf)}[what]
    out = np.empty(shape, dtype=np.float64, order="F")
    _need_handle(gp)
    _lib.check(gp._ctx.lib.gpl_fit_get(gp._ctx.h, gp._handle, what.encode(), _lib._ptr(out)))
    return out


# ------------------------------------------------------------------------------- energy
def gp_energy(gp, include_prior=True):
    """R/gp.R:138-147."""
    if not is_fitted(gp, "analytic"):
        raise RuntimeError("The GP must be fitted. Call gp_fit first.")
    e = -gp.fit["log_evidence"]
    if include_prior:
        e -= lpdf_prior(gp)
    return e


def gp_energy_grad(gp, x, y, jitter=None, include_prior=True, ctx=None, device_ptrs=None, trials=None,
                   offset=None):
    """Energy and its gradient w.r.t. get_param(gp) -- the new capability (the reference's
    gp_optim is derivative-free).  method_full with lik_gaussian, or with any other likelihood under
    approx_laplace (implicit-function gradient at the Newton mode: use a tight approx_laplace(tol=);
    the likelihood's own parameters must be fixed).

    device_ptrs=(x_ptr, y_ptr, n, d): use data already resident in HBM (bench.py's `value`)."""
    ctx = ctx or _lib.default_context()
    if gp.method.name != "full":
        raise NotImplementedError("the analytic gradient is implemented for method_full")
    if gp.lik.name != "lik_gaussian":
        return _energy_grad_laplace(gp, x, y, jitter, include_prior, ctx, trials, offset)
    if offset is not None and np.any(np.asarray(offset) != 0):
        # gp_fit / gp_energy put the offset into the Gaussian log-likelihood (R/lik.R:301-303) while the
        # pseudo-data ignore it (R/lik.R:188-192); the analytic gradient does not cover that extra term
        raise NotImplementedError("gp_energy_grad: offset is not supported for lik_gaussian "
                                  "(use method='Nelder-Mead' in gp_optim)")
    items = M.get_param_objs(gp.cfs, gp.lik)
    if device_ptrs is None:
        x = _as_matrix(x)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
        n, d = x.shape
        for cf in gp.cfs:
            M.learn_scales(cf, x)
        xp, yp, on_dev = _lib._ptr(x), _lib._ptr(y), 0
    else:
        xp, yp, n, d = device_ptrs
        xp, yp, on_dev = C.c_void_p(xp), C.c_void_p(yp), 1
    factors, where = M.flatten_cfs(gp.cfs)
    pm = _lib.PackedModel(factors)
    lik = _pack_lik(gp.lik)
    pf = np.empty(len(items), dtype=np.int32)
    ps = np.empty(len(items), dtype=np.int32)
    for j, (_, owner, attr) in enumerate(items):
        if owner is gp.lik:
            pf[j], ps[j] = -1, 0
        else:
            pf[j], ps[j] = where[(id(owner), attr)]
    energy = C.c_double()
    grad = np.zeros(len(items), dtype=np.float64)
    jit = 1e-6 if jitter is None else float(jitter)
    _lib.check(ctx.lib.gpl_energy_grad(ctx.h, pm.ref(), C.byref(lik), xp, n, d, yp, on_dev, jit,
                                       len(items), _lib._ptr(pf), _lib._ptr(ps), C.byref(energy),
                                       _lib._ptr(grad), None))
    e = energy.value
    if include_prior:
        for j, (_, owner, attr) in enumerate(items):
            lp = math.log(getattr(owner, attr))
            e -= M.lpdf_prior_one(owner.priors[attr], lp)
            grad[j] -= M.dlpdf_prior_one(owner.priors[attr], lp)
    return e, grad


def _energy_grad_laplace(gp, x, y, jitter, include_prior, ctx, trials, offset):
    if gp.approx.name != "laplace":
        raise NotImplementedError("the analytic gradient needs approx_laplace for non-Gaussian likelihoods")
    items = M.get_param_objs(gp.cfs, gp.lik)
    if any(owner is gp.lik for _, owner, _ in items):
        raise NotImplementedError("fix the likelihood's parameters (prior_fixed) to use the analytic gradient")
    x = _as_matrix(x)
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
    n, d = x.shape
    for cf in gp.cfs:
        M.learn_scales(cf, x)
    factors, where = M.flatten_cfs(gp.cfs)
    pm = _lib.PackedModel(factors)
    lik = _pack_lik(gp.lik)
    ap, keep = _pack_approx(gp)
    pf = np.empty(len(items), dtype=np.int32)
    ps = np.empty(len(items), dtype=np.int32)
    for j, (_, owner, attr) in enumerate(items):
        pf[j], ps[j] = where[(id(owner), attr)]
    energy = C.c_double()
    grad = np.zeros(max(len(items), 1), dtype=np.float64)
    out = _lib.FitOut()
    jit = 1e-6 if jitter is None else float(jitter)
    tr, off = _vec(trials, n), _vec(offset, n)
    _lib.check(ctx.lib.gpl_energy_grad_laplace(ctx.h, pm.ref(), C.byref(lik), C.byref(ap), _lib._ptr(x), n, d,
                                               _lib._ptr(y), _lib._ptr(tr), _lib._ptr(off), jit, len(items),
                                               _lib._ptr(pf), _lib._ptr(ps), C.byref(energy), _lib._ptr(grad),
                                               C.byref(out)))
    e, grad = energy.value, grad[:len(items)]
    if include_prior:
        for j, (_, owner, attr) in enumerate(items):
            lp = math.log(getattr(owner, attr))
            e -= M.lpdf_prior_one(owner.priors[attr], lp)
            grad[j] -= M.dlpdf_prior_one(owner.priors[attr], lp)
    return e, grad


# ------------------------------------------------------------------------------- gp_optim
def gp_optim(gp, x, y, tol=1e-4, tol_param=1e-1, maxiter=500, restarts=1, verbose=True,
             warn=True, method="Nelder-Mead", **fit_args):
    """R/gp_optim.R:63-125.  method="Nelder-Mead" reproduces the reference's derivative-free
    search (scipy's simplex in place of stats::optim, fnscale = n); method="L-BFGS-B" uses the
    analytic gradient (method_full: lik_gaussian, or any likelihood with approx_laplace)."""
    from scipy.optimize import minimize
    names = list(get_param(gp).keys())
    n = len(np.asarray(y).ravel())
    it = [0]

    def energy(p):
        set_param(gp, p)
        gp_fit(gp, x, y, **fit_args)
        e = gp_energy(gp)
        if verbose:
            print(" ".join(f"{v:8.2f}" for v in p), f"{e:10.2f}", f"{it[0]:9d}")
        it[0] += 1
        return e

    if verbose:
        print("Optimizing parameters")
        print("\n".join(f"p{i + 1}: log {nm}" for i, nm in enumerate(names)))
    p0 = get_param_vector(gp)
    if method == "Nelder-Mead":
        res = minimize(lambda p: energy(p) / n, p0, method="Nelder-Mead",
                       options=dict(fatol=tol, xatol=np.inf, maxiter=maxiter))
        e_opt = res.fun * n
    else:
        def eg(p):
            set_param(gp, p)
            e, g = gp_energy_grad(gp, x, y, **fit_args)
            if verbose:
                print(" ".join(f"{v:8.2f}" for v in p), f"{e:10.2f}", f"{it[0]:9d}")
            it[0] += 1
            return e / n, g / n
        res = minimize(eg, p0, jac=True, method="L-BFGS-B", options=dict(maxiter=maxiter, ftol=tol))
        e_opt = res.fun * n
    if not res.success and warn and res.get("nit", 0) >= maxiter:
        warnings.warn("Maximum number of iterations reached, optimization may not have converged.")
    set_param(gp, res.x)
    gp_fit(gp, x, y, **fit_args)
    if warn:
        converged, msg = _check_convergence(energy, e_opt, res.x, delta=tol_param)
        if not converged:
            if restarts > 0:
                if verbose:
                    print("Parameters not yet converged, restarting...\n")
                set_param(gp, res.x)          # R/gp_optim.R:104-113 restarts from param_opt
                return gp_optim(gp, x, y, tol=tol, tol_param=tol_param, maxiter=maxiter,
                                restarts=restarts - 1, verbose=verbose, warn=warn, method=method,
                                **fit_args)
            warnings.warn(msg)
        elif verbose:
            print(msg)
        set_param(gp, res.x)
        gp_fit(gp, x, y, **fit_args)
    return gp


def _check_convergence(lossfun, loss_opt, param_opt, delta=1e-1, tol=1e-2):
    """R/gp_optim.R:127-160."""
    for k in range(len(param_opt)):
        dp = np.zeros(len(param_opt))
        dp[k] = delta
        l1 = lossfun(param_opt + dp)
        l2 = lossfun(param_opt - dp)
        if loss_opt - min(l1, l2) > tol:
            return False, (f"Not all hyperparameters have reached convergence within tolerance "
                           f"tol_param = {delta:f}. Try increasing the number of restarts, or "
                           f"reducing tol.")
    return True, "Optimization converged."


# ------------------------------------------------------------------------------- gp_pred
def _norm_ppf(q, mean, sd):
    from scipy.special import ndtri
    return mean + sd * ndtri(q)


def _get_response(lik, f):
    """R/lik.R:458-498 (host side: only used on nt-sized vectors)."""
    if lik.name == "lik_gaussian":
        return f
    if lik.name == "lik_poisson":
        return np.exp(f)
    if lik.link == "probit":
        from scipy.special import ndtr
        return ndtr(f)
    return 1.0 / (1.0 + np.exp(-f))


def gp_pred(gp, xnew, var=False, quantiles=None, transform=False, cfind=None, jitter=None,
            quad_order=15, offset=None, ctx=None):
    """R/gp_pred.R:88-122."""
    if quantiles is not None or transform:
        var = True
    xnew = _as_matrix(xnew)
    nt = xnew.shape[0]
    if not is_fitted(gp, "analytic"):
        # prior prediction, R/gp_pred.R:142-156
        pred = dict(mean=np.zeros(nt))
        if var:
            pred["var"] = eval_cf(gp.cfs, xnew, xnew, cfind=cfind, diag_only=True, ctx=ctx)
    else:
        if "alpha" not in gp.fit and "wmean" not in gp.fit:
            raise RuntimeError("the fit failed (log_evidence = -Inf); cannot predict")
        _need_handle(gp)
        c = gp._ctx
        factors, _ = M.flatten_cfs(gp.cfs)
        pm = _lib.PackedModel(factors)
        mean = np.empty(nt)
        v = np.empty(nt) if var else None
        fn = {"full": c.lib.gpl_pred_full, "fitc": c.lib.gpl_pred_fitc, "rf": c.lib.gpl_pred_rf}[gp.method.name]
        if xnew.shape[1] != gp.x.shape[1]:
            raise ValueError("xnew has a different number of columns than the training inputs")
        _lib.check(fn(c.h, gp._handle, pm.ref(), _lib._ptr(xnew), nt, _term_mask(gp, cfind),
                      _lib._ptr(mean), _lib._ptr(v)))
        pred = dict(mean=mean)
        if var:
            pred["var"] = v
    if offset is not None:
        pred["mean"] = pred["mean"] + offset
    if quantiles is not None:
        q = np.stack([_norm_ppf(qq, pred["mean"], np.sqrt(pred["var"])) for qq in quantiles], axis=1)
        pred["quantiles"] = _get_response(gp.lik, q) if transform else q
    if transform:
        xq, wq = gauss_hermite_points(quad_order)
        fgrid = xq[None, :] * np.sqrt(pred["var"])[:, None] + pred["mean"][:, None]
        ft = _get_response(gp.lik, fgrid)
        mt = ft @ wq
        pred["var"] = ((ft - mt[:, None]) ** 2) @ wq
        pred["mean"] = mt
    return pred


# ------------------------------------------------------------------------------- gp_draw
def gp_pred_cov(gp, xnew, cfind=None, jitter=None, ctx=None):
    """gp_pred_post / gp_pred_prior with cov = TRUE (R/gp_pred.R:142-156, 202-232, 262-273):
    dict(mean, cov) of the latent, jitter on the diagonal of cov."""
    xnew = _as_matrix(xnew)
    nt = xnew.shape[0]
    jit = 1e-6 if jitter is None else float(jitter)
    if not is_fitted(gp, "analytic"):
        K = eval_cf(gp.cfs, xnew, xnew, cfind=cfind, ctx=ctx)
        return dict(mean=np.zeros(nt), cov=K + jit * np.eye(nt))
    if gp.method.name not in ("full", "fitc"):
        raise NotImplementedError("joint predictive covariance is implemented for method_full and method_fitc")
    _need_handle(gp)
    c = gp._ctx
    factors, _ = M.flatten_cfs(gp.cfs)
    pm = _lib.PackedModel(factors)
    mean = np.empty(nt)
    cov = np.empty((nt, nt), order="F")
    _lib.check(c.lib.gpl_pred_full_cov(c.h, gp._handle, pm.ref(), _lib._ptr(xnew), nt, _term_mask(gp, cfind),
                                       jit, _lib._ptr(mean), _lib._ptr(cov)))
    return dict(mean=mean, cov=cov)


def generate_target(lik, f, rng, trials=None):
    """R/lik.R:502-537: sample observations given latent values (one column per draw)."""
    mu = _get_response(lik, f)
    if lik.name == "lik_gaussian":
        return rng.standard_normal(f.shape) * lik.sigma + f
    if lik.name == "lik_poisson":
        return rng.poisson(mu).astype(float)
    t = np.ones(f.shape[0]) if lik.name == "lik_bernoulli" else np.asarray(trials, dtype=np.float64)
    t = np.broadcast_to(t.reshape(-1, *([1] * (f.ndim - 1))), f.shape).astype(np.int64)
    if lik.name == "lik_betabinom":
        mu = rng.beta(mu / lik.phi, (1 - mu) / lik.phi)
    return rng.binomial(t, mu).astype(float)


def gp_draw(gp, xnew, draws=1, transform=True, target=False, marginal=False, cfind=None, jitter=None,
            seed=None, offset=None, trials=None, normals=None, ctx=None):
    """R/gp_draw.R:8-123: draws from the posterior (fitted model) or prior of the latent function at
    xnew, nt x draws.  The joint draw mean + t(chol(cov)) z runs on the device (gpl_draw_full);
    z comes from numpy's default_rng(seed) -- or `normals` -- instead of R's rnorm stream."""
    xnew = _as_matrix(xnew)
    nt = xnew.shape[0]
    draws = int(draws or 1)
    rng = np.random.default_rng(seed)
    jit = 1e-6 if jitter is None else float(jitter)
    fitted = is_fitted(gp, "analytic")
    rf = fitted and gp.method.name == "rf"         # draws live in weight space: normals is M_selected x draws
    if not rf:
        if normals is None:
            normals = rng.standard_normal((nt, draws))
        normals = np.asfortranarray(np.asarray(normals, dtype=np.float64).reshape(nt, draws))
    if fitted and "alpha" not in gp.fit and "wmean" not in gp.fit:
        raise RuntimeError("the fit failed (log_evidence = -Inf); cannot draw")
    if fitted:
        _need_handle(gp)
    if rf:                                                       # R/gp_draw.R:145-163
        # w ~ N(wmean, wcov) restricted to the selected summands, sample = featuremap(xt, cfind) w; the
        # M x M factor and the M-vector draws are host-sized, the features come from the device
        c = gp._ctx
        factors, _ = M.flatten_cfs(gp.cfs)
        pm = _lib.PackedModel(factors)
        nb = list(gp.method.num_basis)
        ends = np.concatenate([[0], np.cumsum(nb)])
        idx = range(len(nb)) if cfind is None else list(cfind)
        inds = np.concatenate([np.arange(ends[k], ends[k + 1]) for k in idx])
        Mtot = int(ends[-1])
        phi = np.empty((nt, Mtot), order="F")
        _lib.check(c.lib.gpl_rf_features(c.h, gp._handle, pm.ref(), _lib._ptr(xnew), nt, _term_mask(gp, cfind),
                                         _lib._ptr(phi)))
        wcov = fit_member(gp, "wcov")[np.ix_(inds, inds)]
        zw = rng.standard_normal((len(inds), draws)) if normals is None \
            else np.asarray(normals, dtype=np.float64).reshape(len(inds), draws)
        w = gp.fit["wmean"][inds][:, None] + np.linalg.cholesky(wcov) @ zw
        sample = phi[:, inds] @ w
    elif fitted and marginal:                                    # R/gp_draw.R:107-110,127-130
        pred = gp_pred(gp, xnew, var=True, cfind=cfind, jitter=jitter, ctx=ctx)
        sample = pred["mean"][:, None] + np.sqrt(pred["var"])[:, None] * normals
    else:
        if gp.method.name not in ("full", "fitc") or (gp.method.name == "fitc" and not fitted):
            raise NotImplementedError("joint draws are implemented for method_full (prior and posterior) "
                                      "and fitted method_fitc models; use marginal=True otherwise")
        c = gp._ctx if fitted else (ctx or _lib.default_context())
        factors, _ = M.flatten_cfs(gp.cfs)
        pm = _lib.PackedModel(factors)
        if fitted and xnew.shape[1] != gp.x.shape[1]:
            raise ValueError("xnew has a different number of columns than the training inputs")
        sample = np.empty((nt, draws), order="F")
        _lib.check(c.lib.gpl_draw_full(c.h, gp._handle if fitted else None, pm.ref(), _lib._ptr(xnew), nt,
                                       xnew.shape[1], _term_mask(gp, cfind), jit, _lib._ptr(normals), draws,
                                       _lib._ptr(sample)))
    if offset is not None:
        sample = sample + np.asarray(offset, dtype=np.float64).reshape(-1, 1)
    if target:
        sample = generate_target(gp.lik, sample, rng, trials=trials)
    elif transform:
        sample = _get_response(gp.lik, sample)
    return sample


# ------------------------------------------------------------------------------- gp_loo
def _get_loglik_grid(lik, fgrid, y, trials=None, offset=None):
    """get_loglik(..., sum = FALSE) on an n x q grid (R/lik.R:297-366).  Host side: n*q values."""
    from scipy.special import gammaln, xlogy, xlog1py
    f = fgrid if offset is None else fgrid + np.asarray(offset, dtype=np.float64)[:, None]
    yb = np.asarray(y, dtype=np.float64)[:, None]
    if lik.name == "lik_gaussian":
        s = lik.sigma
        return -0.5 * math.log(2 * math.pi) - math.log(s) - 0.5 * ((yb - f) / s) ** 2
    if lik.name == "lik_poisson":
        mu = np.exp(f)
        return xlogy(yb, mu) - mu - gammaln(yb + 1)
    t = np.ones_like(yb) if lik.name == "lik_bernoulli" else np.asarray(trials, dtype=np.float64)[:, None]
    mu = _get_response(lik, f)
    if lik.name == "lik_betabinom":
        a, b = mu / lik.phi, (1 - mu) / lik.phi
        return ((gammaln(t + 1) - gammaln(yb + 1) - gammaln(t - yb + 1))
                + (gammaln(yb + a) + gammaln(t - yb + b) - gammaln(t + a + b))
                + (gammaln(a + b) - gammaln(a) - gammaln(b)))
    lch = gammaln(t + 1) - gammaln(yb + 1) - gammaln(t - yb + 1)
    return lch + xlogy(yb, mu) + xlog1py(t - yb, -mu)


def gp_loo(gp, x, y, quad_order=11, trials=None, offset=None):
    """Leave-one-out predictive densities, R/gp_loo.R:73-133 (quadrature = TRUE branch), for
    every method.  The expensive part -- the posterior variances at the training points for the
    Laplace approximation (gp_pred_post with train = TRUE, R/gp_loo.R:123) -- runs on the GPU: through
    gp_pred for method_full / method_rf, and as the fit's own marginal variances for method_fitc; the
    n x q quadrature is host side."""
    if not is_fitted(gp, "analytic"):
        raise RuntimeError("The provied GP model is not fitted. Please call gp_fit (or gp_optim) first.")
    y = np.asarray(y, dtype=np.float64).ravel()
    n = len(y)
    if gp.approx.name == "ep":                                     # R/gp_loo.R:131-133
        loo_mean, loo_var = gp.fit["cavity_mean"], gp.fit["cavity_var"]
    else:                                                          # R/gp_loo.R:119-129
        fhat, z, V = gp.fit["fmean"], gp.fit["z"], gp.fit["V"]
        grad = (z - fhat) / V
        post_var = fit_member(gp, "var_train") if gp.method.name == "fitc" else gp_pred(gp, x, var=True)["var"]
        loo_var = 1.0 / (1.0 / post_var - 1.0 / V)
        loo_mean = fhat - loo_var * grad
    xq, wq = gauss_hermite_points(quad_order)
    fgrid = xq[None, :] * np.sqrt(loo_var)[:, None] + loo_mean[:, None]
    loglik = _get_loglik_grid(gp.lik, fgrid, y, trials=trials, offset=offset)
    c = loglik.max(axis=1, keepdims=True)
    loos = (c + np.log(np.sum(wq[None, :] * np.exp(loglik - c), axis=1, keepdims=True)))[:, 0]
    return dict(loo=float(loos.sum()), sd=float(n * (loos.std(ddof=1) / math.sqrt(n))), loos=loos)


def gp_compare(*loos, ref=None):
    """R/gp_loo.R:147-175."""
    vals = [l["loo"] for l in loos]
    if ref is None:
        ref = int(np.argmax(vals))
    n = len(loos[0]["loos"])
    out = np.zeros((len(loos), 2))
    for i, l in enumerate(loos):
        dlt = l["loos"] - loos[ref]["loos"]
        out[i, 0] = n * dlt.mean()
        out[i, 1] = n * (dlt.std(ddof=1) / math.sqrt(n))
    return out

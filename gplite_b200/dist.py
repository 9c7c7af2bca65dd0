"""Multi-GPU full GP (BASELINE config 5: n = 200k, method_full): thin host wrappers over the
`gpl_dist_*` entries of the C ABI.  One process per GPU; the NCCL communicator, the block-cyclic
layout, the panel broadcasts and the Newton / EP / gradient loops all live inside
libgplite_b200.so (csrc/dist.cu, csrc/dist_sched.hpp) -- this module only carries the 128-byte NCCL
id from rank 0 to the other ranks and marshals arguments, exactly what the R glue would do.

    comm = DistComm.from_env()                    # under torchrun: RANK / WORLD_SIZE / LOCAL_RANK
    res = dist_fit_laplace(gp, x, y, comm=comm)   # every rank passes the same x, y and gets the same result

Reference semantics: fit_laplace / fit_ep .method_full (R/gp_fit.R:91-99,150-161), laplace.method_full
(R/laplace.R:13-38,274-297), ep.method_full (R/ep.R:16-49,286-330); the energy gradient has no
counterpart in the reference (its gp_optim is derivative-free, R/gp_optim.R:94).

The row-sharded FITC fit (dist_fit_fitc) is the same kind of wrapper over gpl_dist_fit_fitc.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _lib
from . import model as M


class DistComm:
    """One rank's handle on the library's communicator (gpl_dist).  `exchange(id_or_None) -> id` must
    return rank 0's 128-byte id on every rank; from_env / from_torch provide it through a
    torch.distributed process group (any backend), a file-based rendezvous works just as well."""

    def __init__(self, rank, world, exchange, ctx=None, device=None, have_id=False):
        self.lib = _lib.load()
        self.rank, self.world = int(rank), int(world)
        if ctx is None:
            ctx = _lib.Context(int(os.environ.get("LOCAL_RANK", rank)) if device is None else device)
        self.ctx = ctx
        buf = C.create_string_buffer(128)
        if self.world > 1:
            if self.rank == 0 and not have_id:
                _lib.check(self.lib.gpl_dist_unique_id(buf))
            raw = exchange(bytes(buf.raw) if self.rank == 0 else None)
            buf = C.create_string_buffer(bytes(raw), 128)
        h = C.c_void_p()
        _lib.check(self.lib.gpl_dist_init(ctx.h, buf, self.rank, self.world, C.byref(h)))
        self.h = h

    @staticmethod
    def exchange_id(dist):
        """rank 0 draws the NCCL id inside the library; the process group carries its 128 bytes"""
        raw = None
        if dist.get_rank() == 0:
            buf = C.create_string_buffer(128)
            _lib.check(_lib.load().gpl_dist_unique_id(buf))
            raw = bytes(buf.raw)
        box = [raw]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    @classmethod
    def from_torch(cls, dist, ctx=None, device=None):
        """`dist`: an initialised torch.distributed module (gloo or nccl)."""
        raw = cls.exchange_id(dist) if dist.get_world_size() > 1 else None
        return cls(dist.get_rank(), dist.get_world_size(), lambda _: raw, ctx=ctx, device=device, have_id=True)

    @classmethod
    def single(cls, ctx=None):
        return cls(0, 1, None, ctx=ctx or _lib.default_context())

    def release(self):
        """free the cached n x n / world slabs"""
        _lib.check(self.lib.gpl_dist_release(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpl_dist_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _timings_dict(tm):
    return {k: getattr(tm, k) for k, _ in _lib.DistTimings._fields_}


def _prep(gp, x, y):
    x = np.asfortranarray(np.asarray(x, dtype=np.float64))
    if x.ndim == 1:
        x = np.asfortranarray(x[:, None])
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
    if len(y) != x.shape[0]:
        raise ValueError("x and y have different numbers of observations")
    for cf in gp.cfs:
        M.learn_scales(cf, x)
    factors, where = M.flatten_cfs(gp.cfs)
    return x, y, _lib.PackedModel(factors), where


def _dist_fit(gp, x, y, trials, offset, jitter, nb, comm, timings):
    from .gp import _pack_lik, _pack_approx, _vec
    if gp.method.name != "full":
        raise NotImplementedError("the sharded factorisation covers method_full (FITC: dist_fit_fitc)")
    x, y, pm, _ = _prep(gp, x, y)
    n, d = x.shape
    lik = _pack_lik(gp.lik)
    ap, keep = _pack_approx(gp)
    tr, off = _vec(trials, n), _vec(offset, n)
    out = _lib.FitOut()
    tm = _lib.DistTimings()
    vec = {k: np.empty(n) for k in ("fmean", "z", "V", "alpha", "fvar", "cavity_mean", "cavity_var")}
    jit = 1e-6 if jitter is None else float(jitter)
    status = comm.lib.gpl_dist_fit_full(comm.h, pm.ref(), C.byref(lik), C.byref(ap), _lib._ptr(x), n, d,
                                        _lib._ptr(y), _lib._ptr(tr), _lib._ptr(off), jit, int(nb), C.byref(out),
                                        *[_lib._ptr(vec[k]) for k in ("fmean", "z", "V", "alpha", "fvar",
                                                                      "cavity_mean", "cavity_var")],
                                        C.byref(tm))
    if timings is not None:
        timings.update(_timings_dict(tm))
    if status == _lib.GPL_ERR_NONFINITE:                       # R/ep.R:30-32
        return dict(log_evidence=-math.inf, iters=int(out.iters), converged=False, quad_ok=bool(out.quad_ok))
    _lib.check(status)
    res = dict(log_evidence=out.log_evidence, alpha=vec["alpha"], fmean=vec["fmean"], z=vec["z"], V=vec["V"],
               iters=int(out.iters), converged=bool(out.converged))
    if gp.approx.name == "ep":
        res.update(fvar=vec["fvar"], cavity_mean=vec["cavity_mean"], cavity_var=vec["cavity_var"],
                   P=1.0 / vec["V"], quad_ok=bool(out.quad_ok), accepted=int(out.accepted),
                   damp=float(out.damping_final))
    return res


def dist_fit_laplace(gp, x, y, trials=None, offset=None, jitter=None, nb=1024, comm=None, timings=None):
    """Distributed gp_fit for method_full + approx_laplace, any likelihood with log-concave sites
    (gpl_dist_fit_full).  Returns dict(log_evidence, alpha, fmean, z, V, iters, converged), numpy
    arrays, identical on every rank."""
    if gp.approx.name != "laplace":
        raise NotImplementedError("dist_fit_laplace covers approx_laplace (EP: dist_fit_ep)")
    return _dist_fit(gp, x, y, trials, offset, jitter, nb, comm, timings)


def dist_fit_gaussian(gp, x, y, jitter=None, nb=1024, comm=None, timings=None):
    if gp.lik.name != "lik_gaussian":
        raise NotImplementedError("dist_fit_gaussian is for lik_gaussian; use dist_fit_laplace")
    return dist_fit_laplace(gp, x, y, jitter=jitter, nb=nb, comm=comm, timings=timings)


def dist_fit_ep(gp, x, y, trials=None, offset=None, jitter=None, nb=1024, comm=None, timings=None):
    """Distributed gp_fit for method_full + approx_ep: parallel EP (R/ep.R:16-49, 286-330) over the
    sharded factor; marginal variances from diag(B^-1) (one more panel sweep per EP sweep)."""
    if gp.approx.name != "ep":
        raise NotImplementedError("dist_fit_ep covers approx_ep")
    return _dist_fit(gp, x, y, trials, offset, jitter, nb, comm, timings)


def dist_energy_grad(gp, x, y, jitter=None, include_prior=True, nb=1024, comm=None, timings=None, grad=True,
                     want_alpha=False):
    """Multi-GPU gp_energy_grad (lik_gaussian + method_full, gpl_dist_energy_grad): energy and its
    gradient w.r.t. get_param(gp), same conventions as gp.gp_energy_grad; every rank returns the same
    values.  grad=False: energy only (one n x n / world slab per GPU instead of two)."""
    from .gp import _pack_lik
    if gp.method.name != "full" or gp.lik.name != "lik_gaussian":
        raise NotImplementedError("dist_energy_grad covers method_full + lik_gaussian")
    x, y, pm, where = _prep(gp, x, y)
    n, d = x.shape
    items = M.get_param_objs(gp.cfs, gp.lik)
    pf = np.empty(len(items), dtype=np.int32)
    ps = np.empty(len(items), dtype=np.int32)
    for j, (_, owner, attr) in enumerate(items):
        pf[j], ps[j] = (-1, 0) if owner is gp.lik else where[(id(owner), attr)]
    lik = _pack_lik(gp.lik)
    energy = C.c_double()
    g = np.zeros(max(len(items), 1))
    alpha = np.empty(n) if want_alpha else None
    tm = _lib.DistTimings()
    jit = 1e-6 if jitter is None else float(jitter)
    _lib.check(comm.lib.gpl_dist_energy_grad(comm.h, pm.ref(), C.byref(lik), _lib._ptr(x), n, d, _lib._ptr(y), jit,
                                             int(nb), len(items) if grad else 0, _lib._ptr(pf), _lib._ptr(ps),
                                             C.byref(energy), _lib._ptr(g), _lib._ptr(alpha), C.byref(tm)))
    if timings is not None:
        timings.update(_timings_dict(tm))
    e = energy.value
    g = g[:len(items)]
    if include_prior:
        for j, (_, owner, attr) in enumerate(items):
            lp = math.log(getattr(owner, attr))
            e -= M.lpdf_prior_one(owner.priors[attr], lp)
            g[j] -= M.dlpdf_prior_one(owner.priors[attr], lp)
    out = (e, g if grad else None)
    return out + (alpha,) if want_alpha else out


# ================================================================================================
# FITC across GPUs ("shards naturally, one reduction")
# ================================================================================================

def dist_fit_fitc(gp, x, y, trials=None, offset=None, jitter=None, comm=None):
    """Row-sharded gp_fit for method_fitc, approx_laplace or approx_ep (gpl_dist_fit_fitc): rank r works on
    rows [n r / P, n (r + 1) / P) of x; Kz and the m x m factorisations are replicated; per Newton step /
    EP sweep the data plane is one all-reduce of the m x m matrix Kzx S^-1 Kxz plus a few m-vectors and
    scalars, inside the library's own NCCL communicator.  Every rank passes the same x, y and gets the
    same n-vectors back; `gp` is left fitted (rank-local device handle), so gp_pred works on any rank."""
    from .gp import _pack_lik, _pack_approx, _vec, get_inducing
    if gp.method.name != "fitc":
        raise NotImplementedError("dist_fit_fitc covers method_fitc")
    x, y, pm, _ = _prep(gp, x, y)
    n, d = x.shape
    if gp.method.inducing is None:
        gp.method.inducing = get_inducing(gp, x, ctx=comm.ctx)          # same seed -> same points on every rank
        gp.method.num_inducing = gp.method.inducing.shape[0]
    zind = np.asfortranarray(np.asarray(gp.method.inducing, dtype=np.float64))
    if zind.ndim == 1:
        zind = np.asfortranarray(zind[:, None])
    m = zind.shape[0]
    lik = _pack_lik(gp.lik)
    ap, keep = _pack_approx(gp, fitc=True)
    tr, off = _vec(trials, n), _vec(offset, n)
    out = _lib.FitOut()
    vec = {k: np.empty(n) for k in ("fmean", "z", "V", "alpha", "fvar", "cavity_mean", "cavity_var")}
    handle = C.c_void_p()
    jit = 1e-6 if jitter is None else float(jitter)
    gp._release()
    status = comm.lib.gpl_dist_fit_fitc(comm.h, pm.ref(), C.byref(lik), C.byref(ap), _lib._ptr(x), n, d,
                                        _lib._ptr(zind), m, _lib._ptr(y), _lib._ptr(tr), _lib._ptr(off), jit,
                                        C.byref(out), *[_lib._ptr(vec[k]) for k in
                                                        ("fmean", "z", "V", "alpha", "fvar", "cavity_mean",
                                                         "cavity_var")], C.byref(handle))
    if status in (_lib.GPL_ERR_NOT_PD, _lib.GPL_ERR_NONFINITE):       # R/gp_fit.R:113-121: swallowed into -Inf
        gp.fit = dict(log_evidence=-math.inf, type="analytic")
        gp.fitted, gp.x = True, x
        return gp.fit
    _lib.check(status)
    fit = dict(fmean=vec["fmean"], z=vec["z"], V=vec["V"], alpha=vec["alpha"], log_evidence=out.log_evidence,
               iters=int(out.iters), converged=bool(out.converged), n_bad=0, type="analytic")
    if gp.approx.name == "ep":
        fit.update(fvar=vec["fvar"], cavity_mean=vec["cavity_mean"], cavity_var=vec["cavity_var"],
                   P=1.0 / vec["V"], quad_ok=bool(out.quad_ok), accepted=int(out.accepted),
                   damp=float(out.damping_final))
    gp.fit, gp._handle, gp._ctx, gp.x, gp.fitted = fit, handle, comm.ctx, x, True
    return fit


def dist_fit_fitc_gaussian(gp, x, y, jitter=None, comm=None):
    if gp.lik.name != "lik_gaussian":
        raise NotImplementedError("dist_fit_fitc_gaussian is for lik_gaussian; use dist_fit_fitc")
    return dist_fit_fitc(gp, x, y, jitter=jitter, comm=comm)

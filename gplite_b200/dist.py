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

The row-sharded FITC driver further down still moves its m x m all-reduce through torch.distributed.
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

class FitcOps:
    """C-ABI calls of the row-sharded FITC driver on torch CUDA tensors.  Matrices are column-major
    slabs: a (rows x cols) matrix is a tensor of shape (cols, ld), element (i, j) at t[j, i]."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib

    @staticmethod
    def _p(t, offset_elems=0):
        return C.c_void_p(t.data_ptr() + 8 * offset_elems)

    def ksym(self, pm, Z_t, m, d, jitter, out_t):
        """out (m x m) = K(z, z) + jitter I"""
        _lib.check(self.lib.gpl_cf_eval_dev(self.ctx.h, pm.ref(), self._p(Z_t), m, self._p(Z_t), m, d, 1, 0, 0,
                                            float(jitter), self._p(out_t), out_t.shape[1]))

    def kblock(self, pm, X_t, nr, ldx, Z_t, m, d, out_t):
        """out (nr x m) = K(x_local, z)"""
        _lib.check(self.lib.gpl_cf_eval_block_dev(self.ctx.h, pm.ref(), self._p(X_t), nr, ldx, self._p(Z_t), m, m,
                                                  d, self._p(out_t), out_t.shape[1]))

    def kdiag(self, pm, X_t, nr, d, jitter, out_t):
        _lib.check(self.lib.gpl_cf_eval_dev(self.ctx.h, pm.ref(), self._p(X_t), nr, self._p(X_t), nr, d, 1, 0, 1,
                                            float(jitter), self._p(out_t), nr))

    def potrf(self, A_t, m, dinv_t):
        info = C.c_int32()
        logdet = C.c_double()
        st = self.lib.gpl_dpotrf_dev(self.ctx.h, m, self._p(A_t), A_t.shape[1], C.byref(info), C.byref(logdet))
        if st == _lib.GPL_OK:
            _lib.check(self.lib.gpl_dpotrf_get_dinv_dev(self.ctx.h, m, self._p(dinv_t)))
        return st, info.value, logdet.value

    def rownorms_after_solve(self, W_t, rows, m, L_t, dinv_t, out_t, out_off):
        """out[off:off+rows] = row sums of squares of W L^-T  (W is destroyed)"""
        _lib.check(self.lib.gpl_dtrsm_right_dinv_dev(self.ctx.h, 1, rows, m, self._p(L_t), L_t.shape[1],
                                                     self._p(dinv_t), self._p(W_t), W_t.shape[1]))
        _lib.check(self.lib.gpl_rowsumsq_dev(self.ctx.h, rows, m, self._p(W_t), W_t.shape[1], self._p(out_t, out_off)))

    def scale_rows(self, W_t, rows, m, s_t, s_off):
        _lib.check(self.lib.gpl_scale_block_dev(self.ctx.h, rows, m, self._p(W_t), W_t.shape[1], self._p(s_t, s_off),
                                                None, 1.0, 0.0, -(1 << 40)))

    def syrk_acc(self, W_t, rows, m, A_t):
        """A (lower) += W^T W"""
        _lib.check(self.lib.gpl_dgemm_dev(self.ctx.h, 1, 0, m, m, rows, 1.0, self._p(W_t), W_t.shape[1],
                                          self._p(W_t), W_t.shape[1], 1.0, self._p(A_t), A_t.shape[1], 1))

    def gemv(self, trans, rows, m, K_t, x_t, y_t):
        _lib.check(self.lib.gpl_dgemv_dev(self.ctx.h, trans, rows, m, self._p(K_t), K_t.shape[1], self._p(x_t),
                                          self._p(y_t)))

    def pseudodata(self, lik, f_t, y_t, tr_t, off_t, z_t, V_t):
        """z, V <- get_pseudodata_la(lik, f, y) on device n-vectors (R/lik.R:188-201)"""
        _lib.check(self.lib.gpl_lik_pseudodata_la_dev(
            self.ctx.h, C.byref(lik), f_t.numel(), self._p(f_t), self._p(y_t),
            self._p(tr_t) if tr_t is not None else None, self._p(off_t) if off_t is not None else None,
            self._p(z_t), self._p(V_t)))

    def loglik(self, lik, f_t, y_t, tr_t, off_t):
        out = C.c_double()
        _lib.check(self.lib.gpl_lik_loglik_dev(
            self.ctx.h, C.byref(lik), f_t.numel(), self._p(f_t), self._p(y_t),
            self._p(tr_t) if tr_t is not None else None, self._p(off_t) if off_t is not None else None,
            C.byref(out)))
        return out.value

    def trsv(self, trans, m, L_t, dinv_t, b_t, x_t):
        _lib.check(self.lib.gpl_dtrsv_dev(self.ctx.h, trans, m, self._p(L_t), L_t.shape[1], self._p(dinv_t),
                                          self._p(b_t), self._p(x_t)))

    def sync(self):
        self.ctx.sync()


def dist_fit_fitc_gaussian(gp, x, y, jitter=None, ops=None, dist=None, device="cuda", chunk_rows=None):
    """lik_gaussian entry point kept for callers of the first version: see dist_fit_fitc."""
    if gp.lik.name != "lik_gaussian":
        raise NotImplementedError("dist_fit_fitc_gaussian is for lik_gaussian; use dist_fit_fitc")
    return dist_fit_fitc(gp, x, y, jitter=jitter, ops=ops, dist=dist, device=device, chunk_rows=chunk_rows)


def dist_fit_fitc(gp, x, y, trials=None, offset=None, jitter=None, ops=None, dist=None, device="cuda",
                  chunk_rows=None):
    """Row-sharded FITC fit with approx_laplace for any likelihood with log-concave sites
    (R/gp_fit.R:101-123 + R/laplace.R:122-152,299-325 + R/linear_algebra.R:3-43): rank r owns rows
    [r n/P, (r+1) n/P) of x and the site vectors of those rows; Kz and the m x m factorisations are
    replicated; per Newton step the only data-path collectives are one all-reduce of the m x m matrix
    Kzx S^-1 Kxz and a few m-vectors / scalars.  Every rank passes the full x, y."""
    import torch
    from .gp import _pack_lik
    if gp.method.name != "fitc" or gp.approx.name != "laplace":
        raise NotImplementedError("dist_fit_fitc covers method_fitc + approx_laplace")
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    y = np.asarray(y, dtype=np.float64).ravel()
    zind = np.asarray(gp.method.inducing, dtype=np.float64)
    if zind.ndim == 1:
        zind = zind[:, None]
    n, d = x.shape
    m = zind.shape[0]
    for cf in gp.cfs:
        M.learn_scales(cf, x)                       # training means / scales use ALL rows
    jit = 1e-6 if jitter is None else float(jitter)
    gaussian = gp.lik.name == "lik_gaussian"
    lik = _pack_lik(gp.lik)
    factors, _ = M.flatten_cfs(gp.cfs)
    pm = _lib.PackedModel(factors)
    if ops is None:
        ops = FitcOps(_lib.Context(torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream or 1))
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    nr = hi - lo
    f64 = dict(dtype=torch.float64, device=device)
    ldr = (max(nr, 1) + 15) // 16 * 16
    ldm = (m + 15) // 16 * 16
    loc = lambda v: None if v is None else torch.from_numpy(
        np.asarray(v, dtype=np.float64).ravel()[lo:hi].copy()).to(device)
    X_t = torch.from_numpy(np.ascontiguousarray(x[lo:hi].T)).to(device)          # (d, nr)
    Z_t = torch.from_numpy(np.ascontiguousarray(zind.T)).to(device)              # (d, m)
    y_t, tr_t, off_t = loc(y), loc(trials), loc(offset)
    if gp.lik.name in ("lik_binomial", "lik_betabinom") and tr_t is None:
        raise ValueError("trials must be provided for the binomial / beta-binomial likelihood")
    nblk = (m + 127) // 128
    Kz = torch.zeros((m, ldm), **f64)
    Lz = torch.zeros((m, ldm), **f64)
    A = torch.zeros((m, ldm), **f64)
    dinv_z = torch.empty(nblk * 128 * 128, **f64)
    dinv_a = torch.empty(nblk * 128 * 128, **f64)
    Kxz = torch.zeros((m, ldr), **f64)
    ops.ksym(pm, Z_t, m, d, jit, Kz)
    Lz.copy_(Kz)
    st, info, logdet_z = ops.potrf(Lz, m, dinv_z)
    if st != _lib.GPL_OK:
        raise _lib.NotPositiveDefinite(st, "Kz is not positive definite")
    ops.kblock(pm, X_t, nr, nr, Z_t, m, d, Kxz)
    Kdiag = torch.empty(max(nr, 1), **f64)
    ops.kdiag(pm, X_t, nr, d, jit, Kdiag)
    # D = K_diag - |Lz^-1 k_zi|^2, streamed through a row-chunk workspace
    if chunk_rows is None:
        chunk_rows = max(1024, int(2.0e9 / (8.0 * m)) // 128 * 128)
    chunk_rows = min(chunk_rows, max(nr, 1))
    ldw = (chunk_rows + 15) // 16 * 16
    W = torch.empty((m, ldw), **f64)
    q = torch.empty(max(nr, 1), **f64)
    for r0 in range(0, nr, chunk_rows):
        c = min(chunk_rows, nr - r0)
        W[:, :c].copy_(Kxz[:, r0:r0 + c])
        ops.rownorms_after_solve(W, c, m, Lz, dinv_z, q, r0)
    D = Kdiag[:nr] - q[:nr]
    f_loc = torch.zeros(nr, **f64)                                     # R/laplace.R:305
    z_loc = torch.empty(nr, **f64)
    V_loc = torch.empty(nr, **f64)
    w1 = torch.empty(m, **f64)
    maxiter = 1 if gaussian else max(int(gp.approx.maxiter), 1)        # R/laplace.R:306-308
    tol = float(gp.approx.tol)
    log_post_old = -math.inf
    it = 0
    for it in range(1, maxiter + 1):
        # laplace_iter.method_fitc, R/laplace.R:122-152, on the local rows
        if nr > 0:
            ops.pseudodata(lik, f_loc, y_t, tr_t, off_t, z_loc, V_loc)
        vmin = torch.tensor([float(V_loc.min().item()) if nr > 0 else math.inf], **f64)
        dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
        if not vmin.item() > 0:
            raise _lib.GplError(_lib.GPL_ERR_NOT_LOGCONCAVE,
                                "pseudo-variance <= 0: the non-log-concave branch (R/laplace.R:153-232) "
                                "is single-GPU only")
        S = D + V_loc
        rs = torch.rsqrt(S).contiguous()
        # A = Kz + sum over ranks of Kzx S^-1 Kxz  (lower triangle)
        part = torch.zeros((m, ldm), **f64)
        for r0 in range(0, nr, chunk_rows):
            c = min(chunk_rows, nr - r0)
            W[:, :c].copy_(Kxz[:, r0:r0 + c])
            ops.scale_rows(W, c, m, rs, r0)
            ops.syrk_acc(W, c, m, part)
        dist.all_reduce(part)
        A.copy_(Kz + part)
        st, info, logdet_a = ops.potrf(A, m, dinv_a)
        if st != _lib.GPL_OK:
            raise _lib.NotPositiveDefinite(st, "Kz + Kzx S^-1 Kxz is not positive definite")
        # alpha = z/S - S^-1 Kxz A^-1 Kzx (z/S)
        t = (z_loc / S).contiguous()
        u = torch.zeros(m, **f64)
        if nr > 0:
            ops.gemv(1, nr, m, Kxz, t, u)
        dist.all_reduce(u)
        v = torch.empty(m, **f64)
        ops.trsv(0, m, A, dinv_a, u.clone(), w1)
        ops.trsv(1, m, A, dinv_a, w1.clone(), v)
        kv = torch.zeros(max(nr, 1), **f64)
        if nr > 0:
            ops.gemv(0, nr, m, Kxz, v, kv)
        alpha = t - kv[:nr] / S
        # f = Kxz Kz^-1 Kzx alpha + D alpha
        u2 = torch.zeros(m, **f64)
        if nr > 0:
            ops.gemv(1, nr, m, Kxz, alpha.contiguous(), u2)
        dist.all_reduce(u2)
        beta = torch.empty(m, **f64)
        ops.trsv(0, m, Lz, dinv_z, u2.clone(), w1)
        ops.trsv(1, m, Lz, dinv_z, w1.clone(), beta)
        kb = torch.zeros(max(nr, 1), **f64)
        if nr > 0:
            ops.gemv(0, nr, m, Kxz, beta, kb)
        fmean = (kb[:nr] + D * alpha).contiguous()
        ll = ops.loglik(lik, fmean, y_t, tr_t, off_t) if nr > 0 else 0.0
        sc = torch.stack([torch.log(S).sum(), torch.dot(fmean, alpha), torch.log(V_loc).sum(),
                          torch.tensor(ll, **f64)])
        dist.all_reduce(sc)
        sum_log_S, f_invK_f, V_logdet, log_lik = [float(vv) for vv in sc]
        C_logdet = sum_log_S - logdet_z + logdet_a                                # R/linear_algebra.R:32
        log_evidence = log_lik - 0.5 * f_invK_f + 0.5 * (V_logdet - C_logdet)     # R/laplace.R:143-150
        log_post = log_lik - 0.5 * f_invK_f
        f_loc = fmean
        diff = log_post - log_post_old
        log_post_old = log_post
        if abs(diff) < tol:                                                       # R/laplace.R:314-318
            break
    return dict(log_evidence=log_evidence, alpha_local=alpha, fmean_local=f_loc, z_local=z_loc, V_local=V_loc,
                rows=(lo, hi), beta=beta, iters=it, converged=not (maxiter > 1 and it == maxiter))

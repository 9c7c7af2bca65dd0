"""CPU, world_size 2 over gloo: the multi-GPU driver's host logic (block-cyclic ownership, panel
broadcast schedule, distributed triangular solves, evidence assembly) with the C-ABI calls replaced
by a torch-CPU stand-in that lives only here.  The result must equal the oracle's gp_fit."""
import copy
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class TorchOps:
    """Stand-in for gplite_b200.dist.GpuOps on CPU tensors (test-only)."""

    def __init__(self, gp):
        self.gp = gp

    def kblock(self, pm, X_t, n, d, r0, nrows, c0, ncols, A_t, lc0):
        import torch
        from oracle import gplite_oracle as o
        x = X_t.numpy().T
        K = o.eval_cf(self.gp.cfs, x[r0:r0 + nrows], x[c0:c0 + ncols])
        A_t[lc0:lc0 + ncols, r0:r0 + nrows] = torch.from_numpy(np.ascontiguousarray(K.T))

    def scale_block(self, A_t, lc0, ncols, r0, nrows, alpha, diag_add, diag_offset, s_t=None, c0=0):
        A_t[lc0:lc0 + ncols, r0:r0 + nrows] *= alpha
        if s_t is not None:
            A_t[lc0:lc0 + ncols, r0:r0 + nrows] *= s_t[c0:c0 + ncols, None] * s_t[None, r0:r0 + nrows]
        for j in range(ncols):
            i = j + diag_offset
            if 0 <= i < nrows:
                A_t[lc0 + j, r0 + i] += diag_add

    def potrf(self, A_t, lc0, r0, nbk):
        import torch
        blk = A_t[lc0:lc0 + nbk, r0:r0 + nbk].T
        sym = torch.tril(blk) + torch.tril(blk, -1).T
        L = torch.linalg.cholesky(sym)
        A_t[lc0:lc0 + nbk, r0:r0 + nbk] = L.T
        return 0, 0, float(2 * torch.log(torch.diagonal(L)).sum())

    def save_dinv(self, nbk, out_t):
        pass

    def trsm_below(self, A_t, lc0, r0, nbk, m):
        import torch
        L = torch.tril(A_t[lc0:lc0 + nbk, r0:r0 + nbk].T)
        B = A_t[lc0:lc0 + nbk, r0 + nbk:r0 + nbk + m].T          # m x nbk
        X = torch.linalg.solve_triangular(L, B.T, upper=False).T
        A_t[lc0:lc0 + nbk, r0 + nbk:r0 + nbk + m] = X.T

    def update(self, P_t, ldp, nbk, off_rows, m, off_cols, ncols, A_t, lc0, r0):
        Pm = P_t.T                                                # (panel rows) x nbk
        A_t[lc0:lc0 + ncols, r0:r0 + m] -= (Pm[off_rows:off_rows + m] @ Pm[off_cols:off_cols + ncols].T).T

    def trsv(self, trans, nbk, A_t, lc0, r0, dinv_t, b_t, x_t):
        import torch
        L = torch.tril(A_t[lc0:lc0 + nbk, r0:r0 + nbk].T)
        M = L.T if trans else L
        x_t[:nbk] = torch.linalg.solve_triangular(M, b_t[:nbk, None], upper=bool(trans))[:, 0]

    def gemv(self, trans, m, ncols, A_t, lc0, r0, x_t, x_off, y_t, y_off):
        A = A_t[lc0:lc0 + ncols, r0:r0 + m].T                     # m x ncols
        if trans:
            y_t[y_off:y_off + ncols] = A.T @ x_t[x_off:x_off + m]
        else:
            y_t[y_off:y_off + m] = A @ x_t[x_off:x_off + ncols]

    @staticmethod
    def _view(spec, rows, cols):
        import torch
        t, off, ld = spec
        base = t if t.dim() == 1 and t.is_contiguous() else t.reshape(-1) if t.is_contiguous() else None
        assert base is not None
        return torch.as_strided(base, (rows, cols), (1, ld), base.storage_offset() + off)

    def gemm(self, ta, tb, m, n, k, alpha, A, B, beta, Cm, lower=0):
        import torch
        a = self._view(A, k, m).T if ta else self._view(A, m, k)
        b = self._view(B, n, k).T if tb else self._view(B, k, n)
        c = self._view(Cm, m, n)
        r = alpha * (a @ b) + (beta * c if beta != 0 else 0.0)
        if lower:
            mask = torch.tril(torch.ones(m, n, dtype=torch.bool))
            c[mask] = r[mask]
        else:
            c.copy_(r)

    def inv_diag(self, A_t, lc0, k0, w, dinv_t, out_t, ldo):
        import torch
        L = torch.tril(A_t[lc0:lc0 + w, k0:k0 + w].T)
        Li = torch.linalg.solve_triangular(L, torch.eye(w, dtype=torch.float64), upper=False)
        torch.as_strided(out_t, (w, w), (1, ldo)).copy_(Li)

    def grad_trace_block(self, pm, X_t, n, d, c0, w, Y_t, lc0, s_t, alpha_t, pf, ps):
        from oracle import gplite_oracle as o
        import gplite_b200.model as M
        x = X_t.numpy().T
        Y = Y_t[lc0:lc0 + w, :n].numpy().T                       # n x w block column of Binv (lower valid)
        s, a = s_t.numpy(), alpha_t.numpy()
        rows = np.arange(n)[:, None]
        cols = (c0 + np.arange(w))[None, :]
        wgt = np.where(rows > cols, 2.0, np.where(rows == cols, 1.0, 0.0))
        Wm = wgt * (np.outer(a, a[c0:c0 + w]) - s[:, None] * Y * s[None, c0:c0 + w])
        Wm[wgt == 0] = 0.0
        items = M.get_param_objs(self.gp.cfs, self.gp.lik)
        out = np.zeros(len(pf))
        for p, (_, owner, attr) in enumerate(items):
            if pf[p] < 0:
                continue
            dK = o._dK_dlogparam(self.gp.cfs, owner, attr, x, x[c0:c0 + w])
            out[p] = np.sum(Wm * dK)
        dsum = float(np.sum(s[c0:c0 + w] ** 2 * Y[c0 + np.arange(w), np.arange(w)]))
        return out, dsum

    def colsumsq(self, A_t, lc0, r0, rows, ncols, out_t, out_off):
        blk = A_t[lc0:lc0 + ncols, r0:r0 + rows]
        out_t[out_off:out_off + ncols] = (blk * blk).sum(dim=1)

    def kdiag(self, pm, X_t, n, d, jitter, out_t):
        import torch
        from oracle import gplite_oracle as o
        x = X_t.numpy().T
        out_t.copy_(torch.from_numpy(o.eval_cf(self.gp.cfs, x, x, diag_only=True) + jitter))

    def ep_sites(self, lik, fmean, fvar, z_old, P_old, y_t, tr_t, off_t, damping, xq_t, wq_t,
                 z_t, V_t, P_t, cm_t, cv_t, logZ_t):
        import torch
        from oracle import gplite_oracle as o
        pobs = o.get_pseudodata_ep(self.gp.lik, fmean.numpy(), 1 / fvar.numpy(), z_old.numpy(), P_old.numpy(),
                                   y_t.numpy(), quad_order=xq_t.numel(), damping=damping,
                                   **self._lik_args(tr_t, off_t))
        for t, v in ((z_t, pobs["z"]), (V_t, pobs["var"]), (P_t, 1 / pobs["var"]), (cm_t, pobs["mean_cavity"]),
                     (cv_t, pobs["var_cavity"]), (logZ_t, pobs["log_normalizing_const"])):
            t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float64)))
        fl = 0 if pobs["quad_ok"] else 1
        if np.any(np.isnan(pobs["var"]) | np.isnan(pobs["z"])):
            fl |= 2
        elif not np.all(pobs["var"] > 0):
            fl |= 4
        return fl

    def _lik_args(self, tr_t, off_t):
        return dict(trials=None if tr_t is None else tr_t.numpy(), offset=None if off_t is None else off_t.numpy())

    def pseudodata(self, lik, f_t, y_t, tr_t, off_t, z_t, V_t):
        import torch
        from oracle import gplite_oracle as o
        z, V = o.get_pseudodata_la(self.gp.lik, f_t.numpy(), y_t.numpy(), **self._lik_args(tr_t, off_t))
        z_t.copy_(torch.from_numpy(z))
        V_t.copy_(torch.from_numpy(V))

    def loglik(self, lik, f_t, y_t, tr_t, off_t):
        from oracle import gplite_oracle as o
        return float(o.get_loglik(self.gp.lik, f_t.numpy(), y_t.numpy(), **self._lik_args(tr_t, off_t)))

    def sync(self):
        pass


def _bernoulli_case():
    import gplite_b200 as g
    rng = np.random.default_rng(11)
    n = 500
    X = rng.normal(size=(n, 2))
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(2 * X[:, 0] - X[:, 1])))).astype(float)
    gp = g.gp_init(cfs=g.cf_sexp(lscale=0.8), lik=g.lik_bernoulli(), approx=g.approx_laplace())
    return gp, X, y


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import gplite_b200 as g
    from gplite_b200.dist import dist_fit_gaussian, BlockCyclic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = BlockCyclic(1000, 128, world, rank)
        assert lay.nblk == 8 and lay.owned == list(range(rank, 8, world))
        assert sum(lay.width) == 1000 and lay.width[-1] == 104
        rng = np.random.default_rng(3)
        n = 600
        X = rng.normal(size=(n, 3))
        y = np.sin(X[:, 0]) + 0.4 * X[:, 1] + 0.3 * rng.normal(size=n)
        gp = g.gp_init(cfs=[g.cf_matern52(lscale=1.1), g.cf_lin(magn=0.3)], lik=g.lik_gaussian(sigma=0.4))
        res = dist_fit_gaussian(gp, X, y, nb=128, ops=TorchOps(gp), dist=dist, device="cpu")
        # Newton loop over the sharded factor (non-Gaussian likelihood)
        from gplite_b200.dist import dist_fit_laplace
        gpb, Xb, yb = _bernoulli_case()
        resb = dist_fit_laplace(gpb, Xb, yb, nb=128, ops=TorchOps(gpb), dist=dist, device="cpu")
        # energy gradient: distributed inverse (two more panel sweeps) + block trace
        from gplite_b200.dist import dist_energy_grad
        gpg = g.gp_init(cfs=[g.cf_matern52(lscale=1.1), g.cf_lin(magn=0.3)], lik=g.lik_gaussian(sigma=0.4))
        e, gr = dist_energy_grad(gpg, X, y, nb=128, ops=TorchOps(gpg), dist=dist, device="cpu")
        # parallel EP over the sharded factor
        from gplite_b200.dist import dist_fit_ep
        gpe, Xe, ye = _bernoulli_case()
        gpe.approx = g.approx_ep()
        rese = dist_fit_ep(gpe, Xe[:300], ye[:300], nb=128, ops=TorchOps(gpe), dist=dist, device="cpu")
        q.put((rank, res["log_evidence"], res["alpha"].numpy().copy(), resb["log_evidence"],
               resb["fmean"].numpy().copy(), resb["iters"], e, gr, rese["log_evidence"],
               rese["fvar"].numpy().copy(), rese["iters"]))
    finally:
        dist.destroy_process_group()


def test_distributed_driver_world2_matches_oracle():
    import torch.multiprocessing as mp
    import gplite_b200 as g
    from oracle import gplite_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    n = 600
    X = rng.normal(size=(n, 3))
    y = np.sin(X[:, 0]) + 0.4 * X[:, 1] + 0.3 * rng.normal(size=n)
    gpo = g.gp_init(cfs=[g.cf_matern52(lscale=1.1), g.cf_lin(magn=0.3)], lik=g.lik_gaussian(sigma=0.4))
    o.gp_fit(gpo, X, y)
    gpb, Xb, yb = _bernoulli_case()
    o.gp_fit(gpb, Xb, yb)
    gpg = g.gp_init(cfs=[g.cf_matern52(lscale=1.1), g.cf_lin(magn=0.3)], lik=g.lik_gaussian(sigma=0.4))
    eo, go = o.energy_grad_gaussian_full(gpg, X, y)
    gpe, Xe, ye = _bernoulli_case()
    gpe.approx = g.approx_ep()
    o.gp_fit(gpe, Xe[:300], ye[:300])
    for rank, le, alpha, leb, fb, itb, e, gr, lee, fve, ite in out:
        assert ite == gpe.fit["iters"]
        assert abs(lee - gpe.fit["log_evidence"]) < 1e-8 * abs(gpe.fit["log_evidence"])
        assert np.max(np.abs(fve - gpe.fit["fvar"])) < 1e-8 * np.max(np.abs(gpe.fit["fvar"]))
        assert abs(e - eo) < 1e-9 * abs(eo)
        assert np.max(np.abs(gr - go)) < 1e-6 * max(1.0, np.max(np.abs(go))), (gr, go)
        assert abs(le - gpo.fit["log_evidence"]) < 1e-9 * abs(gpo.fit["log_evidence"])
        assert np.max(np.abs(alpha - gpo.fit["alpha"])) < 1e-7 * np.max(np.abs(gpo.fit["alpha"]))
        assert itb == gpb.fit["iters"] and itb > 2
        assert abs(leb - gpb.fit["log_evidence"]) < 1e-9 * abs(gpb.fit["log_evidence"])
        assert np.max(np.abs(fb - gpb.fit["fmean"])) < 1e-8 * np.max(np.abs(gpb.fit["fmean"]))


# ------------------------------------------------------------------------------ FITC row sharding
class TorchFitcOps:
    """Stand-in for gplite_b200.dist.FitcOps on CPU tensors (test-only)."""

    def __init__(self, gp):
        self.gp = gp

    def _k(self, a, b, **kw):
        from oracle import gplite_oracle as o
        return o.eval_cf(self.gp.cfs, a, b, **kw)

    def ksym(self, pm, Z_t, m, d, jitter, out_t):
        import torch
        z = Z_t.numpy().T
        out_t[:, :m] = torch.from_numpy(self._k(z, z) + jitter * np.eye(m))

    def kblock(self, pm, X_t, nr, ldx, Z_t, m, d, out_t):
        import torch
        K = self._k(X_t.numpy().T, Z_t.numpy().T)          # nr x m
        out_t[:, :nr] = torch.from_numpy(np.ascontiguousarray(K.T))

    def kdiag(self, pm, X_t, nr, d, jitter, out_t):
        import torch
        x = X_t.numpy().T
        out_t[:nr] = torch.from_numpy(self._k(x, x, diag_only=True) + jitter)

    def potrf(self, A_t, m, dinv_t):
        import torch
        blk = A_t[:, :m].T
        L = torch.linalg.cholesky(torch.tril(blk) + torch.tril(blk, -1).T)
        A_t[:, :m] = L.T
        return 0, 0, float(2 * torch.log(torch.diagonal(L)).sum())

    def rownorms_after_solve(self, W_t, rows, m, L_t, dinv_t, out_t, out_off):
        import torch
        L = torch.tril(L_t[:, :m].T)
        Wm = W_t[:, :rows].T                                  # rows x m
        X = torch.linalg.solve_triangular(L, Wm.T, upper=False)   # m x rows
        out_t[out_off:out_off + rows] = (X * X).sum(dim=0)

    def scale_rows(self, W_t, rows, m, s_t, s_off):
        W_t[:, :rows] *= s_t[s_off:s_off + rows][None, :]

    def syrk_acc(self, W_t, rows, m, A_t):
        Wm = W_t[:, :rows].T
        A_t[:, :m] += (Wm.T @ Wm).T

    def gemv(self, trans, rows, m, K_t, x_t, y_t):
        Km = K_t[:, :rows].T                                  # rows x m
        if trans:
            y_t[:m] = Km.T @ x_t[:rows]
        else:
            y_t[:rows] = Km @ x_t[:m]

    def trsv(self, trans, m, L_t, dinv_t, b_t, x_t):
        import torch
        L = torch.tril(L_t[:, :m].T)
        Mx = L.T if trans else L
        x_t[:m] = torch.linalg.solve_triangular(Mx, b_t[:m, None], upper=bool(trans))[:, 0]

    def sync(self):
        pass


TorchFitcOps._lik_args = TorchOps._lik_args
TorchFitcOps.pseudodata = TorchOps.pseudodata
TorchFitcOps.loglik = TorchOps.loglik


def _fitc_problem():
    rng = np.random.default_rng(8)
    n, m = 501, 40
    X = rng.uniform(-3, 3, size=(n, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.4 * rng.normal(size=n)
    Zi = X[rng.permutation(n)[:m]].copy()
    return X, y, Zi


def _fitc_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import gplite_b200 as g
    from gplite_b200.dist import dist_fit_fitc_gaussian
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, y, Zi = _fitc_problem()
        gp = g.gp_init(cfs=[g.cf_sexp(lscale=1.1), g.cf_const(magn=0.2)], lik=g.lik_gaussian(sigma=0.4),
                       method=g.method_fitc(inducing=Zi))
        res = dist_fit_fitc_gaussian(gp, X, y, ops=TorchFitcOps(gp), dist=dist, device="cpu", chunk_rows=100)
        # Newton loop over the row shards (Poisson counts)
        from gplite_b200.dist import dist_fit_fitc
        yc = np.random.default_rng(5).poisson(np.exp(0.4 * y)).astype(float)
        gpc = g.gp_init(cfs=g.cf_sexp(lscale=1.1), lik=g.lik_poisson(), method=g.method_fitc(inducing=Zi))
        resc = dist_fit_fitc(gpc, X, yc, ops=TorchFitcOps(gpc), dist=dist, device="cpu", chunk_rows=100)
        q.put((rank, res["log_evidence"], res["rows"], res["alpha_local"].numpy().copy(),
               res["fmean_local"].numpy().copy(), resc["log_evidence"], resc["fmean_local"].numpy().copy(),
               resc["iters"]))
    finally:
        dist.destroy_process_group()


def test_distributed_fitc_world2_matches_oracle():
    import torch.multiprocessing as mp
    import gplite_b200 as g
    from oracle import gplite_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fitc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, y, Zi = _fitc_problem()
    gpo = g.gp_init(cfs=[g.cf_sexp(lscale=1.1), g.cf_const(magn=0.2)], lik=g.lik_gaussian(sigma=0.4),
                    method=g.method_fitc(inducing=Zi))
    o.gp_fit(gpo, X, y)
    yc = np.random.default_rng(5).poisson(np.exp(0.4 * y)).astype(float)
    gpc = g.gp_init(cfs=g.cf_sexp(lscale=1.1), lik=g.lik_poisson(), method=g.method_fitc(inducing=Zi))
    o.gp_fit(gpc, X, yc)
    for rank, le, (lo, hi), alpha, fmean, lec, fc, itc in out:
        assert itc == gpc.fit["iters"] and itc > 2
        assert abs(lec - gpc.fit["log_evidence"]) < 1e-9 * abs(gpc.fit["log_evidence"])
        assert np.max(np.abs(fc - gpc.fit["fmean"][lo:hi])) < 1e-8 * np.max(np.abs(gpc.fit["fmean"]))
        assert abs(le - gpo.fit["log_evidence"]) < 1e-9 * abs(gpo.fit["log_evidence"])
        assert np.max(np.abs(alpha - gpo.fit["alpha"][lo:hi])) < 1e-7 * np.max(np.abs(gpo.fit["alpha"]))
        assert np.max(np.abs(fmean - gpo.fit["fmean"][lo:hi])) < 1e-8 * np.max(np.abs(gpo.fit["fmean"]))


# ------------------------------------------------------ uneven ownership (a rank without blocks)
def _uneven_worker(rank, world, port, q, n):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import gplite_b200 as g
    from gplite_b200.dist import dist_energy_grad, dist_fit_ep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(n)
        X = rng.normal(size=(n, 2))
        y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.2 * rng.normal(size=n)
        gp = g.gp_init(cfs=g.cf_matern32(lscale=0.9), lik=g.lik_gaussian(sigma=0.3))
        e, gr = dist_energy_grad(gp, X, y, nb=128, ops=TorchOps(gp), dist=dist, device="cpu")
        yb = (y > 0).astype(float)
        gpe = g.gp_init(cfs=g.cf_matern32(lscale=0.9), lik=g.lik_bernoulli(), approx=g.approx_ep())
        rese = dist_fit_ep(gpe, X, yb, nb=128, ops=TorchOps(gpe), dist=dist, device="cpu")
        q.put((rank, e, gr, rese["log_evidence"], rese["iters"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(3, 600), (4, 300)])
def test_distributed_driver_uneven_ownership(world, n):
    """5 blocks on 3 ranks, and 3 blocks on 4 ranks (one rank owns nothing but still takes part in
    every broadcast / all-reduce): energy gradient and EP must still equal the oracle."""
    import torch.multiprocessing as mp
    import gplite_b200 as g
    from oracle import gplite_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_uneven_worker, args=(r, world, port, q, n)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(n)
    X = rng.normal(size=(n, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.2 * rng.normal(size=n)
    eo, go = o.energy_grad_gaussian_full(g.gp_init(cfs=g.cf_matern32(lscale=0.9), lik=g.lik_gaussian(sigma=0.3)), X, y)
    gpe = g.gp_init(cfs=g.cf_matern32(lscale=0.9), lik=g.lik_bernoulli(), approx=g.approx_ep())
    o.gp_fit(gpe, X, (y > 0).astype(float))
    for rank, e, gr, lee, ite in out:
        assert abs(e - eo) < 1e-9 * abs(eo)
        assert np.max(np.abs(gr - go)) < 1e-6 * max(1.0, np.max(np.abs(go)))
        assert ite == gpe.fit["iters"]
        assert abs(lee - gpe.fit["log_evidence"]) < 1e-8 * abs(gpe.fit["log_evidence"])

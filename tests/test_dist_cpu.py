"""CPU, world_size 2 over gloo: host-side logic of the multi-GPU paths that is still driven from the
host -- the row-sharded FITC driver (gplite_b200/dist.py: one m x m all-reduce per Newton step) with
the C-ABI calls replaced by a torch-CPU stand-in that lives only here, and the NCCL-id rendezvous of
DistComm.  (The sharded full-GP schedule is C++ inside the library; its CPU test is
tests/test_dist_sched_cpu.py.)  Results must equal the oracle's gp_fit."""
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
    """likelihood stand-ins shared with TorchFitcOps (test-only)"""

    def __init__(self, gp):
        self.gp = gp

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


# ------------------------------------------------------------------------------ NCCL id rendezvous
def _id_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gplite_b200.dist import DistComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, DistComm.exchange_id(dist)))
    finally:
        dist.destroy_process_group()


def test_nccl_id_reaches_every_rank():
    """rank 0 draws the id inside the library (gpl_dist_unique_id -> ncclGetUniqueId, no GPU needed),
    the process group only carries its 128 bytes"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(out[0]) == 128 and out[0] == out[1] and any(out[0])

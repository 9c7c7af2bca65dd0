"""GPU: the multi-GPU entries of the C ABI (gpl_dist_fit_full, gpl_dist_energy_grad: csrc/dist.cu over
csrc/dist_sched.hpp) against the oracle.  On one device the communicator has world size 1 (same code
path: block-cyclic K build, look-ahead panel factorisation on the side stream, staircase trailing
updates, ragged last block, block substitution, the two inverse sweeps, block traces); with two or
more visible GPUs `test_nccl_two_ranks_equal_single_gpu` spawns one process per GPU and checks the
NCCL path (the driver's bench run repeats that check at every N, see bench.py `sharded.parity`).
The schedule itself is also run on CPUs over gloo: tests/test_dist_sched_cpu.py."""
import copy
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import gplite_b200 as g                      # noqa: E402
from gplite_b200.dist import DistComm, dist_fit_gaussian   # noqa: E402
from oracle import gplite_oracle as o        # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def comm(ctx):
    c = DistComm.single(ctx)
    yield c
    c.close()


@pytest.mark.parametrize("n,nb", [(768, 128), (700, 128), (1111, 256), (1500, 256), (2048, 1024)])   # 1111: odd n
def test_driver_single_rank_matches_oracle(comm, n, nb):
    rng = np.random.default_rng(n)
    X = rng.normal(size=(n, 3))
    y = np.sin(X[:, 0]) + 0.4 * X[:, 1] + 0.3 * rng.normal(size=n)
    gp = g.gp_init(cfs=[g.cf_matern32(lscale=1.1), g.cf_lin(magn=0.3)], lik=g.lik_gaussian(sigma=0.4))
    gpo = copy.deepcopy(gp)
    res = dist_fit_gaussian(gp, X, y, nb=nb, comm=comm)
    o.gp_fit(gpo, X, y)
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    a = res["alpha"]
    assert np.max(np.abs(a - gpo.fit["alpha"])) < 1e-6 * np.max(np.abs(gpo.fit["alpha"]))
    f = res["fmean"]
    assert np.max(np.abs(f - gpo.fit["fmean"])) < 1e-8 * np.max(np.abs(gpo.fit["fmean"]))


@pytest.mark.parametrize("lik", ["bernoulli", "poisson", "binomial"])
def test_laplace_driver_single_rank_matches_oracle(comm, lik):
    """SURVEY 8e "Laplace site sweeps": the Newton loop over the block-cyclic factor, site kernels
    through gpl_lik_*_dev on replicated n-vectors."""
    from gplite_b200.dist import dist_fit_laplace
    rng = np.random.default_rng(21)
    n = 900
    X = rng.normal(size=(n, 3))
    f = 1.5 * X[:, 0] - X[:, 1]
    kw = {}
    if lik == "bernoulli":
        y = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(float)
        L = g.lik_bernoulli()
    elif lik == "poisson":
        y = rng.poisson(np.exp(0.5 * f + 0.3)).astype(float)
        L = g.lik_poisson()
    else:
        kw["trials"] = rng.integers(2, 12, size=n).astype(float)
        y = rng.binomial(kw["trials"].astype(int), 1 / (1 + np.exp(-f))).astype(float)
        L = g.lik_binomial(link="probit")
    gp = g.gp_init(cfs=[g.cf_sexp(lscale=1.1), g.cf_const(magn=0.3)], lik=L, approx=g.approx_laplace())
    gpo = copy.deepcopy(gp)
    res = dist_fit_laplace(gp, X, y, nb=256, comm=comm, **kw)
    o.gp_fit(gpo, X, y, **kw)
    assert res["iters"] == gpo.fit["iters"] and res["iters"] > 2
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    a = res["alpha"]
    assert np.max(np.abs(a - gpo.fit["alpha"])) < 1e-6 * np.max(np.abs(gpo.fit["alpha"]))
    fm = res["fmean"]
    assert np.max(np.abs(fm - gpo.fit["fmean"])) < 1e-8 * np.max(np.abs(gpo.fit["fmean"]))


@pytest.mark.parametrize("n,nb", [(768, 128), (1100, 256), (1101, 128), (2048, 1024)])   # 1101: odd n
def test_energy_grad_driver_single_rank_matches_oracle(comm, n, nb):
    """SURVEY 8e "C^-1 + trace term": block-cyclic trtri / lauum sweeps + gpl_grad_trace_block_dev."""
    from gplite_b200.dist import dist_energy_grad
    rng = np.random.default_rng(n + 1)
    X = rng.normal(size=(n, 3))
    y = np.sin(X[:, 0]) + 0.4 * X[:, 1] + 0.3 * rng.normal(size=n)
    gp = g.gp_init(cfs=[g.cf_matern32(lscale=1.1), g.cf_sexp(lscale=0.7, magn=0.5) * g.cf_lin(prior_magn=g.prior_fixed())],
                   lik=g.lik_gaussian(sigma=0.4))
    gpo = copy.deepcopy(gp)
    e, gr = dist_energy_grad(gp, X, y, nb=nb, comm=comm)
    eo, go = o.energy_grad_gaussian_full(gpo, X, y)
    assert abs(e - eo) < 1e-8 * abs(eo)
    assert np.max(np.abs(gr - go)) < 1e-7 * max(1.0, np.max(np.abs(go))), (gr, go)
    e1, g1 = g.gp_energy_grad(copy.deepcopy(gpo), X, y)
    assert np.max(np.abs(gr - g1)) < 1e-9 * max(1.0, np.max(np.abs(g1)))


@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
def test_ep_driver_single_rank_matches_oracle(comm, lik):
    """SURVEY 8e "EP site sweeps": parallel EP over the block-cyclic factor; marginal variances from
    the distributed X = L^-1 sweep."""
    from gplite_b200.dist import dist_fit_ep
    rng = np.random.default_rng(23)
    n = 700
    X = rng.normal(size=(n, 3))
    f = 1.5 * X[:, 0] - X[:, 1]
    if lik == "bernoulli":
        y, L = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(float), g.lik_bernoulli(link="probit")
    else:
        y, L = rng.poisson(np.exp(0.5 * f + 0.3)).astype(float), g.lik_poisson()
    gp = g.gp_init(cfs=[g.cf_matern32(lscale=1.1), g.cf_const(magn=0.3)], lik=L, approx=g.approx_ep(damping=0.8))
    gpo = copy.deepcopy(gp)
    res = dist_fit_ep(gp, X, y, nb=256, comm=comm)
    o.gp_fit(gpo, X, y)
    assert res["iters"] == gpo.fit["iters"] and res["iters"] > 2
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    for k in ("fmean", "fvar", "alpha"):
        a = res[k]
        assert np.max(np.abs(a - gpo.fit[k])) < 1e-6 * np.max(np.abs(gpo.fit[k])), k


@pytest.mark.parametrize("m", [40, 150])
@pytest.mark.parametrize("streamed", [False, True])
def test_fitc_driver_single_rank_matches_oracle(comm, streamed, m, monkeypatch):
    """m = 40: cond(Kz + jitter I) = 1e6, the oracle's evidence is reproducible to 1e-11 and the 1e-8 bar applies.
    m = 150 inducing points drawn from 1203 crowded training points: cond = 3.4e7 and the ORACLE's own evidence
    moves by ~5e-9 relative when its rows are merely permuted (helpers.oracle_evidence_noise; measured on the
    GPU: resident / streamed / Gram / direct-difference K builds scatter by +-1.4e-8 around the oracle with K
    entries equal to 9e-16, profiles/r02d_fitc_illconditioned_diag.jsonl).  There the bar is the larger of 1e-8
    and 10 x that measured floor (the blocked GPU factorisation solves its panels with explicit 128-block inverses,
    whose constant in front of cond * eps is larger than LAPACK's substitution), itself required to stay < 5e-8."""
    from gplite_b200.dist import dist_fit_fitc_gaussian
    from helpers import oracle_evidence_noise
    monkeypatch.setenv("GPL_FITC_STREAM_GB", "0" if streamed else "1000")
    monkeypatch.setenv("GPL_FITC_CHUNK", "512")
    rng = np.random.default_rng(8)
    n = 1203
    X = rng.uniform(-3, 3, size=(n, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.4 * rng.normal(size=n)
    Zi = X[rng.permutation(n)[:m]].copy()
    mk = lambda Z: g.gp_init(cfs=[g.cf_sexp(lscale=1.1), g.cf_const(magn=0.2)], lik=g.lik_gaussian(sigma=0.4),
                             method=g.method_fitc(inducing=Z))
    gp = mk(Zi)
    gpo = copy.deepcopy(gp)
    res = dist_fit_fitc_gaussian(gp, X, y, comm=comm)
    o.gp_fit(gpo, X, y)

    def oracle_evidence(rows, ind):
        gpp = mk(Zi[ind])
        o.gp_fit(gpp, X[rows], y[rows])
        return gpp.fit["log_evidence"]

    floor = oracle_evidence_noise(oracle_evidence, n, m, reps=6) if m > 40 else 0.0
    assert floor < 5e-8                                         # the instance stays measurable
    tol = max(1e-8, 10.0 * floor)
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < tol * abs(gpo.fit["log_evidence"])
    assert np.max(np.abs(res["alpha"] - gpo.fit["alpha"])) < 1e-6 * np.max(np.abs(gpo.fit["alpha"]))
    assert np.max(np.abs(res["fmean"] - gpo.fit["fmean"])) < tol * np.max(np.abs(gpo.fit["fmean"]))
    Xt = rng.uniform(-3, 3, size=(40, 2))
    p, po = g.gp_pred(gp, Xt, var=True), o.gp_pred(gpo, Xt, var=True)      # the rank-local handle predicts
    assert np.max(np.abs(p["mean"] - po["mean"])) < tol * np.max(np.abs(po["mean"]))
    assert np.max(np.abs(p["var"] - po["var"])) < 1e-7 * np.max(np.abs(po["var"]))


@pytest.mark.parametrize("approx", ["laplace", "ep"])
def test_fitc_laplace_ep_driver_single_rank_matches_oracle(comm, approx):
    """Row-sharded FITC entry with the Newton / EP loops (world 1 here; 2 ranks in the NCCL test below)."""
    from gplite_b200.dist import dist_fit_fitc
    rng = np.random.default_rng(9)
    n, m = 1100, 120
    X = rng.uniform(-3, 3, size=(n, 2))
    f = np.sin(X[:, 0]) + 0.3 * X[:, 1]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-2 * f))).astype(float)
    Zi = X[rng.permutation(n)[:m]].copy()
    ap = g.approx_laplace() if approx == "laplace" else g.approx_ep()
    gp = g.gp_init(cfs=g.cf_matern32(lscale=1.2), lik=g.lik_bernoulli(), method=g.method_fitc(inducing=Zi), approx=ap)
    gpo = copy.deepcopy(gp)
    res = dist_fit_fitc(gp, X, y, comm=comm)
    o.gp_fit(gpo, X, y)
    assert res["iters"] == gpo.fit["iters"] and res["iters"] > 2
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    assert np.max(np.abs(res["fmean"] - gpo.fit["fmean"])) < 1e-7 * np.max(np.abs(gpo.fit["fmean"]))
    if approx == "ep":
        assert np.max(np.abs(res["fvar"] - gpo.fit["fvar"])) < 1e-7 * np.max(np.abs(gpo.fit["fvar"]))


def test_config5_kernel_at_8192(comm):
    """BASELINE config 5's model (cf_matern32, d = 8, lik_gaussian; SURVEY 8d's generator) through the
    sharded path at n = 8192, nb = 1024: log evidence, alpha, energy gradient vs the oracle."""
    from gplite_b200.dist import dist_energy_grad
    rng = np.random.default_rng(5)
    n, d = 8192, 8
    X = rng.normal(size=(n, d))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    gp = g.gp_init(cfs=g.cf_matern32(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
    gpo = copy.deepcopy(gp)
    e, gr, alpha = dist_energy_grad(gp, X, y, nb=1024, comm=comm, want_alpha=True)
    eo, go = o.energy_grad_gaussian_full(gpo, X, y)
    o.gp_fit(gpo, X, y)
    assert abs(e - eo) < 1e-8 * abs(eo)
    assert np.max(np.abs(gr - go)) < 1e-8 * np.max(np.abs(go)), (gr, go)
    assert np.max(np.abs(alpha - gpo.fit["alpha"])) < 1e-6 * np.max(np.abs(gpo.fit["alpha"]))


def test_not_pd_reports_pivot_and_bad_args(comm):
    from gplite_b200 import _lib
    X = np.array([[0.0], [1.0], [np.nan], [2.0]])
    gp = g.gp_init(cfs=g.cf_sexp(), lik=g.lik_gaussian())
    with pytest.raises(g.GplError):
        dist_fit_gaussian(gp, X, np.zeros(4), nb=128, comm=comm)
    with pytest.raises(g.GplError):
        dist_fit_gaussian(gp, np.zeros((4, 1)), np.zeros(4), nb=100, comm=comm)      # nb not a multiple of 128


# ------------------------------------------------------------------------------ NCCL, >= 2 GPUs
def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import gplite_b200 as g
    from gplite_b200.dist import DistComm, dist_energy_grad, dist_fit_ep
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # only carries the 128-byte NCCL id
    try:
        comm = DistComm.from_torch(dist, device=rank)
        rng = np.random.default_rng(5)
        n = 3000
        X = rng.normal(size=(n, 4))
        y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + 0.5 * rng.normal(size=n)
        gp = g.gp_init(cfs=g.cf_matern32(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
        e, gr, alpha = dist_energy_grad(gp, X, y, nb=256, comm=comm, want_alpha=True)
        yb = (y > 0.5).astype(float)
        gpe = g.gp_init(cfs=g.cf_matern32(lscale=1.3), lik=g.lik_bernoulli(), approx=g.approx_ep())
        rese = dist_fit_ep(gpe, X[:1500], yb[:1500], nb=128, comm=comm)
        from gplite_b200.dist import dist_fit_fitc
        Zi = X[:: n // 64][:64].copy()
        resf = {}
        for nm, ap in (("laplace", g.approx_laplace()), ("ep", g.approx_ep())):
            gpf = g.gp_init(cfs=g.cf_sexp(lscale=1.1), lik=g.lik_bernoulli(), method=g.method_fitc(inducing=Zi), approx=ap)
            r = dist_fit_fitc(gpf, X[:2001], yb[:2001], comm=comm)           # odd n: uneven row shards
            resf[nm] = (r["log_evidence"], r["iters"], r["fmean"], g.gp_pred(gpf, X[:7], var=True)["var"])
        q.put((rank, e, gr, alpha, rese["log_evidence"], rese["iters"], rese["fvar"], resf))
        comm.close()
    finally:
        dist.destroy_process_group()


def test_nccl_two_ranks_equal_single_gpu(ctx):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the bench's sharded.parity record covers this on the driver's N > 1 runs)")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [mpc.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    n = 3000
    X = rng.normal(size=(n, 4))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + 0.5 * rng.normal(size=n)
    gp = g.gp_init(cfs=g.cf_matern32(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
    e1, g1 = g.gp_energy_grad(gp, X, y, ctx=ctx)
    gpe = g.gp_init(cfs=g.cf_matern32(lscale=1.3), lik=g.lik_bernoulli(), approx=g.approx_ep())
    g.gp_fit(gpe, X[:1500], (y > 0.5).astype(float)[:1500], ctx=ctx)
    yb = (y > 0.5).astype(float)
    Zi = X[:: n // 64][:64].copy()
    single = {}
    for nm, ap in (("laplace", g.approx_laplace()), ("ep", g.approx_ep())):
        gpf = g.gp_init(cfs=g.cf_sexp(lscale=1.1), lik=g.lik_bernoulli(), method=g.method_fitc(inducing=Zi), approx=ap)
        g.gp_fit(gpf, X[:2001], yb[:2001], ctx=ctx)
        single[nm] = (gpf.fit["log_evidence"], gpf.fit["iters"], gpf.fit["fmean"], g.gp_pred(gpf, X[:7], var=True)["var"])
    for rank, e, gr, alpha, lee, ite, fvar, resf in outs:
        for nm in ("laplace", "ep"):
            le2, it2, fm2, pv2 = resf[nm]
            le1, it1, fm1, pv1 = single[nm]
            assert it2 == it1 and abs(le2 - le1) < 1e-10 * abs(le1)
            assert np.max(np.abs(fm2 - fm1)) < 1e-9 * np.max(np.abs(fm1))
            assert np.max(np.abs(pv2 - pv1)) < 1e-8 * np.max(np.abs(pv1))
        assert abs(e - e1) < 1e-10 * abs(e1)
        assert np.max(np.abs(gr - g1)) < 1e-9 * np.max(np.abs(g1))
        assert ite == gpe.fit["iters"]
        assert abs(lee - gpe.fit["log_evidence"]) < 1e-9 * abs(gpe.fit["log_evidence"])
        assert np.max(np.abs(fvar - gpe.fit["fvar"])) < 1e-8 * np.max(np.abs(gpe.fit["fvar"]))

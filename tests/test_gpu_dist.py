"""GPU: the multi-GPU driver (gplite_b200/dist.py) with the real C-ABI ops on one device
(world size 1 through a trivial stand-in for torch.distributed) against the oracle: exercises the
block-column K build, in-place B, panel factor / TRSM / trailing updates, ragged last block and
the blocked triangular solves.  The NCCL transport itself is covered by tools/dist_fit_bench.py
--check on 2+ GPUs and by the gloo test in tests/test_dist_cpu.py."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import gplite_b200 as g                      # noqa: E402
from gplite_b200.dist import dist_fit_gaussian   # noqa: E402
from oracle import gplite_oracle as o        # noqa: E402


class OneRank:
    class ReduceOp:
        MAX = "max"
        SUM = "sum"
        MIN = "min"

    @staticmethod
    def get_rank():
        return 0

    @staticmethod
    def get_world_size():
        return 1

    @staticmethod
    def broadcast(t, src=0, async_op=False):
        return None

    @staticmethod
    def all_reduce(t, op=None):
        return None


@pytest.mark.parametrize("n,nb", [(768, 128), (700, 128), (1500, 256), (2048, 1024)])
def test_driver_single_rank_matches_oracle(n, nb):
    rng = np.random.default_rng(n)
    X = rng.normal(size=(n, 3))
    y = np.sin(X[:, 0]) + 0.4 * X[:, 1] + 0.3 * rng.normal(size=n)
    gp = g.gp_init(cfs=[g.cf_matern32(lscale=1.1), g.cf_lin(magn=0.3)], lik=g.lik_gaussian(sigma=0.4))
    gpo = copy.deepcopy(gp)
    res = dist_fit_gaussian(gp, X, y, nb=nb, dist=OneRank)
    o.gp_fit(gpo, X, y)
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    a = res["alpha"].cpu().numpy()
    assert np.max(np.abs(a - gpo.fit["alpha"])) < 1e-6 * np.max(np.abs(gpo.fit["alpha"]))
    f = res["fmean"].cpu().numpy()
    assert np.max(np.abs(f - gpo.fit["fmean"])) < 1e-8 * np.max(np.abs(gpo.fit["fmean"]))


@pytest.mark.parametrize("lik", ["bernoulli", "poisson", "binomial"])
def test_laplace_driver_single_rank_matches_oracle(lik):
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
    res = dist_fit_laplace(gp, X, y, nb=256, dist=OneRank, **kw)
    o.gp_fit(gpo, X, y, **kw)
    assert res["iters"] == gpo.fit["iters"] and res["iters"] > 2
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    a = res["alpha"].cpu().numpy()
    assert np.max(np.abs(a - gpo.fit["alpha"])) < 1e-6 * np.max(np.abs(gpo.fit["alpha"]))
    fm = res["fmean"].cpu().numpy()
    assert np.max(np.abs(fm - gpo.fit["fmean"])) < 1e-8 * np.max(np.abs(gpo.fit["fmean"]))


@pytest.mark.parametrize("n,nb", [(768, 128), (1100, 256), (2048, 1024)])
def test_energy_grad_driver_single_rank_matches_oracle(n, nb):
    """SURVEY 8e "C^-1 + trace term": block-cyclic trtri / lauum sweeps + gpl_grad_trace_block_dev."""
    from gplite_b200.dist import dist_energy_grad
    rng = np.random.default_rng(n + 1)
    X = rng.normal(size=(n, 3))
    y = np.sin(X[:, 0]) + 0.4 * X[:, 1] + 0.3 * rng.normal(size=n)
    gp = g.gp_init(cfs=[g.cf_matern32(lscale=1.1), g.cf_sexp(lscale=0.7, magn=0.5) * g.cf_lin(prior_magn=g.prior_fixed())],
                   lik=g.lik_gaussian(sigma=0.4))
    gpo = copy.deepcopy(gp)
    e, gr = dist_energy_grad(gp, X, y, nb=nb, dist=OneRank)
    eo, go = o.energy_grad_gaussian_full(gpo, X, y)
    assert abs(e - eo) < 1e-8 * abs(eo)
    assert np.max(np.abs(gr - go)) < 1e-7 * max(1.0, np.max(np.abs(go))), (gr, go)
    e1, g1 = g.gp_energy_grad(copy.deepcopy(gpo), X, y)
    assert np.max(np.abs(gr - g1)) < 1e-9 * max(1.0, np.max(np.abs(g1)))


@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
def test_ep_driver_single_rank_matches_oracle(lik):
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
    res = dist_fit_ep(gp, X, y, nb=256, dist=OneRank)
    o.gp_fit(gpo, X, y)
    assert res["iters"] == gpo.fit["iters"] and res["iters"] > 2
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    for k in ("fmean", "fvar", "alpha"):
        a = res[k].cpu().numpy()
        assert np.max(np.abs(a - gpo.fit[k])) < 1e-6 * np.max(np.abs(gpo.fit[k])), k


def test_fitc_driver_single_rank_matches_oracle():
    from gplite_b200.dist import dist_fit_fitc_gaussian
    rng = np.random.default_rng(8)
    n, m = 1203, 150
    X = rng.uniform(-3, 3, size=(n, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.4 * rng.normal(size=n)
    Zi = X[rng.permutation(n)[:m]].copy()
    gp = g.gp_init(cfs=[g.cf_sexp(lscale=1.1), g.cf_const(magn=0.2)], lik=g.lik_gaussian(sigma=0.4),
                   method=g.method_fitc(inducing=Zi))
    gpo = copy.deepcopy(gp)
    res = dist_fit_fitc_gaussian(gp, X, y, dist=OneRank, chunk_rows=512)
    o.gp_fit(gpo, X, y)
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    a = res["alpha_local"].cpu().numpy()
    assert np.max(np.abs(a - gpo.fit["alpha"])) < 1e-6 * np.max(np.abs(gpo.fit["alpha"]))
    f = res["fmean_local"].cpu().numpy()
    assert np.max(np.abs(f - gpo.fit["fmean"])) < 1e-8 * np.max(np.abs(gpo.fit["fmean"]))


def test_fitc_laplace_driver_single_rank_matches_oracle():
    """Row-sharded FITC with the Laplace Newton loop on the local rows (gpl_lik_*_dev)."""
    from gplite_b200.dist import dist_fit_fitc
    rng = np.random.default_rng(9)
    n, m = 1100, 120
    X = rng.uniform(-3, 3, size=(n, 2))
    f = np.sin(X[:, 0]) + 0.3 * X[:, 1]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-2 * f))).astype(float)
    Zi = X[rng.permutation(n)[:m]].copy()
    gp = g.gp_init(cfs=g.cf_matern32(lscale=1.2), lik=g.lik_bernoulli(), method=g.method_fitc(inducing=Zi))
    gpo = copy.deepcopy(gp)
    res = dist_fit_fitc(gp, X, y, dist=OneRank, chunk_rows=512)
    o.gp_fit(gpo, X, y)
    assert res["iters"] == gpo.fit["iters"] and res["iters"] > 2
    assert abs(res["log_evidence"] - gpo.fit["log_evidence"]) < 1e-8 * abs(gpo.fit["log_evidence"])
    fm = res["fmean_local"].cpu().numpy()
    assert np.max(np.abs(fm - gpo.fit["fmean"])) < 1e-7 * np.max(np.abs(gpo.fit["fmean"]))

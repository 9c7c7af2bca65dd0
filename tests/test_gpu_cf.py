"""GPU: covariance matrices through the C ABI (gpl_cf_eval, gpl_cf_*_c) against the reference's
own compiled loops (golden vectors from oracle/_ref) and the oracle's eval_cf composition.
North-star tolerance: K entries within 1e-12 relative."""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gplite_b200 as g                      # noqa: E402
from gplite_b200 import _lib                 # noqa: E402
from oracle import gplite_oracle as o        # noqa: E402
from helpers import model_grid, rel_err      # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-12


def _native(ctx, kind, x1, x2, p, diag_only=False):
    fn = getattr(ctx.lib, {"cf_sexp": "gpl_cf_sexp_c", "cf_matern32": "gpl_cf_matern32_c",
                           "cf_matern52": "gpl_cf_matern52_c", "cf_nn": "gpl_cf_nn_c"}[kind])
    x1 = np.asfortranarray(x1)
    x2 = np.asfortranarray(x2)
    n1, d = x1.shape
    n2 = x2.shape[0]
    out = np.empty(min(n1, n2)) if diag_only else np.empty((n1, n2), order="F")
    args = [ctx.h, _lib._ptr(x1), n1, _lib._ptr(x2), n2, d] + [float(v) for v in p] + \
           [1 if diag_only else 0, _lib._ptr(out)]
    _lib.check(fn(*args))
    return out


def test_native_entry_points_match_reference_golden(ctx):
    z = np.load(os.path.join(GOLD, "cf_native.npz"))
    for d in (1, 3, 8):
        x1, x2 = z[f"x1_d{d}"], z[f"x2_d{d}"]
        for kind in ("cf_sexp", "cf_matern32", "cf_matern52", "cf_nn"):
            p = z[f"{kind}_d{d}_p"]
            pp = p if kind == "cf_nn" else p[:2]
            assert rel_err(_native(ctx, kind, x1, x2, pp), z[f"{kind}_d{d}_full"]) < TOL
            assert rel_err(_native(ctx, kind, x1, x2, pp, True), z[f"{kind}_d{d}_diag"]) < TOL
            # symmetric fast path (mirror tiles) against the same reference output
            cf = {"cf_sexp": g.cf_sexp, "cf_matern32": g.cf_matern32, "cf_matern52": g.cf_matern52}.get(kind)
            if cf is not None:
                K = g.eval_cf(cf(lscale=p[0], magn=p[1]), x1, ctx=ctx)
            else:
                K = g.eval_cf(g.cf_nn(sigma0=p[0], sigma=p[1], magn=p[2], normalize=False), x1, ctx=ctx)
            assert rel_err(K, z[f"{kind}_d{d}_sym"]) < TOL
            assert np.array_equal(K, K.T)


@pytest.mark.parametrize("n1,n2", [(1, 1), (127, 129), (128, 128), (300, 257), (1000, 3)])
def test_ragged_shapes_and_jitter(ctx, n1, n2):
    rng = np.random.default_rng(n1 * 7 + n2)
    d = 4
    x1 = rng.normal(size=(n1, d))
    x2 = rng.normal(size=(n2, d))
    cf = g.cf_matern52(lscale=1.1, magn=0.9)
    K = g.eval_cf(cf, x1, x2, ctx=ctx)
    assert rel_err(K, o.eval_cf([cf], x1, x2)) < TOL
    Ks = g.eval_cf(cf, x1, jitter=1e-6, ctx=ctx)
    want = o.eval_cf([cf], x1, x1) + 1e-6 * np.eye(n1)
    assert rel_err(Ks, want) < TOL


def test_empty_inputs(ctx):
    cf = g.cf_sexp()
    assert g.eval_cf(cf, np.zeros((0, 2)), np.zeros((5, 2)), ctx=ctx).shape == (0, 5)
    assert g.eval_cf(cf, np.zeros((4, 2)), np.zeros((0, 2)), ctx=ctx).shape == (4, 0)


def test_model_grid_composition(ctx):
    """All 7 cfs, 21 sums, 35 triple products (the grid of tests/testthat/test_basic.R:15-53) on
    2-D inputs with column subsets, normalisation and cfind."""
    rng = np.random.default_rng(11)
    x = rng.normal(size=(150, 1)) * 2 + 0.5
    xt = rng.normal(size=(131, 1))
    for cfs in model_grid():
        gp = g.gp_init(cfs=cfs)
        for cf in gp.cfs:
            g.model.learn_scales(cf, x)
        for a, b in ((x, None), (xt, x)):
            K = g.eval_cf(gp.cfs, a, b, ctx=ctx)
            want = o.eval_cf(gp.cfs, a, a if b is None else b)
            assert np.max(np.abs(K - want)) <= TOL * np.max(np.abs(want)), cfs
            # every entry must also agree relative to the size of ITS OWN terms (a sum with a cf_lin
            # term can cancel, so |K_ij| itself is not the scale of the rounding errors)
            bb = a if b is None else b
            scale = sum(np.abs(o.eval_cf(gp.cfs, a, bb, cfind=[k])) for k in range(len(gp.cfs)))
            ok = np.abs(K - want) <= 1e-11 * np.maximum(scale, 1e-300 + 1e-6 * np.max(scale))
            assert np.all(ok), cfs
        kd = g.eval_cf(gp.cfs, xt, x, diag_only=True, ctx=ctx)
        assert np.max(np.abs(kd - o.eval_cf(gp.cfs, xt, x, diag_only=True))) < TOL * max(1, np.max(np.abs(kd)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfs = [g.cf_sexp(vars=[0], lscale=0.8) * g.cf_lin(vars=[1, 2], normalize=True),
               g.cf_nn(vars=[2]), g.cf_periodic(vars=[1], period=1.3, cf_base=g.cf_matern32()),
               g.cf_const(magn=0.3)]
    X = rng.normal(size=(200, 3)) * [1.0, 3.0, 0.5] + [0.0, 2.0, -1.0]
    Xt = rng.normal(size=(77, 3))
    for cf in cfs:
        g.model.learn_scales(cf, X)
    for cfind in (None, [0], [1, 3], [2]):
        K = g.eval_cf(cfs, Xt, X, cfind=cfind, ctx=ctx)
        want = o.eval_cf(cfs, Xt, X, cfind=cfind)
        assert np.max(np.abs(K - want)) <= TOL * np.max(np.abs(want))


def test_large_symmetric_matches_properties(ctx):
    """Full-size style property check (size-independent): symmetry, unit diagonal + jitter, and a
    sampled comparison against the oracle."""
    rng = np.random.default_rng(2)
    n, d = 4096 + 37, 8
    x = rng.normal(size=(n, d))
    cf = g.cf_matern52(lscale=1.3, magn=1.0)
    K = g.eval_cf(cf, x, jitter=1e-6, ctx=ctx)
    assert np.array_equal(K, K.T)
    assert np.max(np.abs(np.diag(K) - (1.0 + 1e-6))) < 1e-15
    idx = rng.choice(n, size=200, replace=False)
    want = o.cf_native_np("cf_matern52", x[idx], x, 1.3, 1.0)
    sub = K[idx] - 1e-6 * (idx[:, None] == np.arange(n)[None, :])
    assert rel_err(sub, want) < TOL

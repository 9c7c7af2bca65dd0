"""Generates tests/golden/*.npz.  Run in the build container (needs oracle/_ref, i.e. the
reference's own src/cf.cpp compiled verbatim by oracle/Makefile):

    python tests/golden/make_golden.py

cf_native.npz : outputs of the REFERENCE binary (cf_sexp_c, cf_matern32_c, cf_matern52_c, cf_nn_c;
                full and diag_only) on seeded inputs -> pins K entries.
gaussian_kat.npz : for the model grid of tests/testthat/test_basic.R (7 cfs alone, 21 sums,
                35 triple products) the textbook log evidence and predictive mean/sd exactly as
                that test computes them (R&W eqs; tests/testthat/test_basic.R:82-90,119-129), with K
                from the reference binary where a native kind is involved.
"""
import os
import sys
import itertools

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gplite_oracle as o       # noqa: E402
import gplite_b200 as g                     # noqa: E402  (constructors only; no GPU needed)

HERE = os.path.dirname(os.path.abspath(__file__))


def native():
    assert o.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20260001)
    out = {}
    for d in (1, 3, 8):
        x1 = rng.normal(size=(67, d))
        x2 = rng.normal(size=(45, d)) * 1.5 + 0.3
        out[f"x1_d{d}"] = x1
        out[f"x2_d{d}"] = x2
        for kind, p in (("cf_sexp", (0.7, 1.3, 0.0)), ("cf_matern32", (0.9, 0.8, 0.0)),
                        ("cf_matern52", (1.3, 1.1, 0.0)), ("cf_nn", (1.2, 2.5, 0.9))):
            out[f"{kind}_d{d}_p"] = np.array(p)
            out[f"{kind}_d{d}_full"] = o.cf_native_ref(kind, x1, x2, *p)
            out[f"{kind}_d{d}_sym"] = o.cf_native_ref(kind, x1, x1, *p)
            out[f"{kind}_d{d}_diag"] = o.cf_native_ref(kind, x1, x2, *p, diag_only=True)
    np.savez_compressed(os.path.join(HERE, "cf_native.npz"), **out)


def model_grid():
    mk = [lambda: g.cf_const(), lambda: g.cf_lin(), lambda: g.cf_sexp(), lambda: g.cf_matern32(),
          lambda: g.cf_matern52(), lambda: g.cf_nn(), lambda: g.cf_periodic()]
    grid = [[m()] for m in mk]
    grid += [[mk[i](), mk[j]()] for i, j in itertools.combinations(range(7), 2)]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        grid += [[mk[i]() * mk[j]() * mk[k]()] for i, j, k in itertools.combinations(range(7), 3)]
    return grid


def kat():
    rng = np.random.default_rng(1234)
    n, nt = 20, 30
    x = rng.normal(size=n)
    xt = rng.normal(size=nt)
    y = x ** 2 + 0.2 * rng.normal(size=n)
    prng = np.random.default_rng(2309)
    out = dict(x=x, xt=xt, y=y)
    recs = []
    for mi, cfs in enumerate(model_grid()):
        gp = g.gp_init(cfs=cfs, lik=g.lik_gaussian())
        p0 = g.get_param_vector(gp)
        for j in range(3):
            p = p0 + prng.normal(size=len(p0)) * 0.3
            g.set_param(gp, p)
            for cf in gp.cfs:
                o.learn_scales_cf(cf, x[:, None])
            K = o.eval_cf(gp.cfs, x, x)
            Kt = o.eval_cf(gp.cfs, xt, x)
            Ktt = o.eval_cf(gp.cfs, xt, xt)
            Cm = K + gp.lik.sigma ** 2 * np.eye(n)
            L = sla.cholesky(Cm, lower=True)
            aux = sla.solve_triangular(L, y, lower=True)
            le = -0.5 * aux @ aux - np.sum(np.log(np.diag(L))) - n / 2 * np.log(2 * np.pi)
            alpha = sla.solve_triangular(L, aux, lower=True, trans="T")
            a2 = sla.solve_triangular(L, Kt.T, lower=True)
            recs.append((mi, j))
            out[f"m{mi}_{j}_param"] = p
            out[f"m{mi}_{j}_le"] = le
            out[f"m{mi}_{j}_mean"] = Kt @ alpha
            out[f"m{mi}_{j}_sd"] = np.sqrt(np.diag(Ktt - a2.T @ a2))
    out["index"] = np.array(recs)
    np.savez_compressed(os.path.join(HERE, "gaussian_kat.npz"), **out)


if __name__ == "__main__":
    native()
    kat()
    print("golden fixtures written")

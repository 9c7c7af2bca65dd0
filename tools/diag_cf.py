import os, sys, numpy as np, warnings
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
warnings.simplefilter("ignore")
import gplite_b200 as g
from oracle import gplite_oracle as o
from helpers import model_grid, rel_err
ctx = g.default_context()
z = np.load("tests/golden/cf_native.npz")
for d in (1, 3, 8):
    x1, x2 = z[f"x1_d{d}"], z[f"x2_d{d}"]
    for kind in ("cf_sexp", "cf_matern32", "cf_matern52", "cf_nn"):
        p = z[f"{kind}_d{d}_p"]
        cf = {"cf_sexp": g.cf_sexp, "cf_matern32": g.cf_matern32, "cf_matern52": g.cf_matern52}.get(kind)
        c = cf(lscale=p[0], magn=p[1]) if cf else g.cf_nn(sigma0=p[0], sigma=p[1], magn=p[2], normalize=False)
        K = g.eval_cf(c, x1, x2, ctx=ctx)
        want = z[f"{kind}_d{d}_full"]
        re = np.abs(K - want) / np.maximum(np.abs(want), 1e-300)
        i = np.unravel_index(np.argmax(re), re.shape)
        Ks = g.eval_cf(c, x1, ctx=ctx)
        print(d, kind, p, "rect relerr", re.max(), "at", i, "want", want[i], "got", K[i], "| sym relerr", rel_err(Ks, z[f"{kind}_d{d}_sym"]), "symmetric", np.array_equal(Ks, Ks.T), "maxasym", np.max(np.abs(Ks-Ks.T)))
rng = np.random.default_rng(11)
x = rng.normal(size=(150, 1)) * 2 + 0.5
xt = rng.normal(size=(131, 1))
nbad = 0
for cfs in model_grid():
    gp = g.gp_init(cfs=cfs)
    for cf in gp.cfs:
        g.model.learn_scales(cf, x)
    for a, b in ((x, None), (xt, x)):
        K = g.eval_cf(gp.cfs, a, b, ctx=ctx)
        want = o.eval_cf(gp.cfs, a, a if b is None else b)
        e1 = np.max(np.abs(K - want)) / np.max(np.abs(want))
        big = np.abs(want) > 1e-6 * np.max(np.abs(want))
        e2 = rel_err(K[big], want[big])
        if e1 > 1e-12 or e2 > 1e-11:
            nbad += 1
            if nbad < 15: print("GRID FAIL", cfs, "sym" if b is None else "rect", e1, e2)
print("grid failures", nbad)

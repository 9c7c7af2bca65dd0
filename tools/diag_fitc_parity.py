"""Diagnostic for the ill-conditioned FITC parity case (tests/test_gpu_dist.py::test_fitc_driver_single_rank_
matches_oracle: inducing points = a subset of the training points, cond(Kz + jitter) = 3.4e7): log evidence and
K-entry errors of the GPU fit against the oracle, for whatever K-build path the environment selects
(GPL_KBUILD_V1, GPL_KBUILD_DIRECT, GPL_FITC_STREAM_GB).  One JSON line.

    GPL_KBUILD_V1=1 python tools/diag_fitc_parity.py
"""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                       # noqa: E402
from oracle import gplite_oracle as o         # noqa: E402

rng = np.random.default_rng(8)
n, m = 1203, 150
X = rng.uniform(-3, 3, size=(n, 2))
y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.4 * rng.normal(size=n)
Zi = X[rng.permutation(n)[:m]].copy()
gp = g.gp_init(cfs=[g.cf_sexp(lscale=1.1), g.cf_const(magn=0.2)], lik=g.lik_gaussian(sigma=0.4),
               method=g.method_fitc(inducing=Zi))
gpo = copy.deepcopy(gp)
g.gp_fit(gp, X, y)
o.gp_fit(gpo, X, y)
out = {"env": {k: os.environ[k] for k in os.environ if k.startswith("GPL_")},
       "log_evidence": gp.fit["log_evidence"], "oracle": gpo.fit["log_evidence"],
       "diff": gp.fit["log_evidence"] - gpo.fit["log_evidence"],
       "rel": abs(gp.fit["log_evidence"] - gpo.fit["log_evidence"]) / abs(gpo.fit["log_evidence"])}
try:
    Kxz = g.fit_member(gp, "Kxz")
    out["Kxz_max_rel_err"] = float(np.max(np.abs(Kxz - gpo.fit["Kxz"]) / np.abs(gpo.fit["Kxz"])))
    out["Kxz_mean_rel_err"] = float(np.mean((Kxz - gpo.fit["Kxz"]) / np.abs(gpo.fit["Kxz"])))
except Exception as ex:
    out["Kxz"] = repr(ex)[:80]
Lz = g.fit_member(gp, "Kz_chol")
out["Kz_chol_scaled_err"] = float(np.max(np.abs(np.tril(Lz) - np.tril(gpo.fit["Kz_chol"]))) / np.max(np.abs(gpo.fit["Kz_chol"])))
Kz_g, Kz_o = np.tril(Lz) @ np.tril(Lz).T, gpo.fit["Kz_chol"] @ gpo.fit["Kz_chol"].T
out["Kz_recon_max_abs_diff"] = float(np.max(np.abs(Kz_g - Kz_o)))
for k in ("alpha", "fmean"):
    out[k + "_scaled_err"] = float(np.max(np.abs(gp.fit[k] - gpo.fit[k])) / np.max(np.abs(gpo.fit[k])))
print(json.dumps(out), flush=True)

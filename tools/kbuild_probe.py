"""K-build timings (CUDA events on the library's stream) for the config kernels, symmetric and
rectangular, n x n at d = 8 / 10; prints one JSON line per case with GB/s against MEASURED_PEAKS.

    python tools/kbuild_probe.py [n]
"""
import ctypes as C
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                    # noqa: E402
from gplite_b200 import _lib               # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
REPS = int(os.environ.get("GPL_PROBE_REPS", "5"))       # 1 under ncu: one warm-up + one timed launch per case
ctx = _lib.Context(0, 1)      # the legacy default stream, so that torch.cuda.Event sees the kernels
P = lambda t: C.c_void_p(t.data_ptr())
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0
warnings.simplefilter("ignore")
cases = {
    "matern52 d8": ([g.cf_matern52(lscale=1.3)], 8),
    "matern32 d8": ([g.cf_matern32(lscale=1.3)], 8),
    "sexp d8": ([g.cf_sexp(lscale=1.3)], 8),
    "sexp d6 (cfg4 Kxz)": ([g.cf_sexp(lscale=1.0)], 6),
    "sexp*lin d10 (cfg3)": ([g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed())], 10),
    "nn d8": ([g.cf_nn(normalize=False)], 8),
    "lin d8": ([g.cf_lin()], 8),
    "matern52 + lin + const d8": ([g.cf_matern52(lscale=1.3), g.cf_lin(), g.cf_const()], 8),
}
K = torch.empty((n, n), dtype=torch.float64, device="cuda")
for name, (cfs, d) in cases.items():
    if only and not any(o in name for o in only):
        continue
    rng = np.random.default_rng(2)
    x = torch.from_numpy(rng.normal(size=(d, n))).cuda()
    factors, _ = g.model.flatten_cfs(cfs)
    pm = _lib.PackedModel(factors)
    for same in (1, 0):
        def run():
            _lib.check(ctx.lib.gpl_cf_eval_dev(ctx.h, pm.ref(), P(x), n, P(x), n, d, same, 0, 0,
                                               1e-6 if same else 0.0, P(K), n))
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / REPS
        gbs = (8.0 * n * n + 16.0 * d * n) / (ms * 1e-3) / 1e9
        print(json.dumps({"case": name, "shape": "symmetric" if same else "rectangular", "n": n, "ms": round(ms, 4),
                          "gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 4),
                          "kernel": os.environ.get("GPL_KBUILD_V1", "0") == "1" and "v1 (fp64 pipe)" or "v2 (tensor path)",
                          "variant": os.environ.get("GPL_GRAM_VARIANT", "0")}),
              flush=True)

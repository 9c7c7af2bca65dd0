"""Per-launch GEMM log of one Cholesky factorisation and one SPD inverse (run under gpurun):

    GPL_GEMM_LOG is set per phase; writes gpurun_out/gemmlog_{potrf,potri}_<n>.csv
    columns: m,n,k,lower,kskip,tile,ms
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                    # noqa: E402
from gplite_b200 import _lib               # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
ctx = g.default_context()
q = torch.randn(n, 64, dtype=torch.float64, device="cuda")
a0 = q @ q.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
del q
a = a0.clone()
info, ld = C.c_int32(), C.c_double()


def potrf():
    _lib.check(ctx.lib.gpl_dpotrf_dev(ctx.h, n, C.c_void_p(a.data_ptr()), n, C.byref(info), C.byref(ld)))


def potri():
    _lib.check(ctx.lib.gpl_dpotri_dev(ctx.h, n, C.c_void_p(a.data_ptr()), n))


potrf()
potri()                      # warm: lazy kernel loading, workspace
ctx.sync()
a.copy_(a0)
torch.cuda.synchronize()
for name, fn in (("potrf", potrf), ("potri", potri)):
    path = os.path.join(out, "gemmlog_%s_%d.csv" % (name, n))
    if os.path.exists(path):
        os.remove(path)
    os.environ["GPL_GEMM_LOG"] = path
    ctx.profile(True)
    fn()
    r = ctx.profile_read()
    ctx.profile(False)
    print(name, r, "TF %.2f" % (r["gemm_flops"] / r["gemm_ms"] / 1e9))

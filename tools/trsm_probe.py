"""Probe: right-TRSM of a tall panel (FITC shapes) under the automatic tile choice and with each tile
forced; prints ms and TFLOP/s.   python tools/trsm_probe.py [rows] [m]"""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                    # noqa: E402
from gplite_b200 import _lib               # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 61440
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ctx = g.default_context()
q = torch.randn(m, 64, dtype=torch.float64, device="cuda")
A = q @ q.T + m * torch.eye(m, dtype=torch.float64, device="cuda")
info = C.c_int32()
_lib.check(ctx.lib.gpl_dpotrf_dev(ctx.h, m, C.c_void_p(A.data_ptr()), m, C.byref(info), None))
ctx.sync()
B0 = torch.randn(m, rows, dtype=torch.float64, device="cuda")      # column-major rows x m
B = B0.clone()
for cfg, name in ((0, "auto"), (1, "L"), (2, "M"), (3, "S"), (0, "auto")):
    _lib.check(ctx.lib.gpl_ctx_gemm_config(ctx.h, cfg))
    ts = []
    for rep in range(2):
        B.copy_(B0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(ctx.lib.gpl_dtrsm_right_dev(ctx.h, 1, rows, m, C.c_void_p(A.data_ptr()), m,
                                               C.c_void_p(B.data_ptr()), rows))
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(name, "ms %.2f" % (t * 1e3), "TF %.2f" % (rows * m * m / t / 1e12), flush=True)
_lib.check(ctx.lib.gpl_ctx_gemm_config(ctx.h, 0))

"""Tile-choice tuning table for the FP64 GEMM (run under gpurun): for a grid of (m, n, k, lower)
shapes, time the 128/64/32 CTA tiles with per-launch CUDA events and print one JSON line each.

    python tools/gemm_tune.py > gpurun_out/gemm_tune.jsonl
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                    # noqa: E402
from gplite_b200 import _lib               # noqa: E402


def time_shape(ctx, A, B, Cm, m, n, k, lower, cfg, reps=6):
    _lib.check(ctx.lib.gpl_ctx_gemm_config(ctx.h, cfg))
    lda, ldb, ldc = A.shape[1], B.shape[1], Cm.shape[1]

    def f():
        _lib.check(ctx.lib.gpl_dgemm_dev(ctx.h, 0, 1, m, n, k, -1.0, C.c_void_p(A.data_ptr()), lda,
                                         C.c_void_p(B.data_ptr()), ldb, 1.0, C.c_void_p(Cm.data_ptr()), ldc, lower))
    f()
    f()
    ctx.profile(True)
    for _ in range(reps):
        f()
    r = ctx.profile_read()
    ctx.profile(False)
    return r["gemm_ms"] / r["gemm_launches"]


def main():
    ctx = g.default_context()
    big = 16384
    A = torch.randn(8192, big, dtype=torch.float64, device="cuda")     # col-major (big x 4096)
    B = torch.randn(8192, big, dtype=torch.float64, device="cuda")
    Cm = torch.zeros(big, big, dtype=torch.float64, device="cuda")
    sizes = [128, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, 12288, 16384]
    ks = [128, 256, 512, 1024, 2048, 4096, 8192]
    for lower in (0, 1):
        for m in sizes:
            for n in ([m] if lower else [s for s in sizes if s <= m and (s == m or s in (128, 256, 512, 1024, 2048))]):
                for k in ks:
                    if k > 4 * max(m, n) and k > 512:
                        continue
                    t = [time_shape(ctx, A, B, Cm, m, n, k, lower, cfg) for cfg in (1, 2, 3)]
                    _lib.check(ctx.lib.gpl_ctx_gemm_config(ctx.h, 0))
                    ta = time_shape(ctx, A, B, Cm, m, n, k, lower, 0)
                    print(json.dumps(dict(m=m, n=n, k=k, lower=lower, L=round(t[0] * 1e3, 1), M=round(t[1] * 1e3, 1),
                                          S=round(t[2] * 1e3, 1), auto=round(ta * 1e3, 1))), flush=True)


if __name__ == "__main__":
    main()

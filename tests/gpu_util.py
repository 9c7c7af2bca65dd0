"""Helpers for the -m gpu tests: torch is used only as device memory + an independent fp64
reference for the dense linear algebra (the GP maths is checked against oracle/)."""
import ctypes as C

import numpy as np
import torch

from gplite_b200 import _lib


def dev_colmajor(a_np):
    """numpy (m x n) -> torch tensor holding it column-major with an even leading dimension.
    Returns (tensor_T, ld): tensor_T has shape (n, ld) row-major, i.e. element (i, j) of the
    matrix lives at tensor_T[j, i]."""
    m, n = a_np.shape
    ld = (m + 15) // 16 * 16
    t = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    t[:, :m] = torch.from_numpy(np.ascontiguousarray(a_np.T)).cuda()
    return t, ld


def to_np(t, m):
    return t[:, :m].T.cpu().numpy()


def ptr(t):
    return C.c_void_p(t.data_ptr())


def sync(ctx):
    ctx.sync()
    torch.cuda.synchronize()

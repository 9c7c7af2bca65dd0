"""Gauss-Hermite rule of gplite (R/quadrature.R:1-44): nodes are the roots of the probabilists'
Hermite polynomial He_p, weights p! / (p^2 He_{p-1}(x)^2) (they sum to one).  Host-side, p <= 64;
the nodes/weights cross the C ABI in gpl_approx."""
from __future__ import annotations

import math

import numpy as np


def gauss_hermite_points(p: int):
    if p < 1:
        raise ValueError("quadrature order must be >= 1")
    # roots of He_p = eigenvalues of the symmetric Jacobi matrix with off-diagonals sqrt(k)
    off = np.sqrt(np.arange(1, p, dtype=np.float64))
    x = np.linalg.eigvalsh(np.diag(off, 1) + np.diag(off, -1)) if p > 1 else np.zeros(1)
    x = 0.5 * (x - x[::-1])                       # enforce the exact symmetry of the roots
    # two Newton steps on He_p polish the eigenvalues to full precision
    for _ in range(2):
        h0, h1 = np.ones_like(x), x.copy()
        for k in range(1, p):
            h0, h1 = h1, x * h1 - k * h0
        he_p, he_pm1 = (h1, h0) if p >= 1 else (h0, h0)
        x = x - he_p / (p * he_pm1)               # He_p' = p He_{p-1}
    h0, h1 = np.ones_like(x), x.copy()
    for k in range(1, p - 1):
        h0, h1 = h1, x * h1 - k * h0
    he_pm1 = h1 if p >= 2 else h0
    w = math.factorial(p) / (p ** 2 * he_pm1 ** 2)
    return x, w

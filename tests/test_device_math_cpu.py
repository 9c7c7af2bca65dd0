"""CPU: the accuracy budget of the branch-free device math behind the tensor-path K build and the fused gradient
trace (gplite_b200/csrc/gram_math.cuh), checked by restating its arithmetic in numpy from the constants parsed
out of the header itself.  K entries must stay within 1e-12 relative of the reference's libm-based loops
(src/cf.cpp:37-38,75-76,114-115); the budget per entry is exp_m <= 3e-14, sqrt_pos <= 2 ulp, plus the Gram
cancellation bounded by the kernel's tile guard (cf_gram.cu: scaled norms < 250 -> <= 5e-13)."""
import math
import os
import re
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "gplite_b200", "csrc", "gram_math.cuh")).read()


def _table():
    body = SRC[SRC.index("kExpTab256[256] = {") + len("kExpTab256[256] = {"):]
    body = body[:body.index("};")]
    return np.array([float(x) for x in re.findall(r"[0-9]+\.[0-9]+(?:e[-+]?\d+)?|[0-9]+\.(?![0-9])", body)])


def _const(pattern):
    m = re.search(pattern, SRC)
    assert m, pattern
    return float(m.group(1))


def test_exp_table_is_correctly_rounded():
    tab = _table()
    assert tab.shape == (256,)
    exact = np.exp2(np.arange(256) / 256.0)
    assert np.max(np.abs(tab / exact - 1.0)) <= 2.3e-16


def test_exp_m_restated_error_bound():
    """exp_m(u) = 2^(k >> 8) * T[k & 255] * p3(r), k = rint(-u * 256 / ln 2), r = -u - k ln2/256 (two-part)."""
    tab = _table()
    c_scale = _const(r"fma\(u, (-[0-9.]+), magic\)")
    hi = _const(r"fma\(kd, (-[0-9.e-]+), -u\)")
    lo = _const(r"fma\(kd, (-[0-9.e-]+), r\)")
    c3, c2 = [float(x) for x in re.search(r"fma\(([0-9.]+), r, ([0-9.]+)\)", SRC).groups()]
    assert abs(-c_scale - 256.0 / math.log(2.0)) < 1e-12
    assert abs((-hi) + (-lo) - math.log(2.0) / 256.0) < 1e-19
    # the high part is short enough that kd * hi is exact for every k the kernel can see (|k| < 2^19)
    mant = int(math.frexp(-hi)[0] * 2 ** 53)
    assert (mant & ((1 << 30) - 1)) == 0
    rng = np.random.default_rng(0)
    u = np.concatenate([rng.uniform(0, 30, 200000), rng.uniform(0, 708, 200000), rng.uniform(0, 1e-3, 20000),
                        np.arange(0, 2000) * (math.log(2.0) / 512.0)])        # incl. the interval end points
    kd = np.rint(u * c_scale)
    k = kd.astype(np.int64)
    r = kd * hi - u                   # exact product (short hi), one rounding like the device fma
    r = r + kd * lo
    assert np.max(np.abs(r)) <= math.log(2.0) / 512.0 * (1 + 1e-9)
    p = c3 * r + c2
    p = p * r + 1.0
    p = p * r + 1.0
    v = tab[k & 255] * p * np.exp2((k >> 8).astype(np.float64))
    err = v / np.exp(-u) - 1.0
    assert np.max(np.abs(err)) < 3e-14
    assert np.all(np.isfinite(v)) and np.all(v > 0)
    assert v[u == 0.0][0] == 1.0      # k(x, x) = magn^2 exactly


def test_sqrt_pos_restated_error_bound():
    """sqrt_pos: seed y = rsqrt(u)(1 + d), |d| <= 2^-22 (MUFU.RSQ64H), one third-order step."""
    rng = np.random.default_rng(1)
    u = np.exp(rng.uniform(math.log(1e-300), math.log(1e6), 400000))
    for sign in (-1.0, 1.0, None):
        d = (rng.uniform(-1, 1, u.shape) if sign is None else sign) * 2.0 ** -22
        y = (1.0 / np.sqrt(u)) * (1.0 + d)
        g0 = u * y
        e = 1.0 - g0 * y
        p = 0.375 * e + 0.5
        s = g0 + g0 * (p * e)
        assert np.max(np.abs(s / np.sqrt(u) - 1.0)) < 4.5e-16


def test_clamp_constants():
    lo = struct.unpack("<d", struct.pack("<II", 0, int(re.search(r"max\(__double2hiint\(u\), (0x[0-9a-f]+)\)", SRC).group(1), 16)))[0]
    hi = struct.unpack("<d", struct.pack("<II", 0, int(re.search(r"min\(__double2hiint\(x\), (0x[0-9a-f]+)\)", SRC).group(1), 16)))[0]
    assert 1e-301 < lo < 1e-299                       # sqrt_pos never sees 0 * inf
    assert 708.0 <= hi < 708.5 and math.exp(-hi) > 2.3e-308       # exp_m's exponent arithmetic cannot wrap


def test_gram_contraction_cancellation_stays_inside_the_tile_guard():
    """cf_gram.cu forms u = |a|^2 + |b|^2 - 2 a.b on shifted, scaled rows and falls back to direct differences when
    a tile's scaled norms reach 250.  Below the guard the absolute error of u -- which IS the relative error of
    cf_sexp's K and bounds the Matern kinds' -- must stay under 5e-13."""
    rng = np.random.default_rng(2)
    worst = 0.0
    for d in (1, 2, 6, 8, 10, 38):
        for scale in (1.0, 5.0, 11.0, 15.5):
            a = rng.normal(size=(400, d))
            b = a[rng.integers(0, 400, 400)] + rng.normal(size=(400, d)) * rng.choice([1e-6, 1e-2, 1.0], size=(400, 1))
            a = a / np.linalg.norm(a, axis=1, keepdims=True) * rng.uniform(0, scale, (400, 1))
            b = b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-300) * rng.uniform(0, scale, (400, 1))
            na = np.zeros(400)
            nb = np.zeros(400)
            for k in range(d):                      # per-row norms, fma order of the staging loop
                na = na + a[:, k] * a[:, k]
                nb = nb + b[:, k] * b[:, k]
            keep = (na < 250.0) & (nb < 250.0)      # the kernel's guard
            u = na + nb                              # the two augmented columns come first in the DMMA chain
            for k in range(d):
                u = u + a[:, k] * (-2.0 * b[:, k])
            al, bl = a.astype(np.longdouble), b.astype(np.longdouble)
            exact = np.sum((al - bl) ** 2, axis=1)
            err = np.abs(u.astype(np.longdouble) - exact)[keep]
            worst = max(worst, float(err.max()))
    assert worst < 5e-13

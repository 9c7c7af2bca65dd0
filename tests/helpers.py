"""Shared data / model-grid builders for the tests (seeded numpy stand-ins for the R test data:
tests/testthat/test_basic.R:6-11, test_pred.R:6-12, test_normalization.R:6-13)."""
import itertools
import warnings

import numpy as np

import gplite_b200 as g

CF_MAKERS = [lambda: g.cf_const(), lambda: g.cf_lin(), lambda: g.cf_sexp(), lambda: g.cf_matern32(),
             lambda: g.cf_matern52(), lambda: g.cf_nn(), lambda: g.cf_periodic()]


def model_grid(triples=True):
    mk = CF_MAKERS
    grid = [[m()] for m in mk]
    grid += [[mk[i](), mk[j]()] for i, j in itertools.combinations(range(7), 2)]
    if triples:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            grid += [[mk[i]() * mk[j]() * mk[k]()] for i, j, k in itertools.combinations(range(7), 3)]
    return grid


def basic_data(seed=1234, n=20, nt=30):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=n)
    xt = rng.normal(size=nt)
    y = x ** 2 + 0.2 * rng.normal(size=n)
    return x, xt, y


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b|: the relative error norm used for vectors with entries near zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0


def outlier_counts(seed, n=150, n_out=15):
    """Beta-binomial data with gross outliers: drives some Laplace / EP pseudo-variances negative
    (the reference's non-log-concave branch, R/laplace.R:40-117, R/ep.R:50-138)."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, 2))
    f = 1.5 * X[:, 0] - 1.0 * X[:, 1]
    trials = rng.integers(10, 40, size=n).astype(float)
    y = rng.binomial(trials.astype(int), 1 / (1 + np.exp(-f))).astype(float)
    idx = rng.choice(n, n_out, replace=False)
    y[idx] = np.where(f[idx] > 0, 0.0, trials[idx])
    return X, y, trials


def oracle_evidence_noise(fit_evidence, n, m, reps=5, seed=0):
    """Relative spread of the ORACLE's own log evidence under mathematically neutral permutations of the
    training rows and of the inducing points (only the floating-point summation order changes).
    `fit_evidence(row_perm, inducing_perm)` refits the oracle and returns its log evidence.  Nearly singular
    Kz + jitter I (inducing points drawn from crowded training points) makes the FITC evidence -- the
    difference of two log-determinants with eigenvalues down at the jitter -- uncertain at cond(Kz) * eps;
    a parity tolerance below this floor cannot be met by ANY second fp64 implementation, the reference's
    own R code included."""
    rs = np.random.default_rng(seed)
    vals = [fit_evidence(np.arange(n), np.arange(m))]
    for _ in range(reps - 1):
        vals.append(fit_evidence(rs.permutation(n), rs.permutation(m)))
    vals = np.asarray(vals)
    return float((vals.max() - vals.min()) / abs(vals[0]))

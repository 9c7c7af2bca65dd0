"""The R side of the boundary: r_glue/gplite_b200_glue.c compiled (strict C99) against a stand-in
Rinternals.h (tests/r_shim/, R is not installed here) and RUN through that stand-in runtime.

CPU part: the glue builds, R_init_gplite registers the reference's four .Call symbols with the
reference's arities (src/RcppExports.cpp:76-87: 5, 5, 5, 6) and switches dynamic lookup off
(:86), as the reference does.
GPU part: the registered routines are called with R-style arguments (column-major REALSXP matrices,
length-1 scalars, named lists) and must reproduce the reference binary's golden K entries and the
Python host layer's fit / prediction -- i.e. the .Call path works end to end.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF_ARITY = {"_gplite_cf_sexp_c": 5, "_gplite_cf_matern32_c": 5, "_gplite_cf_matern52_c": 5,
             "_gplite_cf_nn_c": 6}            # src/RcppExports.cpp:76-82


@pytest.fixture(scope="module")
def glue(tmp_path_factory):
    import gplite_b200
    gplite_b200.load_library()               # the product library must exist (no fallback)
    out = str(tmp_path_factory.mktemp("glue") / "libgplite_glue_test.so")
    libdir = os.path.join(ROOT, "gplite_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-O1", "-shared",
           "-fPIC", "-I" + os.path.join(ROOT, "tests", "r_shim"), "-I" + os.path.join(ROOT, "include"),
           "-o", out, os.path.join(ROOT, "r_glue", "gplite_b200_glue.c"),
           os.path.join(ROOT, "tests", "r_shim", "rshim.c"), "-L" + libdir, "-lgplite_b200",
           "-Wl,-rpath," + libdir, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    P = C.c_void_p
    lib.shim_routine_name.restype = C.c_char_p
    lib.shim_last_error.restype = C.c_char_p
    for f in ("shim_nil", "shim_real_matrix", "shim_int_vector", "shim_list", "shim_list_get", "shim_call",
              "Rf_ScalarReal", "Rf_ScalarLogical", "Rf_ScalarInteger", "Rf_mkString", "REAL", "VECTOR_ELT"):
        getattr(lib, f).restype = P
    lib.shim_real_matrix.argtypes = [P, C.c_int, C.c_int]
    lib.shim_int_vector.argtypes = [P, C.c_int]
    lib.shim_list.argtypes = [C.c_int]
    lib.shim_list_set.argtypes = [P, C.c_int, C.c_char_p, P]
    lib.shim_list_get.argtypes = [P, C.c_char_p]
    lib.shim_call.argtypes = [C.c_char_p, C.c_int, P]
    lib.Rf_ScalarReal.argtypes = [C.c_double]
    lib.Rf_ScalarLogical.argtypes = [C.c_int]
    lib.Rf_ScalarInteger.argtypes = [C.c_int]
    lib.Rf_mkString.argtypes = [C.c_char_p]
    lib.REAL.argtypes = [P]
    lib.Rf_nrows.argtypes = [P]
    lib.Rf_ncols.argtypes = [P]
    lib.XLENGTH.argtypes = [P]
    lib.XLENGTH.restype = C.c_ssize_t
    lib.Rf_asReal.argtypes = [P]
    lib.Rf_asReal.restype = C.c_double
    lib.shim_run_finalizer.argtypes = [P]
    lib.R_init_gplite(None)
    return lib


def test_registration_matches_reference(glue):
    table = {glue.shim_routine_name(i).decode(): glue.shim_routine_nargs(i)
             for i in range(glue.shim_routine_count())}
    for name, nargs in REF_ARITY.items():
        assert table.get(name) == nargs, (name, table.get(name))
    assert glue.shim_dynamic_symbols() == 0                  # R_useDynamicSymbols(dll, FALSE)
    for extra in ("gplb_fit_full", "gplb_fit_fitc", "gplb_energy_grad", "gplb_pred_full", "gplb_pred_fitc",
                  "gplb_fit_get"):
        assert extra in table
    # the multi-GPU routines
    assert {k: table.get(k) for k in ("gplb_set_device", "gplb_dist_unique_id", "gplb_dist_init", "gplb_dist_free",
                                      "gplb_dist_fit_full", "gplb_dist_energy_grad")} == {
        "gplb_set_device": 1, "gplb_dist_unique_id": 0, "gplb_dist_init": 3, "gplb_dist_free": 1,
        "gplb_dist_fit_full": 10, "gplb_dist_energy_grad": 9}
    ref = "/root/reference/src/RcppExports.cpp"              # only in the build container
    if os.path.exists(ref):
        found = dict((m.group(1), int(m.group(2))) for m in
                     re.finditer(r'\{"(_gplite_\w+)", \(DL_FUNC\) &\w+, (\d+)\}', open(ref).read()))
        assert found == REF_ARITY


def test_unique_id_is_128_raw_bytes(glue):
    """gplb_dist_unique_id() -> the RAWSXP an R host ships to its workers (ncclGetUniqueId needs no GPU)."""
    out = _call(glue, "gplb_dist_unique_id")
    assert out, glue.shim_last_error()
    assert glue.XLENGTH(out) == 128
    glue.RAW.restype = C.c_void_p
    glue.RAW.argtypes = [C.c_void_p]
    raw = bytes((C.c_ubyte * 128).from_address(glue.RAW(out)))
    assert any(raw)
    again = bytes((C.c_ubyte * 128).from_address(glue.RAW(_call(glue, "gplb_dist_unique_id"))))
    assert again != raw                                        # ids are unique per call


def test_r_side_helpers_are_documented():
    """INTEGRATION.md names the glue file and every routine it registers."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "r_glue/gplite_b200_glue.c" in text
    for name in list(REF_ARITY) + ["gplb_fit_full", "gplb_pred_full", "gplb_energy_grad", "gplb_dist_init",
                                   "gplb_dist_fit_full", "gplb_dist_unique_id"]:
        assert name in text


# ------------------------------------------------------------------------------------------- GPU
def _mat(glue, a):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if a.ndim == 1:
        a = a[:, None]
    return glue.shim_real_matrix(a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1])


def _call(glue, name, *args):
    arr = (C.c_void_p * len(args))(*args)
    out = glue.shim_call(name.encode(), len(args), arr)
    assert glue.shim_protect_balance() == 0, "PROTECT / UNPROTECT imbalance in " + name
    return out


def _real(glue, sexp):
    n = glue.XLENGTH(sexp)
    buf = (C.c_double * n).from_address(glue.REAL(sexp))
    return np.array(buf, dtype=np.float64).reshape((glue.Rf_nrows(sexp), glue.Rf_ncols(sexp)), order="F")


def _named(glue, **kw):
    lst = glue.shim_list(len(kw))
    for i, (k, v) in enumerate(kw.items()):
        glue.shim_list_set(lst, i, k.encode(), v)
    return lst


@pytest.mark.gpu
def test_dot_call_cf_matches_reference_golden(glue):
    """.Call(_gplite_cf_*_c, x1, x2, ..., diag_only) -> the reference binary's own output (1e-12)."""
    from helpers import rel_err
    z = np.load(os.path.join(GOLD, "cf_native.npz"))
    x1, x2 = z["x1_d3"], z["x2_d3"]
    for kind in ("cf_sexp", "cf_matern32", "cf_matern52", "cf_nn"):
        p = z[f"{kind}_d3_p"]
        scal = [glue.Rf_ScalarReal(float(v)) for v in (p if kind == "cf_nn" else p[:2])]
        for diag, key in ((0, "full"), (1, "diag")):
            out = _call(glue, f"_gplite_{kind}_c", _mat(glue, x1), _mat(glue, x2), *scal, glue.Rf_ScalarLogical(diag))
            assert out, glue.shim_last_error()
            got = _real(glue, out)
            want = z[f"{kind}_d3_{key}"]
            assert got.size == want.size
            assert rel_err(got.reshape(want.shape, order="F"), want) < 1e-12


@pytest.mark.gpu
def test_dot_call_fit_and_predict_equal_host_layer(glue, ctx):
    """gplb_fit_full / gplb_pred_full / gplb_energy_grad / gplb_fit_get driven as R would drive them,
    against the Python host layer over the same C ABI; the failing-fit path raises an R error."""
    import gplite_b200 as g
    rng = np.random.default_rng(32004)
    n = 200
    x = rng.normal(size=n)
    y = np.sin(3 * x) * np.exp(-np.abs(x)) + 0.1 * rng.normal(size=n)
    xt = np.linspace(-4, 4, 300)
    gp = g.gp_init(cfs=g.cf_sexp(lscale=0.4, magn=1.1), lik=g.lik_gaussian(sigma=0.3))
    g.gp_fit(gp, x, y, ctx=ctx)
    pred = g.gp_pred(gp, xt, var=True)
    e, grad = g.gp_energy_grad(gp, x, y, include_prior=False, ctx=ctx)

    factor = _named(glue, kind=glue.Rf_ScalarInteger(2), term=glue.Rf_ScalarInteger(0),
                    p=_mat(glue, np.array([0.4, 1.1])))
    model = glue.shim_list(1)
    glue.shim_list_set(model, 0, b"", factor)
    lik = _named(glue, kind=glue.Rf_ScalarInteger(0), link=glue.Rf_ScalarInteger(0), param=glue.Rf_ScalarReal(0.3))
    approx = _named(glue, kind=glue.Rf_ScalarInteger(0), maxiter=glue.Rf_ScalarInteger(30), tol=glue.Rf_ScalarReal(1e-4))
    nil = glue.shim_nil()
    fit = _call(glue, "gplb_fit_full", model, lik, approx, _mat(glue, x), _mat(glue, y), nil, nil,
                glue.Rf_ScalarReal(1e-6))
    assert fit, glue.shim_last_error()
    le = glue.Rf_asReal(glue.shim_list_get(fit, b"log_evidence"))
    assert abs(le - gp.fit["log_evidence"]) <= 1e-12 * abs(le)
    assert np.array_equal(_real(glue, glue.shim_list_get(fit, b"alpha"))[:, 0], gp.fit["alpha"])
    handle = glue.shim_list_get(fit, b"handle")
    p = _call(glue, "gplb_pred_full", handle, model, _mat(glue, xt), nil, glue.Rf_ScalarLogical(1))
    assert p, glue.shim_last_error()
    assert np.array_equal(_real(glue, glue.shim_list_get(p, b"mean"))[:, 0], pred["mean"])
    assert np.array_equal(_real(glue, glue.shim_list_get(p, b"var"))[:, 0], pred["var"])
    cchol = _call(glue, "gplb_fit_get", handle, glue.Rf_mkString(b"C_chol"), glue.Rf_ScalarInteger(n),
                  glue.Rf_ScalarInteger(n))
    assert np.array_equal(_real(glue, cchol), g.fit_member(gp, "C_chol"))
    pf = np.array([0, 0, -1], dtype=np.int32)
    ps = np.array([0, 1, 0], dtype=np.int32)
    eg = _call(glue, "gplb_energy_grad", model, lik, _mat(glue, x), _mat(glue, y), glue.Rf_ScalarReal(1e-6),
               glue.shim_int_vector(pf.ctypes.data_as(C.c_void_p), 3),
               glue.shim_int_vector(ps.ctypes.data_as(C.c_void_p), 3))
    assert eg, glue.shim_last_error()
    assert abs(glue.Rf_asReal(glue.shim_list_get(eg, b"energy")) - e) <= 1e-12 * abs(e)
    assert np.allclose(_real(glue, glue.shim_list_get(eg, b"grad"))[:, 0], grad, rtol=1e-12, atol=0)
    glue.shim_run_finalizer(handle)                           # what R's GC would do
    # a NaN input makes B non-PD: method_full aborts with an R error (R/gp_fit.R:95-97)
    xbad = x.copy()
    xbad[5] = np.nan
    bad = _call(glue, "gplb_fit_full", model, lik, approx, _mat(glue, xbad), _mat(glue, y), nil, nil,
                glue.Rf_ScalarReal(1e-6))
    assert not bad
    assert b"gplite_b200" in glue.shim_last_error()


@pytest.mark.gpu
def test_dot_call_multi_gpu_routines_world_one(glue, ctx):
    """The multi-GPU .Call routines driven as an R worker would drive them, with a one-rank communicator (no
    NCCL needed): gplb_dist_fit_full / gplb_dist_energy_grad return what gplb_fit_full / gplb_energy_grad
    return for the same data (the block-cyclic schedule with world = 1 is a different summation order, hence
    1e-10 instead of bit equality), and a released communicator raises an R error."""
    rng = np.random.default_rng(32005)
    n, d = 700, 3
    x = rng.normal(size=(n, d))
    y = np.sin(x[:, 0]) + 0.3 * x[:, 1] + 0.2 * rng.normal(size=n)
    factor = _named(glue, kind=glue.Rf_ScalarInteger(3), term=glue.Rf_ScalarInteger(0),
                    p=_mat(glue, np.array([1.3, 0.9])))                       # cf_matern32 (config 5's kernel)
    model = glue.shim_list(1)
    glue.shim_list_set(model, 0, b"", factor)
    lik = _named(glue, kind=glue.Rf_ScalarInteger(0), link=glue.Rf_ScalarInteger(0), param=glue.Rf_ScalarReal(0.4))
    approx = _named(glue, kind=glue.Rf_ScalarInteger(0), maxiter=glue.Rf_ScalarInteger(30), tol=glue.Rf_ScalarReal(1e-4))
    nil = glue.shim_nil()
    jit = glue.Rf_ScalarReal(1e-6)
    one = _call(glue, "gplb_fit_full", model, lik, approx, _mat(glue, x), _mat(glue, y), nil, nil, jit)
    assert one, glue.shim_last_error()
    dist = _call(glue, "gplb_dist_init", nil, glue.Rf_ScalarInteger(0), glue.Rf_ScalarInteger(1))
    assert dist, glue.shim_last_error()
    many = _call(glue, "gplb_dist_fit_full", dist, model, lik, approx, _mat(glue, x), _mat(glue, y), nil, nil, jit,
                 glue.Rf_ScalarInteger(256))
    assert many, glue.shim_last_error()
    le1 = glue.Rf_asReal(glue.shim_list_get(one, b"log_evidence"))
    leN = glue.Rf_asReal(glue.shim_list_get(many, b"log_evidence"))
    assert abs(le1 - leN) <= 1e-10 * abs(le1)
    a1 = _real(glue, glue.shim_list_get(one, b"alpha"))[:, 0]
    aN = _real(glue, glue.shim_list_get(many, b"alpha"))[:, 0]
    assert np.max(np.abs(a1 - aN)) <= 1e-8 * np.max(np.abs(a1))
    f1 = _real(glue, glue.shim_list_get(one, b"fmean"))[:, 0]
    fN = _real(glue, glue.shim_list_get(many, b"fmean"))[:, 0]
    assert np.max(np.abs(f1 - fN)) <= 1e-9 * np.max(np.abs(f1))
    pf = np.array([0, 0, -1], dtype=np.int32)
    ps = np.array([0, 1, 0], dtype=np.int32)
    ivec = lambda a: glue.shim_int_vector(a.ctypes.data_as(C.c_void_p), len(a))
    eg1 = _call(glue, "gplb_energy_grad", model, lik, _mat(glue, x), _mat(glue, y), jit, ivec(pf), ivec(ps))
    egN = _call(glue, "gplb_dist_energy_grad", dist, model, lik, _mat(glue, x), _mat(glue, y), jit,
                glue.Rf_ScalarInteger(256), ivec(pf), ivec(ps))
    assert eg1 and egN, glue.shim_last_error()
    e1 = glue.Rf_asReal(glue.shim_list_get(eg1, b"energy"))
    eN = glue.Rf_asReal(glue.shim_list_get(egN, b"energy"))
    assert abs(e1 - eN) <= 1e-10 * abs(e1)
    g1 = _real(glue, glue.shim_list_get(eg1, b"grad"))[:, 0]
    gN = _real(glue, glue.shim_list_get(egN, b"grad"))[:, 0]
    assert np.max(np.abs(g1 - gN)) <= 1e-8 * np.max(np.abs(g1))
    _call(glue, "gplb_dist_free", dist)
    gone = _call(glue, "gplb_dist_fit_full", dist, model, lik, approx, _mat(glue, x), _mat(glue, y), nil, nil, jit,
                 glue.Rf_ScalarInteger(256))
    assert not gone and b"released" in glue.shim_last_error()
    glue.shim_run_finalizer(glue.shim_list_get(one, b"handle"))

"""ctypes binding of libgplite_b200.so (include/gplite_b200.h).

The library is the product; this module only marshals arguments.  There is no CPU fallback:
if the shared library is missing, or no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgplite_b200.so")

GPL_OK, GPL_ERR_BAD_ARG, GPL_ERR_NOT_PD, GPL_ERR_NONFINITE, GPL_ERR_CUDA, GPL_ERR_NCCL, \
    GPL_ERR_NOMEM, GPL_ERR_NOT_LOGCONCAVE = range(8)

CF_KINDS = {"cf_const": 0, "cf_lin": 1, "cf_sexp": 2, "cf_matern32": 3, "cf_matern52": 4, "cf_nn": 5}
LIK_KINDS = {"lik_gaussian": 0, "lik_bernoulli": 1, "lik_binomial": 2, "lik_betabinom": 3,
             "lik_poisson": 4}
LINKS = {"identity": 0, "logit": 1, "probit": 2, "log": 3}


class GplError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libgplite_b200 status {status}: {msg}")
        self.status = status


class NotPositiveDefinite(GplError):
    pass


class CfFactor(C.Structure):
    _fields_ = [("kind", C.c_int32), ("term", C.c_int32), ("periodic", C.c_int32),
                ("normalize", C.c_int32), ("nvars", C.c_int32), ("reserved", C.c_int32),
                ("vars", C.POINTER(C.c_int32)), ("means", C.POINTER(C.c_double)),
                ("scales", C.POINTER(C.c_double)), ("p", C.c_double * 3), ("period", C.c_double)]


class CfModel(C.Structure):
    _fields_ = [("nfactors", C.c_int32), ("nterms", C.c_int32), ("factors", C.POINTER(CfFactor))]


class Lik(C.Structure):
    _fields_ = [("kind", C.c_int32), ("link", C.c_int32), ("param", C.c_double)]


class Approx(C.Structure):
    _fields_ = [("kind", C.c_int32), ("maxiter", C.c_int32), ("tol", C.c_double),
                ("damping", C.c_double), ("damping_min", C.c_double), ("quad_order", C.c_int32),
                ("reserved", C.c_int32), ("quad_x", C.POINTER(C.c_double)),
                ("quad_w", C.POINTER(C.c_double))]


class RfSpec(C.Structure):
    _fields_ = [("nterms", C.c_int32), ("num_basis", C.POINTER(C.c_int32)),
                ("weights", C.POINTER(C.c_double)), ("nweights", C.c_int64)]


class FitOut(C.Structure):
    _fields_ = [("log_evidence", C.c_double), ("iters", C.c_int32), ("converged", C.c_int32),
                ("quad_ok", C.c_int32), ("n_bad", C.c_int32), ("accepted", C.c_int32),
                ("reserved", C.c_int32), ("damping_final", C.c_double)]


class DistTimings(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("kbuild_s", "factor_s", "solve_s", "trtri_s", "lauum_s", "trace_s",
                                         "total_s", "nccl_bytes")]


_P = C.c_void_p
_D = C.c_double
_I64 = C.c_int64
_I32 = C.c_int32
_U64 = C.c_uint64

# symbol -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "gpl_abi_version": (C.c_int, []),
    "gpl_last_error": (C.c_char_p, []),
    "gpl_device_count": (C.c_int, []),
    "gpl_ctx_create": (C.c_int, [C.c_int, _U64, C.POINTER(_P)]),
    "gpl_ctx_destroy": (C.c_int, [_P]),
    "gpl_ctx_sync": (C.c_int, [_P]),
    "gpl_ctx_launch_count": (_I64, [_P]),
    "gpl_ctx_profile": (C.c_int, [_P, _I32]),
    "gpl_ctx_gemm_config": (C.c_int, [_P, _I32]),
    "gpl_ctx_profile_read": (C.c_int, [_P, C.POINTER(_D), C.POINTER(_D), C.POINTER(_I64)]),
    "gpl_dev_alloc": (C.c_int, [_P, _U64, C.POINTER(_P)]),
    "gpl_dev_free": (C.c_int, [_P, _P]),
    "gpl_memcpy_h2d": (C.c_int, [_P, _P, _P, _U64]),
    "gpl_memcpy_d2h": (C.c_int, [_P, _P, _P, _U64]),
    "gpl_cf_eval": (C.c_int, [_P, C.POINTER(CfModel), _P, _I64, _P, _I64, _I32, _I32, _U64, _I32, _D, _P]),
    "gpl_cf_eval_dev": (C.c_int, [_P, C.POINTER(CfModel), _P, _I64, _P, _I64, _I32, _I32, _U64, _I32,
                                  _D, _P, _I64]),
    "gpl_cf_eval_block_dev": (C.c_int, [_P, C.POINTER(CfModel), _P, _I64, _I64, _P, _I64, _I64, _I32, _P, _I64]),
    "gpl_cf_sexp_c": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _D, _I32, _P]),
    "gpl_cf_matern32_c": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _D, _I32, _P]),
    "gpl_cf_matern52_c": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _D, _I32, _P]),
    "gpl_cf_nn_c": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _D, _D, _I32, _P]),
    "gpl_dgemm_dev": (C.c_int, [_P, _I32, _I32, _I64, _I64, _I64, _D, _P, _I64, _P, _I64, _D, _P, _I64,
                                _I32]),
    "gpl_dpotrf_dev": (C.c_int, [_P, _I64, _P, _I64, C.POINTER(_I32), C.POINTER(_D)]),
    "gpl_dtrsm_right_dev": (C.c_int, [_P, _I32, _I64, _I64, _P, _I64, _P, _I64]),
    "gpl_dpotri_dev": (C.c_int, [_P, _I64, _P, _I64]),
    "gpl_dtrsm_right_dinv_dev": (C.c_int, [_P, _I32, _I64, _I64, _P, _I64, _P, _P, _I64]),
    "gpl_rowsumsq_dev": (C.c_int, [_P, _I64, _I64, _P, _I64, _P]),
    "gpl_dpotrf_get_dinv_dev": (C.c_int, [_P, _I64, _P]),
    "gpl_dtrsv_dev": (C.c_int, [_P, _I32, _I64, _P, _I64, _P, _P, _P]),
    "gpl_dgemv_dev": (C.c_int, [_P, _I32, _I64, _I64, _P, _I64, _P, _P]),
    "gpl_scale_block_dev": (C.c_int, [_P, _I64, _I64, _P, _I64, _P, _P, _D, _D, _I64]),
    "gpl_lik_pseudodata_la_dev": (C.c_int, [_P, C.POINTER(Lik), _I64, _P, _P, _P, _P, _P, _P]),
    "gpl_grad_trace_block_dev": (C.c_int, [_P, C.POINTER(CfModel), _P, _I64, _I32, _I64, _I64, _P, _I64, _P, _P,
                                           _I32, _P, _P, _P, C.POINTER(_D)]),
    "gpl_pred_full_cov": (C.c_int, [_P, _P, C.POINTER(CfModel), _P, _I64, _U64, _D, _P, _P]),
    "gpl_draw_full": (C.c_int, [_P, _P, C.POINTER(CfModel), _P, _I64, _I32, _U64, _D, _P, _I64, _P]),
    "gpl_fit_rf": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), C.POINTER(Approx), C.POINTER(RfSpec), _P,
                             _I64, _I32, _P, _P, _P, C.POINTER(FitOut), _P, _P, _P, _P, C.POINTER(_P)]),
    "gpl_rf_features": (C.c_int, [_P, _P, C.POINTER(CfModel), _P, _I64, _U64, _P]),
    "gpl_pred_rf": (C.c_int, [_P, _P, C.POINTER(CfModel), _P, _I64, _U64, _P, _P]),
    "gpl_energy_grad_laplace": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), C.POINTER(Approx), _P, _I64, _I32,
                                          _P, _P, _P, _D, _I32, _P, _P, C.POINTER(_D), _P, C.POINTER(FitOut)]),
    "gpl_kmeans": (C.c_int, [_P, _P, _I64, _I32, _P, _I64, _I32, _P, _P, C.POINTER(_I32)]),
    "gpl_lik_ep_sites_dev": (C.c_int, [_P, C.POINTER(Lik), _I64, _P, _P, _P, _P, _P, _P, _P, _D, _I32, _P, _P,
                                       _P, _P, _P, _P, _P, _P, C.POINTER(_I32)]),
    "gpl_colsumsq_dev": (C.c_int, [_P, _I64, _I64, _P, _I64, _P]),
    "gpl_lik_loglik_dev": (C.c_int, [_P, C.POINTER(Lik), _I64, _P, _P, _P, _P, C.POINTER(_D)]),
    "gpl_fit_full": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), C.POINTER(Approx), _P, _I64,
                               _I32, _P, _P, _P, _D, C.POINTER(FitOut), _P, _P, _P, _P, _P, _P, _P,
                               C.POINTER(_P)]),
    "gpl_fit_fitc": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), C.POINTER(Approx), _P, _I64,
                               _I32, _P, _I64, _P, _P, _P, _D, C.POINTER(FitOut), _P, _P, _P, _P, _P,
                               _P, _P, C.POINTER(_P)]),
    "gpl_energy_grad": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), _P, _I64, _I32, _P, _I32, _D,
                                  _I32, _P, _P, C.POINTER(_D), _P, C.POINTER(_P)]),
    "gpl_pred_full": (C.c_int, [_P, _P, C.POINTER(CfModel), _P, _I64, _U64, _P, _P]),
    "gpl_pred_fitc": (C.c_int, [_P, _P, C.POINTER(CfModel), _P, _I64, _U64, _P, _P]),
    "gpl_fit_get": (C.c_int, [_P, _P, C.c_char_p, _P]),
    "gpl_fit_free": (C.c_int, [_P, _P]),
    "gpl_ctx_release_cache": (C.c_int, [_P]),
    "gpl_dist_unique_id": (C.c_int, [_P]),
    "gpl_dist_init": (C.c_int, [_P, _P, _I32, _I32, C.POINTER(_P)]),
    "gpl_dist_destroy": (C.c_int, [_P]),
    "gpl_dist_release": (C.c_int, [_P]),
    "gpl_dist_bcast_bench": (C.c_int, [_P, _I64, _I32, C.POINTER(_D)]),
    "gpl_dist_rank": (C.c_int, [_P]),
    "gpl_dist_world": (C.c_int, [_P]),
    "gpl_dist_fit_full": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), C.POINTER(Approx), _P, _I64, _I32, _P,
                                    _P, _P, _D, _I64, C.POINTER(FitOut), _P, _P, _P, _P, _P, _P, _P,
                                    C.POINTER(DistTimings)]),
    "gpl_dist_fit_fitc": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), C.POINTER(Approx), _P, _I64, _I32, _P,
                                    _I64, _P, _P, _P, _D, C.POINTER(FitOut), _P, _P, _P, _P, _P, _P, _P,
                                    C.POINTER(_P)]),
    "gpl_dist_energy_grad": (C.c_int, [_P, C.POINTER(CfModel), C.POINTER(Lik), _P, _I64, _I32, _P, _D, _I64, _I32,
                                       _P, _P, C.POINTER(_D), _P, _P, C.POINTER(DistTimings)]),
}

_lib = None


def load():
    """Load the shared library (no GPU needed for this; compute calls need one)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  gplite_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status == GPL_OK:
        return
    msg = load().gpl_last_error().decode("utf-8", "replace")
    if status == GPL_ERR_NOT_PD:
        raise NotPositiveDefinite(status, msg)
    raise GplError(status, msg)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=[order, "A"])


class Context:
    """One per (process, device)."""

    def __init__(self, device=0, stream=0):
        self.lib = load()
        h = _P()
        check(self.lib.gpl_ctx_create(int(device), int(stream), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpl_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.gpl_ctx_sync(self.h))

    def release_cache(self):
        check(self.lib.gpl_ctx_release_cache(self.h))

    def launch_count(self):
        return int(self.lib.gpl_ctx_launch_count(self.h))

    def profile(self, enable=True):
        check(self.lib.gpl_ctx_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        ms, fl, nl = C.c_double(), C.c_double(), C.c_int64()
        check(self.lib.gpl_ctx_profile_read(self.h, C.byref(ms), C.byref(fl), C.byref(nl)))
        return dict(gemm_ms=ms.value, gemm_flops=fl.value, gemm_launches=nl.value)

    def alloc(self, nbytes):
        p = _P()
        check(self.lib.gpl_dev_alloc(self.h, int(nbytes), C.byref(p)))
        return p

    def free(self, p):
        check(self.lib.gpl_dev_free(self.h, p))

    def h2d(self, dptr, arr):
        check(self.lib.gpl_memcpy_h2d(self.h, dptr, _ptr(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        check(self.lib.gpl_memcpy_d2h(self.h, _ptr(arr), dptr, arr.nbytes))


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
        ndev = load().gpl_device_count()
        if ndev > 0:
            device %= ndev
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class PackedModel:
    """Keeps the numpy buffers referenced by a CfModel alive."""

    def __init__(self, factors):
        # factors: list of dict(kind, term, periodic, normalize, vars, means, scales, p, period)
        self.keep = []
        arr = (CfFactor * max(len(factors), 1))()
        for i, f in enumerate(factors):
            cf = arr[i]
            cf.kind = CF_KINDS[f["kind"]]
            cf.term = f["term"]
            cf.periodic = 1 if f.get("periodic") else 0
            cf.normalize = 1 if f.get("normalize") else 0
            v = f.get("vars")
            if v is not None:
                va = np.ascontiguousarray(v, dtype=np.int32)
                self.keep.append(va)
                cf.nvars = len(va)
                cf.vars = va.ctypes.data_as(C.POINTER(C.c_int32))
            else:
                cf.nvars = 0
                cf.vars = None
            if cf.normalize and f.get("means") is not None and f.get("scales") is not None:
                m = np.ascontiguousarray(f["means"], dtype=np.float64)
                s = np.ascontiguousarray(f["scales"], dtype=np.float64)
                self.keep += [m, s]
                cf.means = m.ctypes.data_as(C.POINTER(C.c_double))
                cf.scales = s.ctypes.data_as(C.POINTER(C.c_double))
            else:
                cf.normalize = 0
                cf.means = None
                cf.scales = None
            p = list(f["p"]) + [0.0] * (3 - len(f["p"]))
            for k in range(3):
                cf.p[k] = float(p[k])
            cf.period = float(f.get("period", 1.0))
        self.arr = arr
        self.model = CfModel(len(factors), (max(f["term"] for f in factors) + 1) if factors else 0, arr)
        self.nterms = self.model.nterms

    def ref(self):
        return C.byref(self.model)

/*
 * gplite_b200_glue.c -- the R side of the drop-in boundary: .Call stubs that bind
 * libgplite_b200.so (include/gplite_b200.h) into the gplite package.  Pure C against
 * Rinternals.h, no Rcpp.  It replaces src/RcppExports.cpp of the reference:
 *
 *   _gplite_cf_sexp_c / _gplite_cf_matern32_c / _gplite_cf_matern52_c (5 args) and
 *   _gplite_cf_nn_c (6 args)          src/RcppExports.cpp:16-74, registered :76-87,
 *                                     called from R/RcppExports.R:4-18
 *   gplb_fit_full / gplb_fit_fitc     fit_laplace / fit_ep .method_full / .method_fitc
 *                                     (R/gp_fit.R:91-123,150-188)
 *   gplb_energy_grad                  gp_energy through gp_fit (R/gp.R:138-147) + gradient
 *   gplb_pred_full / gplb_pred_fitc   gp_pred_post.method_full / .method_fitc (R/gp_pred.R:202-296)
 *   gplb_fit_get                      gp$fit$C_chol, $Kxz ... materialised on demand
 *   gplb_set_device, gplb_dist_unique_id, gplb_dist_init, gplb_dist_fit_full,
 *   gplb_dist_energy_grad, gplb_dist_free
 *                                     the same fits for problems larger than one GPU: one R worker
 *                                     process per GPU (parallel::makeCluster), NCCL inside the library
 *                                     
 *
 * Build inside the package:  PKG_CPPFLAGS = -I$(GPLITE_B200_HOME)/include
 *                            PKG_LIBS     = -L$(GPLITE_B200_HOME) -lgplite_b200
 * R is not installed in the build image of this repository; tests/test_r_glue.py compiles this
 * file against a stand-in Rinternals.h (tests/r_shim/) and runs it.
 *
 * Error rule: every C resource is released before Rf_error (which longjmps).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include <stdlib.h>
#include <string.h>

#include "gplite_b200.h"

static gpl_ctx* ctx_ = NULL;
static int device_ = 0;      /* gplb_set_device(): the GPU of this R process (multi-GPU workers) */

static gpl_ctx* ctx(void) {
  if (!ctx_ && gpl_ctx_create(device_, 0, &ctx_) != GPL_OK) Rf_error("gplite_b200: %s", gpl_last_error());
  return ctx_;
}

static SEXP list_get(SEXP list, const char* name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (Rf_isNull(names)) return R_NilValue;
  for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}

static double get_real(SEXP list, const char* name, double dflt) {
  SEXP v = list_get(list, name);
  return Rf_isNull(v) ? dflt : Rf_asReal(v);
}

static int get_int(SEXP list, const char* name, int dflt) {
  SEXP v = list_get(list, name);
  return Rf_isNull(v) ? dflt : Rf_asInteger(v);
}

static const double* opt_real(SEXP v) { return Rf_isNull(v) ? NULL : REAL(v); }

/* ---- the reference's four .Call targets -------------------------------------------------- */

typedef int (*cf2_fn)(gpl_ctx*, const double*, int64_t, const double*, int64_t, int32_t, double, double,
                      int32_t, double*);

static SEXP cf_native2(cf2_fn fn, SEXP x1, SEXP x2, SEXP p0, SEXP p1, SEXP diag_only) {
  const int64_t n1 = Rf_nrows(x1), n2 = Rf_nrows(x2);
  const int d = Rf_ncols(x1);
  const int diag = Rf_asLogical(diag_only);
  SEXP out = PROTECT(diag ? Rf_allocMatrix(REALSXP, (int)(n1 < n2 ? n1 : n2), 1)
                          : Rf_allocMatrix(REALSXP, (int)n1, (int)n2));
  const int st = fn(ctx(), REAL(x1), n1, REAL(x2), n2, d, Rf_asReal(p0), Rf_asReal(p1), diag, REAL(out));
  UNPROTECT(1);
  if (st != GPL_OK) Rf_error("gplite_b200: %s", gpl_last_error());
  return out;
}

SEXP _gplite_cf_sexp_c(SEXP x1, SEXP x2, SEXP lscale, SEXP magn, SEXP diag_only) {
  return cf_native2(gpl_cf_sexp_c, x1, x2, lscale, magn, diag_only);
}
SEXP _gplite_cf_matern32_c(SEXP x1, SEXP x2, SEXP lscale, SEXP magn, SEXP diag_only) {
  return cf_native2(gpl_cf_matern32_c, x1, x2, lscale, magn, diag_only);
}
SEXP _gplite_cf_matern52_c(SEXP x1, SEXP x2, SEXP lscale, SEXP magn, SEXP diag_only) {
  return cf_native2(gpl_cf_matern52_c, x1, x2, lscale, magn, diag_only);
}
SEXP _gplite_cf_nn_c(SEXP x1, SEXP x2, SEXP sigma0, SEXP sigma, SEXP magn, SEXP diag_only) {
  const int64_t n1 = Rf_nrows(x1), n2 = Rf_nrows(x2);
  const int d = Rf_ncols(x1);
  const int diag = Rf_asLogical(diag_only);
  SEXP out = PROTECT(diag ? Rf_allocMatrix(REALSXP, (int)(n1 < n2 ? n1 : n2), 1)
                          : Rf_allocMatrix(REALSXP, (int)n1, (int)n2));
  const int st = gpl_cf_nn_c(ctx(), REAL(x1), n1, REAL(x2), n2, d, Rf_asReal(sigma0), Rf_asReal(sigma),
                             Rf_asReal(magn), diag, REAL(out));
  UNPROTECT(1);
  if (st != GPL_OK) Rf_error("gplite_b200: %s", gpl_last_error());
  return out;
}

/* ---- model descriptors from R lists ------------------------------------------------------
 * `model`: list of factors, each list(kind = <0..5>, term = <0-based>, periodic =, normalize =,
 * vars = <0-based integer vector or NULL>, means =, scales =, p = <numeric>, period =) built by
 * b200_model(gp$cfs) on the R side. */
typedef struct {
  gpl_cf_model m;
  gpl_cf_factor* f;
} model_buf;

static int unpack_model(SEXP model, model_buf* mb) {
  const int nf = (int)XLENGTH(model);
  mb->f = (gpl_cf_factor*)calloc((size_t)(nf > 0 ? nf : 1), sizeof(gpl_cf_factor));
  if (!mb->f) return 0;
  int nterms = 0;
  for (int i = 0; i < nf; ++i) {
    SEXP fi = VECTOR_ELT(model, i);
    gpl_cf_factor* f = &mb->f[i];
    f->kind = get_int(fi, "kind", 0);
    f->term = get_int(fi, "term", 0);
    f->periodic = get_int(fi, "periodic", 0);
    f->normalize = get_int(fi, "normalize", 0);
    SEXP vars = list_get(fi, "vars");
    f->nvars = Rf_isNull(vars) ? 0 : (int)XLENGTH(vars);
    f->vars = Rf_isNull(vars) ? NULL : (const int32_t*)INTEGER(vars);
    f->means = opt_real(list_get(fi, "means"));
    f->scales = opt_real(list_get(fi, "scales"));
    SEXP p = list_get(fi, "p");
    for (int k = 0; k < 3; ++k) f->p[k] = (!Rf_isNull(p) && k < XLENGTH(p)) ? REAL(p)[k] : 0.0;
    f->period = get_real(fi, "period", 1.0);
    if (f->term + 1 > nterms) nterms = f->term + 1;
  }
  mb->m.nfactors = nf;
  mb->m.nterms = nterms;
  mb->m.factors = mb->f;
  return 1;
}

static void unpack_lik(SEXP lik, gpl_lik* lk) {
  lk->kind = get_int(lik, "kind", GPL_LIK_GAUSSIAN);
  lk->link = get_int(lik, "link", GPL_LINK_IDENTITY);
  lk->param = get_real(lik, "param", 0.0);
}

static void unpack_approx(SEXP approx, gpl_approx* ap) {
  memset(ap, 0, sizeof(*ap));
  ap->kind = get_int(approx, "kind", GPL_APPROX_LAPLACE);
  ap->maxiter = get_int(approx, "maxiter", 30);
  ap->tol = get_real(approx, "tol", 1e-4);
  ap->damping = get_real(approx, "damping", 0.9);
  ap->damping_min = get_real(approx, "damping_min", 0.1);
  SEXP qx = list_get(approx, "quad_x"), qw = list_get(approx, "quad_w");
  ap->quad_order = Rf_isNull(qx) ? 0 : (int)XLENGTH(qx);
  ap->quad_x = opt_real(qx);
  ap->quad_w = opt_real(qw);
}

static uint64_t term_mask(SEXP cfind) {   /* 0-based indices; NULL = all terms */
  uint64_t mask = 0;
  if (Rf_isNull(cfind)) return 0;
  for (R_xlen_t i = 0; i < XLENGTH(cfind); ++i) mask |= (uint64_t)1 << INTEGER(cfind)[i];
  return mask;
}

/* ---- fits ---------------------------------------------------------------------------------- */

static void fit_finalizer(SEXP p) {
  gpl_fit* f = (gpl_fit*)R_ExternalPtrAddr(p);
  if (f) {
    gpl_fit_free(ctx_, f);
    R_ClearExternalPtr(p);
  }
}

static SEXP named_list(int n, const char** names) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, n));
  for (int i = 0; i < n; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

/* shared tail of gplb_fit_full / gplb_fit_fitc: gp$fit as the reference's iterations return it
 * (R/laplace.R:119, R/ep.R:140-145), C_chol / Kxz behind `handle` */
static SEXP fit_result(int st, int fitc, const gpl_fit_out* o, gpl_fit* h, SEXP* vec /* 7 protected */) {
  static const char* names[] = {"fmean", "z", "V", "alpha", "fvar", "cavity_mean", "cavity_var",
                                "log_evidence", "iters", "converged", "quad_ok", "n_bad", "handle"};
  if (st == GPL_ERR_NONFINITE || (fitc && st == GPL_ERR_NOT_PD)) {
    /* R/ep.R:30-32, R/laplace.R:73-80; FITC swallows factorisation errors (R/gp_fit.R:113-121) */
    static const char* nm1[] = {"log_evidence"};
    SEXP out = PROTECT(named_list(1, nm1));
    SET_VECTOR_ELT(out, 0, Rf_ScalarReal(R_NegInf));
    UNPROTECT(1);
    return out;
  }
  if (st != GPL_OK) return R_NilValue;   /* caller raises */
  SEXP out = PROTECT(named_list(13, names));
  for (int i = 0; i < 7; ++i) SET_VECTOR_ELT(out, i, vec[i]);
  SET_VECTOR_ELT(out, 7, Rf_ScalarReal(o->log_evidence));
  SET_VECTOR_ELT(out, 8, Rf_ScalarInteger(o->iters));
  SET_VECTOR_ELT(out, 9, Rf_ScalarLogical(o->converged));
  SET_VECTOR_ELT(out, 10, Rf_ScalarLogical(o->quad_ok));
  SET_VECTOR_ELT(out, 11, Rf_ScalarInteger(o->n_bad));
  SEXP hp = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(hp, fit_finalizer, TRUE);
  SET_VECTOR_ELT(out, 12, hp);
  UNPROTECT(2);
  return out;
}

SEXP gplb_fit_full(SEXP model, SEXP lik, SEXP approx, SEXP x, SEXP y, SEXP trials, SEXP offset,
                   SEXP jitter) {
  model_buf mb;
  gpl_lik lk;
  gpl_approx ap;
  if (!unpack_model(model, &mb)) Rf_error("gplite_b200: out of memory");
  unpack_lik(lik, &lk);
  unpack_approx(approx, &ap);
  const int64_t n = Rf_nrows(x);
  const int d = Rf_ncols(x);
  SEXP vec[7];
  for (int i = 0; i < 7; ++i) vec[i] = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
  gpl_fit_out o;
  gpl_fit* h = NULL;
  const int st = gpl_fit_full(ctx(), &mb.m, &lk, &ap, REAL(x), n, d, REAL(y), opt_real(trials),
                              opt_real(offset), Rf_asReal(jitter), &o, REAL(vec[0]), REAL(vec[1]),
                              REAL(vec[2]), REAL(vec[3]), REAL(vec[4]), REAL(vec[5]), REAL(vec[6]), &h);
  free(mb.f);
  SEXP out = PROTECT(fit_result(st, 0, &o, h, vec));
  UNPROTECT(8);
  /* a non-PD B aborts for method_full, as R's chol() does (R/gp_fit.R:95-97) */
  if (Rf_isNull(out)) Rf_error("gplite_b200: %s", gpl_last_error());
  return out;
}

SEXP gplb_fit_fitc(SEXP model, SEXP lik, SEXP approx, SEXP x, SEXP z, SEXP y, SEXP trials, SEXP offset,
                   SEXP jitter) {
  model_buf mb;
  gpl_lik lk;
  gpl_approx ap;
  if (!unpack_model(model, &mb)) Rf_error("gplite_b200: out of memory");
  unpack_lik(lik, &lk);
  unpack_approx(approx, &ap);
  const int64_t n = Rf_nrows(x), m = Rf_nrows(z);
  const int d = Rf_ncols(x);
  SEXP vec[7];
  for (int i = 0; i < 7; ++i) vec[i] = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
  gpl_fit_out o;
  gpl_fit* h = NULL;
  const int st = gpl_fit_fitc(ctx(), &mb.m, &lk, &ap, REAL(x), n, d, REAL(z), m, REAL(y), opt_real(trials),
                              opt_real(offset), Rf_asReal(jitter), &o, REAL(vec[0]), REAL(vec[1]),
                              REAL(vec[2]), REAL(vec[3]), REAL(vec[4]), REAL(vec[5]), REAL(vec[6]), &h);
  free(mb.f);
  SEXP out = PROTECT(fit_result(st, 1, &o, h, vec));
  UNPROTECT(8);
  if (Rf_isNull(out)) Rf_error("gplite_b200: %s", gpl_last_error());
  return out;
}

/* list(energy =, grad =): -log_evidence without the priors, gradient w.r.t. log-parameters
 * identified by integer vectors (factor, slot); factor -1 = likelihood sigma */
SEXP gplb_energy_grad(SEXP model, SEXP lik, SEXP x, SEXP y, SEXP jitter, SEXP pfactor, SEXP pslot) {
  model_buf mb;
  gpl_lik lk;
  if (!unpack_model(model, &mb)) Rf_error("gplite_b200: out of memory");
  unpack_lik(lik, &lk);
  const int np = (int)XLENGTH(pfactor);
  SEXP grad = PROTECT(Rf_allocVector(REALSXP, np));
  double energy = 0.0;
  const int st = gpl_energy_grad(ctx(), &mb.m, &lk, REAL(x), Rf_nrows(x), Rf_ncols(x), REAL(y), 0,
                                 Rf_asReal(jitter), np, (const int32_t*)INTEGER(pfactor),
                                 (const int32_t*)INTEGER(pslot), &energy, REAL(grad), NULL);
  free(mb.f);
  if (st != GPL_OK) {
    UNPROTECT(1);
    Rf_error("gplite_b200: %s", gpl_last_error());
  }
  static const char* names[] = {"energy", "grad"};
  SEXP out = PROTECT(named_list(2, names));
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal(energy));
  SET_VECTOR_ELT(out, 1, grad);
  UNPROTECT(2);
  return out;
}

typedef int (*pred_fn)(gpl_ctx*, const gpl_fit*, const gpl_cf_model*, const double*, int64_t, uint64_t,
                       double*, double*);

static SEXP pred_common(pred_fn fn, SEXP handle, SEXP model, SEXP xt, SEXP cfind, SEXP var) {
  model_buf mb;
  if (!unpack_model(model, &mb)) Rf_error("gplite_b200: out of memory");
  const int64_t nt = Rf_nrows(xt);
  const int want_var = Rf_asLogical(var);
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nt));
  SEXP v = PROTECT(want_var ? Rf_allocVector(REALSXP, (R_xlen_t)nt) : R_NilValue);
  const int st = fn(ctx(), (const gpl_fit*)R_ExternalPtrAddr(handle), &mb.m, REAL(xt), nt, term_mask(cfind),
                    REAL(mean), want_var ? REAL(v) : NULL);
  free(mb.f);
  if (st != GPL_OK) {
    UNPROTECT(2);
    Rf_error("gplite_b200: %s", gpl_last_error());
  }
  static const char* names[] = {"mean", "var"};
  SEXP out = PROTECT(named_list(want_var ? 2 : 1, names));
  SET_VECTOR_ELT(out, 0, mean);
  if (want_var) SET_VECTOR_ELT(out, 1, v);
  UNPROTECT(3);
  return out;
}

SEXP gplb_pred_full(SEXP handle, SEXP model, SEXP xt, SEXP cfind, SEXP var) {
  return pred_common(gpl_pred_full, handle, model, xt, cfind, var);
}
SEXP gplb_pred_fitc(SEXP handle, SEXP model, SEXP xt, SEXP cfind, SEXP var) {
  return pred_common(gpl_pred_fitc, handle, model, xt, cfind, var);
}

/* gp$fit$C_chol etc. on demand: what = "C_chol" | "Kxz" | "Kz_chol" | "alpha" | ...; rows, cols
 * as the caller knows them */
SEXP gplb_fit_get(SEXP handle, SEXP what, SEXP rows, SEXP cols) {
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(rows), Rf_asInteger(cols)));
  const int st = gpl_fit_get(ctx(), (const gpl_fit*)R_ExternalPtrAddr(handle), CHAR(STRING_ELT(what, 0)),
                             REAL(out));
  UNPROTECT(1);
  if (st != GPL_OK) Rf_error("gplite_b200: %s", gpl_last_error());
  return out;
}

/* ---- multi-GPU (one R process per GPU; include/gplite_b200.h "multi-GPU full GP") ------------- */

/* choose this process's GPU; must precede the first library call of the process */
SEXP gplb_set_device(SEXP dev) {
  if (ctx_) Rf_error("gplite_b200: gplb_set_device must be called before the first fit of this process");
  device_ = Rf_asInteger(dev);
  return R_NilValue;
}

/* rank 0: 128 raw bytes to ship to the other workers (clusterApply argument) */
SEXP gplb_dist_unique_id(void) {
  SEXP id = PROTECT(Rf_allocVector(RAWSXP, GPL_DIST_ID_BYTES));
  const int st = gpl_dist_unique_id(RAW(id));
  UNPROTECT(1);
  if (st != GPL_OK) Rf_error("gplite_b200: %s", gpl_last_error());
  return id;
}

static void dist_finalizer(SEXP p) {
  gpl_dist* d = (gpl_dist*)R_ExternalPtrAddr(p);
  if (d) {
    gpl_dist_destroy(d);
    R_ClearExternalPtr(p);
  }
}

/* collective over the workers: id (raw, may be NULL when world == 1), rank (0-based), world */
SEXP gplb_dist_init(SEXP id, SEXP rank, SEXP world) {
  gpl_dist* d = NULL;
  if (!Rf_isNull(id) && XLENGTH(id) != GPL_DIST_ID_BYTES) Rf_error("gplite_b200: the id must be 128 raw bytes");
  const int st = gpl_dist_init(ctx(), Rf_isNull(id) ? NULL : (const void*)RAW(id), Rf_asInteger(rank),
                               Rf_asInteger(world), &d);
  if (st != GPL_OK) Rf_error("gplite_b200: %s", gpl_last_error());
  SEXP p = PROTECT(R_MakeExternalPtr(d, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, dist_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}

/* explicit release (the destroy is a collective: call it on every worker) */
SEXP gplb_dist_free(SEXP dist) {
  dist_finalizer(dist);
  return R_NilValue;
}

/* fit_laplace / fit_ep .method_full across the workers' GPUs; every worker passes the same data and gets
 * the same gp$fit back (no device handle: predictions go through alpha, as gp_pred_post's mean does) */
SEXP gplb_dist_fit_full(SEXP dist, SEXP model, SEXP lik, SEXP approx, SEXP x, SEXP y, SEXP trials, SEXP offset,
                        SEXP jitter, SEXP nb) {
  model_buf mb;
  gpl_lik lk;
  gpl_approx ap;
  gpl_dist* d = (gpl_dist*)R_ExternalPtrAddr(dist);
  if (!d) Rf_error("gplite_b200: the multi-GPU communicator has been released");
  if (!unpack_model(model, &mb)) Rf_error("gplite_b200: out of memory");
  unpack_lik(lik, &lk);
  unpack_approx(approx, &ap);
  const int64_t n = Rf_nrows(x);
  const int dd = Rf_ncols(x);
  SEXP vec[7];
  for (int i = 0; i < 7; ++i) vec[i] = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
  gpl_fit_out o;
  const int st = gpl_dist_fit_full(d, &mb.m, &lk, &ap, REAL(x), n, dd, REAL(y), opt_real(trials), opt_real(offset),
                                   Rf_asReal(jitter), (int64_t)Rf_asInteger(nb), &o, REAL(vec[0]), REAL(vec[1]),
                                   REAL(vec[2]), REAL(vec[3]), REAL(vec[4]), REAL(vec[5]), REAL(vec[6]), NULL);
  free(mb.f);
  static const char* names[] = {"fmean", "z", "V", "alpha", "fvar", "cavity_mean", "cavity_var",
                                "log_evidence", "iters", "converged", "quad_ok"};
  if (st == GPL_ERR_NONFINITE) {          /* R/ep.R:30-32 */
    static const char* nm1[] = {"log_evidence"};
    SEXP out = PROTECT(named_list(1, nm1));
    SET_VECTOR_ELT(out, 0, Rf_ScalarReal(R_NegInf));
    UNPROTECT(8);
    return out;
  }
  if (st != GPL_OK) {
    UNPROTECT(7);
    Rf_error("gplite_b200: %s", gpl_last_error());
  }
  SEXP out = PROTECT(named_list(11, names));
  for (int i = 0; i < 7; ++i) SET_VECTOR_ELT(out, i, vec[i]);
  SET_VECTOR_ELT(out, 7, Rf_ScalarReal(o.log_evidence));
  SET_VECTOR_ELT(out, 8, Rf_ScalarInteger(o.iters));
  SET_VECTOR_ELT(out, 9, Rf_ScalarLogical(o.converged));
  SET_VECTOR_ELT(out, 10, Rf_ScalarLogical(o.quad_ok));
  UNPROTECT(8);
  return out;
}

/* gplb_energy_grad across the workers' GPUs (lik_gaussian): list(energy =, grad =) */
SEXP gplb_dist_energy_grad(SEXP dist, SEXP model, SEXP lik, SEXP x, SEXP y, SEXP jitter, SEXP nb, SEXP pfactor,
                           SEXP pslot) {
  model_buf mb;
  gpl_lik lk;
  gpl_dist* d = (gpl_dist*)R_ExternalPtrAddr(dist);
  if (!d) Rf_error("gplite_b200: the multi-GPU communicator has been released");
  if (!unpack_model(model, &mb)) Rf_error("gplite_b200: out of memory");
  unpack_lik(lik, &lk);
  const int np = (int)XLENGTH(pfactor);
  SEXP grad = PROTECT(Rf_allocVector(REALSXP, np));
  double energy = 0.0;
  const int st = gpl_dist_energy_grad(d, &mb.m, &lk, REAL(x), Rf_nrows(x), Rf_ncols(x), REAL(y), Rf_asReal(jitter),
                                      (int64_t)Rf_asInteger(nb), np, (const int32_t*)INTEGER(pfactor),
                                      (const int32_t*)INTEGER(pslot), &energy, REAL(grad), NULL, NULL);
  free(mb.f);
  if (st != GPL_OK) {
    UNPROTECT(1);
    Rf_error("gplite_b200: %s", gpl_last_error());
  }
  static const char* names[] = {"energy", "grad"};
  SEXP out = PROTECT(named_list(2, names));
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal(energy));
  SET_VECTOR_ELT(out, 1, grad);
  UNPROTECT(2);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"_gplite_cf_sexp_c", (DL_FUNC)&_gplite_cf_sexp_c, 5},
    {"_gplite_cf_matern32_c", (DL_FUNC)&_gplite_cf_matern32_c, 5},
    {"_gplite_cf_matern52_c", (DL_FUNC)&_gplite_cf_matern52_c, 5},
    {"_gplite_cf_nn_c", (DL_FUNC)&_gplite_cf_nn_c, 6},
    {"gplb_fit_full", (DL_FUNC)&gplb_fit_full, 8},
    {"gplb_fit_fitc", (DL_FUNC)&gplb_fit_fitc, 9},
    {"gplb_energy_grad", (DL_FUNC)&gplb_energy_grad, 7},
    {"gplb_pred_full", (DL_FUNC)&gplb_pred_full, 5},
    {"gplb_pred_fitc", (DL_FUNC)&gplb_pred_fitc, 5},
    {"gplb_fit_get", (DL_FUNC)&gplb_fit_get, 4},
    {"gplb_set_device", (DL_FUNC)&gplb_set_device, 1},
    {"gplb_dist_unique_id", (DL_FUNC)&gplb_dist_unique_id, 0},
    {"gplb_dist_init", (DL_FUNC)&gplb_dist_init, 3},
    {"gplb_dist_free", (DL_FUNC)&gplb_dist_free, 1},
    {"gplb_dist_fit_full", (DL_FUNC)&gplb_dist_fit_full, 10},
    {"gplb_dist_energy_grad", (DL_FUNC)&gplb_dist_energy_grad, 9},
    {NULL, NULL, 0}};

void R_init_gplite(DllInfo* dll) { /* replaces src/RcppExports.cpp:84-87 */
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

/* TEST INFRASTRUCTURE ONLY: the runtime behind tests/r_shim/Rinternals.h plus a few shim_* helpers
 * the Python test uses to build arguments, inspect results and catch Rf_error.  Objects are never
 * freed (a test process is short-lived); PROTECT / UNPROTECT only keep a balance counter that the
 * test asserts returns to zero. */
#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct shim_sexprec {
  SEXPTYPE type;
  R_xlen_t len;
  int nrow, ncol;
  void* data;          /* double / int / SEXP* / char */
  SEXP names;
  R_CFinalizer_t fin;
};

static struct shim_sexprec nil_ = {NILSXP, 0, 0, 0, NULL, NULL, NULL};
static struct shim_sexprec names_sym_ = {SYMSXP, 0, 0, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_;
SEXP R_NamesSymbol = &names_sym_;
double R_NegInf = -INFINITY, R_PosInf = INFINITY;

static int protect_balance = 0;
static char last_error[1024] = "";
static jmp_buf* error_jmp = NULL;
static const R_CallMethodDef* routines = NULL;
static int dynamic_symbols = -1;

static SEXP new_obj(SEXPTYPE t, R_xlen_t n, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(*s));
  s->type = t;
  s->len = n;
  s->nrow = (int)n;
  s->ncol = 1;
  s->data = calloc((size_t)(n > 0 ? n : 1), elt);
  s->names = R_NilValue;
  if (t == VECSXP || t == STRSXP)
    for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}

SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) {
  switch (t) {
    case REALSXP: return new_obj(t, n, sizeof(double));
    case INTSXP: case LGLSXP: return new_obj(t, n, sizeof(int));
    case VECSXP: case STRSXP: return new_obj(t, n, sizeof(SEXP));
    case RAWSXP: return new_obj(t, n, 1);
    default: Rf_error("shim: allocVector type %d", (int)t);
  }
}
SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc) {
  SEXP s = Rf_allocVector(t, (R_xlen_t)nr * nc);
  s->nrow = nr;
  s->ncol = nc;
  return s;
}
double* REAL(SEXP s) { if (s->type != REALSXP) Rf_error("shim: REAL() on type %d", (int)s->type); return (double*)s->data; }
int* INTEGER(SEXP s) { if (s->type != INTSXP) Rf_error("shim: INTEGER() on type %d", (int)s->type); return (int*)s->data; }
Rbyte* RAW(SEXP s) { if (s->type != RAWSXP) Rf_error("shim: RAW() on type %d", (int)s->type); return (Rbyte*)s->data; }
int* LOGICAL(SEXP s) { if (s->type != LGLSXP) Rf_error("shim: LOGICAL() on type %d", (int)s->type); return (int*)s->data; }
int Rf_nrows(SEXP s) { return s->nrow; }
int Rf_ncols(SEXP s) { return s->ncol; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int Rf_length(SEXP s) { return (int)s->len; }
int Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
double Rf_asReal(SEXP s) {
  if (s->len < 1) return NAN;
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) return (double)((int*)s->data)[0];
  Rf_error("shim: asReal on type %d", (int)s->type);
}
int Rf_asInteger(SEXP s) { return (int)Rf_asReal(s); }
int Rf_asLogical(SEXP s) { return Rf_asReal(s) != 0.0; }
SEXP Rf_protect(SEXP s) { ++protect_balance; return s; }
void Rf_unprotect(int n) { protect_balance -= n; }
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error, sizeof(last_error), fmt, ap);
  va_end(ap);
  if (error_jmp) longjmp(*error_jmp, 1);
  fprintf(stderr, "Rf_error outside shim_call: %s\n", last_error);
  abort();
}
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { return ((SEXP*)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP*)s->data)[i] = v; return v; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { return ((SEXP*)s->data)[i]; }
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP*)s->data)[i] = v; }
const char* CHAR(SEXP s) { return (const char*)s->data; }
SEXP Rf_mkChar(const char* c) {
  SEXP s = new_obj(CHARSXP, (R_xlen_t)strlen(c) + 1, 1);
  strcpy((char*)s->data, c);
  return s;
}
SEXP Rf_mkString(const char* c) {
  SEXP s = Rf_allocVector(STRSXP, 1);
  SET_STRING_ELT(s, 0, Rf_mkChar(c));
  return s;
}
SEXP Rf_getAttrib(SEXP s, SEXP what) { return what == R_NamesSymbol ? s->names : R_NilValue; }
SEXP Rf_setAttrib(SEXP s, SEXP what, SEXP v) { if (what == R_NamesSymbol) s->names = v; return v; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v; return s; }
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag; (void)prot;
  SEXP s = new_obj(EXTPTRSXP, 0, 1);
  free(s->data);
  s->data = p;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->fin = f; }
int R_registerRoutines(DllInfo* d, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)d; (void)c; (void)f; (void)e;
  routines = call;
  return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* d, Rboolean v) { (void)d; dynamic_symbols = (int)v; return TRUE; }

/* ---- helpers for the Python side --------------------------------------------------------------- */
int shim_routine_count(void) { int n = 0; while (routines && routines[n].name) ++n; return n; }
const char* shim_routine_name(int i) { return routines[i].name; }
int shim_routine_nargs(int i) { return routines[i].numArgs; }
int shim_dynamic_symbols(void) { return dynamic_symbols; }
int shim_protect_balance(void) { return protect_balance; }
const char* shim_last_error(void) { return last_error; }
SEXP shim_nil(void) { return R_NilValue; }
SEXP shim_real_matrix(const double* v, int nr, int nc) {
  SEXP s = Rf_allocMatrix(REALSXP, nr, nc);
  memcpy(s->data, v, sizeof(double) * (size_t)nr * nc);
  return s;
}
SEXP shim_int_vector(const int* v, int n) {
  SEXP s = Rf_allocVector(INTSXP, n);
  memcpy(s->data, v, sizeof(int) * (size_t)n);
  return s;
}
SEXP shim_list(int n) { return Rf_allocVector(VECSXP, n); }
void shim_list_set(SEXP list, int i, const char* name, SEXP v) {
  if (Rf_isNull(list->names)) list->names = Rf_allocVector(STRSXP, list->len);
  SET_STRING_ELT(list->names, i, Rf_mkChar(name ? name : ""));
  SET_VECTOR_ELT(list, i, v);
}
SEXP shim_list_get(SEXP list, const char* name) {
  for (R_xlen_t i = 0; i < list->len; ++i)
    if (!Rf_isNull(list->names) && strcmp(CHAR(STRING_ELT(list->names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}
int shim_type(SEXP s) { return (int)s->type; }
void shim_run_finalizer(SEXP s) { if (s->fin) s->fin(s); }
/* call a registered .Call routine by name with 0 .. 10 SEXP arguments; returns NULL (C) when the
 * routine raised through Rf_error (message in shim_last_error) */
SEXP shim_call(const char* name, int nargs, SEXP* a) {
  for (int i = 0; routines && routines[i].name; ++i) {
    if (strcmp(routines[i].name, name)) continue;
    if (routines[i].numArgs != nargs) { snprintf(last_error, sizeof(last_error), "arity mismatch"); return NULL; }
    jmp_buf jb;
    error_jmp = &jb;
    SEXP out = NULL;
    if (setjmp(jb) == 0) {
      typedef SEXP (*f0)(void);
      typedef SEXP (*f1)(SEXP);
      typedef SEXP (*f3)(SEXP, SEXP, SEXP);
      typedef SEXP (*f10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
      typedef SEXP (*f5)(SEXP, SEXP, SEXP, SEXP, SEXP);
      typedef SEXP (*f4)(SEXP, SEXP, SEXP, SEXP);
      typedef SEXP (*f6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
      typedef SEXP (*f7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
      typedef SEXP (*f8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
      typedef SEXP (*f9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
      DL_FUNC f = routines[i].fun;
      switch (nargs) {
        case 0: out = ((f0)f)(); break;
        case 1: out = ((f1)f)(a[0]); break;
        case 3: out = ((f3)f)(a[0], a[1], a[2]); break;
        case 10: out = ((f10)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
        case 4: out = ((f4)f)(a[0], a[1], a[2], a[3]); break;
        case 5: out = ((f5)f)(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: out = ((f6)f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: out = ((f7)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        case 8: out = ((f8)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
        case 9: out = ((f9)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
        default: snprintf(last_error, sizeof(last_error), "unsupported arity %d", nargs);
      }
    } else {
      protect_balance = 0;   /* R unwinds the protect stack on error */
    }
    error_jmp = NULL;
    return out;
  }
  snprintf(last_error, sizeof(last_error), "no routine %s", name);
  return NULL;
}

/* TEST INFRASTRUCTURE ONLY -- a stand-in for R's Rinternals.h (R is not installed in this image),
 * just large enough to compile and RUN r_glue/gplite_b200_glue.c from tests/test_r_glue.py: a tiny
 * SEXP implementation (numeric / integer / logical / character / list vectors with a names
 * attribute, external pointers, a routine registry).  Same trick as oracle/shim/RcppArmadillo.h.
 * Nothing here is R code; the names and signatures follow the public R C API so that the glue
 * compiles unchanged against the real headers. */
#ifndef GPLB_RSHIM_RINTERNALS_H
#define GPLB_RSHIM_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef ptrdiff_t R_xlen_t;
typedef struct shim_sexprec* SEXP;
typedef enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, RAWSXP = 24,
               CHARSXP = 9, SYMSXP = 1 } SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;
extern SEXP R_NilValue, R_NamesSymbol;
extern double R_NegInf, R_PosInf;
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
typedef unsigned char Rbyte;
Rbyte* RAW(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
R_xlen_t XLENGTH(SEXP);
int Rf_length(SEXP);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
int Rf_isNull(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...) __attribute__((noreturn));
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
const char* CHAR(SEXP);
SEXP Rf_mkChar(const char*);
SEXP Rf_mkString(const char*);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarLogical(int);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#ifdef __cplusplus
}
#endif
#endif

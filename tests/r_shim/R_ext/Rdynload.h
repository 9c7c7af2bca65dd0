/* TEST INFRASTRUCTURE ONLY: stand-in for R_ext/Rdynload.h (see ../Rinternals.h). */
#ifndef GPLB_RSHIM_RDYNLOAD_H
#define GPLB_RSHIM_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct shim_dllinfo DllInfo;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
Rboolean R_useDynamicSymbols(DllInfo*, Rboolean);
#ifdef __cplusplus
}
#endif
#endif

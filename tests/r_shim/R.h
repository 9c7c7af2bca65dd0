/* TEST INFRASTRUCTURE ONLY: stand-in for R.h (see Rinternals.h in this directory). */
#ifndef GPLB_RSHIM_R_H
#define GPLB_RSHIM_R_H
#include <stdint.h>
#endif

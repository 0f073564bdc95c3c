/* Stand-in for <gsl/gsl_sys.h> (test infrastructure). */
#ifndef MOPE_SHIM_GSL_SYS_H
#define MOPE_SHIM_GSL_SYS_H
#include <math.h>
static inline double gsl_log1p(const double x) { return log1p(x); }
#endif

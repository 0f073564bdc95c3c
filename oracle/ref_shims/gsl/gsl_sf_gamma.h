/* Stand-in for <gsl/gsl_sf_gamma.h>: only what mope/_binom.pyx binds.
 * Test infrastructure (lets the unmodified reference Cython compile without libgsl).
 * Formula follows GSL's published gsl_sf_lnchoose: lnfact(n) - lnfact(m) - lnfact(n-m)
 * with m := min(m, n-m) and the m==0 / m==n short-cut. */
#ifndef MOPE_SHIM_GSL_SF_GAMMA_H
#define MOPE_SHIM_GSL_SF_GAMMA_H
#include <math.h>
static inline double gsl_sf_lnchoose(unsigned int n, unsigned int m) {
    if (m > n) return NAN;
    if (m == n || m == 0) return 0.0;
    if (m * 2 > n) m = n - m;
    return lgamma((double)n + 1.0) - lgamma((double)m + 1.0) - lgamma((double)(n - m) + 1.0);
}
#endif

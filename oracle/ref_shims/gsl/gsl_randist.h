/* Stand-in for <gsl/gsl_randist.h> (test infrastructure): GSL's published
 * gsl_ran_binomial_pdf algorithm, plus the sampler symbol _wf.pyx links. */
#ifndef MOPE_SHIM_GSL_RANDIST_H
#define MOPE_SHIM_GSL_RANDIST_H
#include <math.h>
#include <stdlib.h>
#include "gsl_sf_gamma.h"
#include "gsl_rng.h"
static inline double gsl_ran_binomial_pdf(const unsigned int k, const double p, const unsigned int n) {
    if (k > n) return 0.0;
    if (p == 0.0) return (k == 0) ? 1.0 : 0.0;
    if (p == 1.0) return (k == n) ? 1.0 : 0.0;
    {
        double ln_Cnk = gsl_sf_lnchoose(n, k);
        double P = ln_Cnk + k * log(p) + (n - k) * log1p(-p);
        return exp(P);
    }
}
/* simulation only; never on the likelihood path */
static inline unsigned int gsl_ran_binomial(const gsl_rng* r, double p, unsigned int n) {
    unsigned int i, c = 0; (void)r;
    for (i = 0; i < n; ++i) c += (drand48() < p);
    return c;
}
#endif

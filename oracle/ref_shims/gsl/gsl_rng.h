/* Stand-in for <gsl/gsl_rng.h> (test infrastructure; simulators only). */
#ifndef MOPE_SHIM_GSL_RNG_H
#define MOPE_SHIM_GSL_RNG_H
#include <stdlib.h>
typedef struct { int dummy; } gsl_rng_type;
typedef struct { unsigned long seed; } gsl_rng;
static gsl_rng_type mope_shim_mt19937 = {0};
static gsl_rng_type* gsl_rng_mt19937 = &mope_shim_mt19937;
static inline gsl_rng* gsl_rng_alloc(const gsl_rng_type* T) { (void)T; return (gsl_rng*)calloc(1, sizeof(gsl_rng)); }
static inline void gsl_rng_set(gsl_rng* r, unsigned long s) { r->seed = s; srand48((long)s); }
static inline void gsl_rng_free(gsl_rng* r) { free(r); }
#endif

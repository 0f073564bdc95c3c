"""Stand-in for emcee (TEST INFRASTRUCTURE: emcee is neither vendored in the reference nor installed here).

`EnsembleSampler` has the constructor and `sample` signature the reference's Inference.run_mcmc uses
(mope/inference.py:1350-1360): EnsembleSampler(nwalkers, dim, lnpostfn, a=2.0, pool=None) iterated with
sample(p0, iterations=..., store=False).  Like emcee it never hands the bare function to the pool: every batch goes
through `pool.map(wrapper, positions)` (or the builtin map) with a wrapper OBJECT that has .f / .args / .kwargs and no
__name__ -- exactly what emcee.ensemble._FunctionWrapper looks like to a pool.  The stretch move itself is the repo's
own restatement (mope_b200.ensemble); set `EnsembleSampler.default_seed` for reproducible chains.
"""
import numpy as np


class _FunctionWrapper(object):
    def __init__(self, f, args=None, kwargs=None):
        self.f = f
        self.args = [] if args is None else args
        self.kwargs = {} if kwargs is None else kwargs

    def __call__(self, x):
        return self.f(x, *self.args, **self.kwargs)


class EnsembleSampler(object):
    default_seed = None

    def __init__(self, nwalkers, dim, lnpostfn, a=2.0, args=None, kwargs=None, pool=None, **_ignored):
        from mope_b200.ensemble import EnsembleSampler as _Stretch
        self.pool = pool
        self.lnprobfn = _FunctionWrapper(lnpostfn, args, kwargs)
        self.map_calls = 0
        self._s = _Stretch(nwalkers, dim, self._batch, a=a, seed=self.default_seed)

    def _batch(self, P):
        M = map if self.pool is None else self.pool.map
        self.map_calls += 1
        return np.array(list(M(self.lnprobfn, [P[i] for i in range(P.shape[0])])), dtype=np.float64)

    @property
    def acceptance_fraction(self):
        return self._s.acceptance_fraction

    def sample(self, p0, lnprob0=None, iterations=1, store=False):
        return self._s.sample(p0, lnprob0=lnprob0, iterations=iterations)


class PTSampler(EnsembleSampler):
    pass

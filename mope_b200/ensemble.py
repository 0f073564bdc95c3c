"""Affine-invariant ensemble sampler (Goodman & Weare stretch move) -- the part of emcee that mope's
`Inference.run_mcmc` relies on (mope/inference.py:1350-1363: `emcee.EnsembleSampler(num_walkers, ndim, post_clone,
pool=pool, a=chain_alpha)` iterated with `sample(init_pos, iterations, store=False)`).

emcee is neither vendored in the reference nor installed here; this is a restatement of its published stretch move:
the ensemble is split into two halves, every walker of a half proposes  y = c + z (x - c)  with a partner c drawn from
the other half and z ~ g(z) ∝ 1/sqrt(z) on [1/a, a], accepted with probability min(1, z^(ndim-1) p(y)/p(x)).
Each half-step is ONE batched posterior evaluation (what emcee does with `pool.map`), i.e. one device call.
"""
from __future__ import annotations

from typing import Callable, Iterator, Optional, Tuple

import numpy as np


class EnsembleSampler(object):
    def __init__(self, nwalkers: int, ndim: int, log_prob_batch: Callable[[np.ndarray], np.ndarray], a: float = 2.0,
                 seed: Optional[int] = None, randomize_split: bool = True):
        if nwalkers < 2 * ndim or nwalkers % 2:
            raise ValueError("the stretch move needs an even number of walkers, at least 2*ndim (got {} for ndim {})".format(
                nwalkers, ndim))
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_batch = log_prob_batch
        self.random = np.random.RandomState(seed)
        self.randomize_split = randomize_split
        self.naccepted = np.zeros(self.nwalkers)
        self.iterations = 0

    @property
    def acceptance_fraction(self) -> np.ndarray:
        return self.naccepted / max(self.iterations, 1)

    def sample(self, p0: np.ndarray, lnprob0: Optional[np.ndarray] = None, iterations: int = 1
               ) -> Iterator[Tuple[np.ndarray, np.ndarray, tuple]]:
        p = np.array(p0, dtype=np.float64, copy=True)
        if p.shape != (self.nwalkers, self.ndim):
            raise ValueError("initial positions must have shape (nwalkers, ndim)")
        lnprob = np.asarray(self.log_prob_batch(p), dtype=np.float64) if lnprob0 is None else np.array(lnprob0, dtype=np.float64)
        if np.any(np.isnan(lnprob)):
            raise ValueError("The initial lnprob was NaN.")
        half = self.nwalkers // 2
        for _ in range(int(iterations)):
            self.iterations += 1
            idx = self.random.permutation(self.nwalkers) if self.randomize_split else np.arange(self.nwalkers)
            for S, C in ((idx[:half], idx[half:]), (idx[half:], idx[:half])):
                s, c = p[S], p[C]
                zz = ((self.a - 1.0) * self.random.rand(len(S)) + 1) ** 2.0 / self.a
                factors = (self.ndim - 1.0) * np.log(zz)
                rint = self.random.randint(len(C), size=(len(S),))
                q = c[rint] - (c[rint] - s) * zz[:, None]
                newlp = np.asarray(self.log_prob_batch(q), dtype=np.float64)
                newlp = np.where(np.isnan(newlp), -np.inf, newlp)
                lnpdiff = factors + newlp - lnprob[S]
                accepted = np.log(self.random.rand(len(S))) < lnpdiff
                p[S[accepted]] = q[accepted]
                lnprob[S[accepted]] = newlp[accepted]
                self.naccepted[S[accepted]] += 1
            yield p.copy(), lnprob.copy(), self.random.get_state()

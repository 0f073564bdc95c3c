"""Root (stationary) allele-frequency distribution -- host side, the reference's own SciPy call.

mope/likelihoods.py:27-60: point masses (1-ppoly)/2 at frequency 0 and 1 and ppoly * Beta(ab, ab) in between,
discretised on the N-1 interior Wright-Fisher cells by CDF differences, then summed into the frequency bins with
np.add.reduceat(., breaks).  The CDF differences cancel catastrophically for small ab (their relative error is
~1e-16/ab), so bit-parity with the reference requires the same CDF implementation: scipy.special.betainc is what
scipy.stats.beta.cdf evaluates (checked bit-identical in tests/test_host_logic.py).  The cell edges are shared
between neighbouring cells, so each distinct edge is evaluated once and the whole batch of walkers in one call.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from scipy.special import betainc

_MAX_THREADS = max(1, min(16, (os.cpu_count() or 1)))
_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=_MAX_THREADS)
    return _POOL


def stationary_distribution_batch(breaks: np.ndarray, N: int, ab: np.ndarray, prob_poly: np.ndarray) -> np.ndarray:
    """[W, F] array; row w equals likelihoods.get_stationary_distribution_double_beta(freqs, breaks, N, ab[w], ppoly[w])."""
    ab = np.asarray(ab, dtype=np.float64).reshape(-1)
    pp = np.asarray(prob_poly, dtype=np.float64).reshape(-1)
    N = int(N)
    Ni = N - 2  # discretize_double_beta_distn: newN = N - 2
    # lower = [0, 1, 3, ..., 2Ni-1] / 2Ni ; upper = [1, 3, ..., 2Ni-1, 2Ni] / 2Ni
    edges = np.arange(1, 2 * Ni, 2) / (2 * Ni)
    W = ab.shape[0]
    cdf = np.empty((W, edges.shape[0]))
    nthreads = min(_MAX_THREADS, max(1, W // 8))
    if nthreads <= 1:
        betainc(ab[:, None], ab[:, None], edges[None, :], out=cdf)
    else:
        # the ufunc loop releases the GIL: rows are split over a few threads (the values do not depend on the split)
        bounds = np.linspace(0, W, nthreads + 1).astype(int)

        def work(i):
            lo, hi = bounds[i], bounds[i + 1]
            if hi > lo:
                betainc(ab[lo:hi, None], ab[lo:hi, None], edges[None, :], out=cdf[lo:hi])

        list(_pool().map(work, range(nthreads)))
    cdf_lo = np.concatenate((np.zeros((W, 1)), cdf), axis=1)   # beta.cdf(0) = 0
    cdf_hi = np.concatenate((cdf, np.ones((W, 1))), axis=1)    # beta.cdf(1) = 1
    distn = np.zeros((W, N + 1))
    distn[:, 0] = (1.0 - pp) / 2
    distn[:, -1] = (1.0 - pp) / 2
    distn[:, 1:-1] = (cdf_hi - cdf_lo) * pp[:, None]
    return np.add.reduceat(distn, np.asarray(breaks), axis=1)

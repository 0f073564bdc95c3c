"""Root (stationary) allele-frequency distribution -- host side, the reference's own SciPy call.

mope/likelihoods.py:27-60: point masses (1-ppoly)/2 at frequency 0 and 1 and ppoly * Beta(ab, ab) in between,
discretised on the N-1 interior Wright-Fisher cells by CDF differences, then summed into the frequency bins with
np.add.reduceat(., breaks).  The CDF differences cancel catastrophically for small ab (their relative error is
~1e-16/ab), so bit-parity with the reference requires the same CDF implementation: scipy.special.betainc is what
scipy.stats.beta.cdf evaluates (checked bit-identical in tests/test_host_logic.py).  The cell edges are shared
between neighbouring cells, so each distinct edge is evaluated once and the whole batch of walkers in one call.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
from scipy.special import betainc, cython_special

_MAX_THREADS = int(os.environ.get("MOPE_B200_HOST_THREADS", "0")) or max(1, min(16, (os.cpu_count() or 1)))


def _betainc_entry():
    """Address of SciPy's own double-precision betainc (the function behind the ufunc and scipy.stats.beta.cdf)."""
    cap = cython_special.__pyx_capi__["__pyx_fuse_0betainc"]
    api = ctypes.pythonapi
    api.PyCapsule_GetName.restype = ctypes.c_char_p
    api.PyCapsule_GetName.argtypes = [ctypes.py_object]
    api.PyCapsule_GetPointer.restype = ctypes.c_void_p
    api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    name = api.PyCapsule_GetName(cap)
    if name is None or b"double (double, double, double" not in name:
        raise RuntimeError("unexpected signature of scipy.special.cython_special betainc: %r" % (name,))
    return api.PyCapsule_GetPointer(cap, name)


_ENTRY = None


def beta_cdf_table(ab: np.ndarray, edges: np.ndarray, n_threads: int = _MAX_THREADS) -> np.ndarray:
    """cdf[w, j] = scipy.stats.beta.cdf(edges[j], ab[w], ab[w]), bit for bit, evaluated by native threads of the C-ABI
    library through SciPy's own C entry point (mope_b200_beta_cdf_table): no interpreter lock, no Python thread pool."""
    global _ENTRY
    from . import _lib
    if _ENTRY is None:
        _ENTRY = _betainc_entry()
    ab = np.ascontiguousarray(ab, dtype=np.float64)
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    out = np.empty((ab.shape[0], edges.shape[0]))
    _lib.check(_lib.lib().mope_b200_beta_cdf_table(_ENTRY, _lib.dptr(ab), ab.shape[0], _lib.dptr(edges), edges.shape[0],
                                                   _lib.dptr(out), int(n_threads)), "mope_b200_beta_cdf_table")
    return out


def stationary_distribution_batch(breaks: np.ndarray, N: int, ab: np.ndarray, prob_poly: np.ndarray) -> np.ndarray:
    """[W, F] array; row w equals likelihoods.get_stationary_distribution_double_beta(freqs, breaks, N, ab[w], ppoly[w])."""
    ab = np.asarray(ab, dtype=np.float64).reshape(-1)
    pp = np.asarray(prob_poly, dtype=np.float64).reshape(-1)
    N = int(N)
    Ni = N - 2  # discretize_double_beta_distn: newN = N - 2
    # lower = [0, 1, 3, ..., 2Ni-1] / 2Ni ; upper = [1, 3, ..., 2Ni-1, 2Ni] / 2Ni
    edges = np.arange(1, 2 * Ni, 2) / (2 * Ni)
    W = ab.shape[0]
    cdf = beta_cdf_table(ab, edges)
    cdf_lo = np.concatenate((np.zeros((W, 1)), cdf), axis=1)   # beta.cdf(0) = 0
    cdf_hi = np.concatenate((cdf, np.ones((W, 1))), axis=1)    # beta.cdf(1) = 1
    distn = np.zeros((W, N + 1))
    distn[:, 0] = (1.0 - pp) / 2
    distn[:, -1] = (1.0 - pp) / 2
    distn[:, 1:-1] = (cdf_hi - cdf_lo) * pp[:, None]
    return np.add.reduceat(distn, np.asarray(breaks), axis=1)

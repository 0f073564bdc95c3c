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



def _default_threads() -> int:
    """Native threads of the root-prior table: the host cores divided by the ranks sharing the box (torchrun exports
    LOCAL_WORLD_SIZE), at most 16 -- eight ranks spinning 16 threads each on a 32-core host slowed every rank down."""
    env = int(os.environ.get("MOPE_B200_HOST_THREADS", "0"))
    if env > 0:
        return env
    cores = os.cpu_count() or 1
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, min(16, cores // local_world))


_MAX_THREADS = _default_threads()


def _betainc_entry():
    """Address of SciPy's own double-precision betainc (the function behind the ufunc and scipy.stats.beta.cdf).
    The capsule name is private to SciPy (checked with 1.18); None when it is missing or has another signature --
    beta_cdf_table then evaluates the ufunc itself (same values bit for bit, single-threaded)."""
    cap = getattr(cython_special, "__pyx_capi__", {}).get("__pyx_fuse_0betainc")
    if cap is None:
        return None
    api = ctypes.pythonapi
    api.PyCapsule_GetName.restype = ctypes.c_char_p
    api.PyCapsule_GetName.argtypes = [ctypes.py_object]
    api.PyCapsule_GetPointer.restype = ctypes.c_void_p
    api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    name = api.PyCapsule_GetName(cap)
    if name is None or b"double (double, double, double" not in name:
        return None
    return api.PyCapsule_GetPointer(cap, name)


_ENTRY = False   # not looked up yet


def beta_cdf_table(ab: np.ndarray, edges: np.ndarray, n_threads: int = _MAX_THREADS) -> np.ndarray:
    """cdf[w, j] = scipy.stats.beta.cdf(edges[j], ab[w], ab[w]), bit for bit, evaluated by native threads of the C-ABI
    library through SciPy's own C entry point (mope_b200_beta_cdf_table): no interpreter lock, no Python thread pool."""
    global _ENTRY
    from . import _lib
    if _ENTRY is False:
        _ENTRY = _betainc_entry()
    ab = np.ascontiguousarray(ab, dtype=np.float64)
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    if _ENTRY is None:
        return betainc(ab[:, None], ab[:, None], edges[None, :])
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

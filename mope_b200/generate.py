"""Transition-table generator (offline side of the path).

Produces the same tables as the reference's `mope generate-transitions` + `make-master-transitions`:
Wright-Fisher matrix (mope/generate_transitions.py:57-67), matrix-power chain over the generation list
(:217-228, :386-414), bottleneck products through the doubling schedule (:29-54, mope/_transition.pyx:6-33),
symmetric frequency binning (:133-214).  With a GPU everything heavy stays on the device, in the library's own kernels
(csrc/generate.cuh through the C ABI `mope_b200_gen_*`): the binomial rows of the Wright-Fisher and bottleneck factors
(log-pmf from host libm), the 1001^3 fp64 matrix products (`dgemm_tn_kernel`: TMA ring + mma.sync.m8n8k4.f64, the tensor
pipe of the pruning kernels) and the binning (`bin_matrix_kernel`); torch only owns the buffers and only the F x F tables
come back.  `device="cpu"` runs the same steps in NumPy / SciPy (the build container has no GPU; the CPU tests use it).  The two paths agree to ~1e-11 relative (summation order, lgamma vs SciPy's pmf);
the tables are an input shared by both implementations, so this does not enter the parity of the likelihood.
Powers of the one-generation matrix are cached by repeated squaring, so a step of d generations costs
popcount(d)-1 products instead of a fresh matrix_power.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import scipy.stats as st
from scipy.special import gammaln

from .tables import TransitionTables

DEFAULT_GENS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 45, 50, 55, 60,
                65, 70, 75, 80, 90, 100, 125, 150, 200, 250, 300, 350, 400, 450, 500, 600, 700, 800, 900, 1000, 1200,
                1400, 1600, 1800, 2000, 2200, 2400, 2600, 2800, 3000, 3200, 3400, 3600, 3800, 4000, 4200, 4400, 4600,
                4800, 5000, 5200, 5400, 5600, 5800, 6000, 6200, 6400, 6600, 6800, 7000, 7200, 7400, 7600, 7800, 8000,
                8200, 8400, 8600, 8800, 9000, 9200, 9400, 9600, 9800, 10000]   # generate_transitions.py:495-502
DEFAULT_BOTS = [2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 25, 30, 35, 40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100, 110,
                120, 130, 140, 150, 160, 180, 200, 220, 240, 260, 280, 300, 320, 340, 360, 380, 400, 425, 450, 475,
                500]                                                          # :503-506
DEFAULT_MUTS = [m * 10.0 ** e for e in range(-12, -2) for m in (1.0, 2.5, 5.0, 7.5)] + [1.0e-2, 2.5e-2, 5e-2, 7.5e-2]


def get_breaks_symmetric(N: int, uniform_weight: float, min_bin_size: float) -> np.ndarray:
    """Bin lower bounds, symmetric about N/2 (generate_transitions.py:133-192)."""
    if N % 2 != 0:
        raise ValueError("population size (N) must be even")
    w = uniform_weight
    u = np.ones(N - 1)
    u /= u.sum()
    s = 1 / np.arange(1, N)
    s /= s.sum()
    interior = w * u + (1 - w) * s
    breaks = [0, 1]
    cur = 0.0
    for i in range(1, N):
        cur += interior[i - 1]
        if cur >= min_bin_size:
            breaks.append(i + 1)
            cur = 0.0
        if i >= N / 2 - 1:
            break
    half = N // 2
    if breaks[-1] != half:
        breaks.append(half)
    breaks.append(half + 1)
    lesser = [b for b in breaks[::-1] if b < half]
    for b in lesser[:-1]:
        breaks.append(N - b + 1)
    return np.array(breaks, dtype=np.int64)


def get_binned_frequencies(N: int, breaks: np.ndarray) -> np.ndarray:
    """generate_transitions.py:209-214."""
    full = np.arange(N + 1) / N
    lengths = np.concatenate((np.diff(breaks), [N + 1 - breaks[-1]]))
    return np.add.reduceat(full, breaks) / lengths


def bin_matrix(P: np.ndarray, breaks: np.ndarray) -> np.ndarray:
    """Column sums per bin, rows taken at the bin's middle state (mean of the two middle states for even-sized
    bins) -- generate_transitions.py:194-208."""
    breaks = np.asarray(breaks)
    colsum = np.add.reduceat(P, breaks, axis=1)
    lengths = np.concatenate((np.diff(breaks), [P.shape[1] - breaks[-1]]))
    lo = breaks + np.floor((lengths - 1) / 2).astype(np.int64)
    hi = breaks + np.ceil((lengths - 1) / 2).astype(np.int64)
    out = np.where((lengths % 2 == 0)[:, None], (colsum[lo, :] + colsum[hi, :]) / 2, colsum[lo, :])
    return out


def wright_fisher_matrix(N: int, s: float, u: float, v: float) -> np.ndarray:
    """generate_transitions.py:57-67."""
    i = np.arange(N + 1)
    p = i / N
    pmut = (1 - u) * p + v * (1 - p)
    pstar = np.minimum(pmut * (1 + s) / (pmut * (1 + s) + 1 - pmut), 1.0)
    js = np.arange(N + 1)
    return np.exp(st.binom.logpmf(js[None, :], N, pstar[:, None]))


def var_n_matrix(N1: int, N2: int, N0: int, mu: float) -> np.ndarray:
    """mope/_transition.pyx:6-33: (N0+1)^2 matrix, row i<=N1 = Binomial(N2, mutated i/N1) pmf."""
    P = np.zeros((N0 + 1, N0 + 1))
    f = np.arange(N1 + 1) / N1
    f = (1 - mu) * f + (1 - f) * mu
    P[: N1 + 1, : N2 + 1] = st.binom.pmf(np.arange(N2 + 1)[None, :], N2, f[:, None])
    return P


class _Mat:
    """Matrix chain backend: the library's CUDA kernels (torch tensors as buffers) when a GPU is present, NumPy otherwise."""

    def __init__(self, device: Optional[str]):
        self.torch = None
        if device is None:
            try:
                import torch
                if torch.cuda.is_available():
                    device = "cuda"
            except Exception:  # pragma: no cover
                device = None
        if device is not None and device != "cpu":
            import torch
            from . import _lib
            self.torch = torch
            self.device = torch.device(device)
            self.lib = _lib.lib()           # raises if the CUDA library is missing
            self._check = _lib.check
            self.gemms = 0

    # -- device helpers ----------------------------------------------------------------------------------------------
    def _new(self, n: int):
        """n x n matrix with an even leading dimension (16-byte row pitch for the tensor maps)."""
        ld = n + (n & 1)
        return self.torch.empty((n, ld), dtype=self.torch.float64, device=self.device)

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def put(self, a):
        if not self.torch:
            return np.asarray(a)
        a = np.asarray(a, dtype=np.float64)
        out = self._new(a.shape[0])
        out[:, : a.shape[1]] = self.torch.as_tensor(a, device=self.device)
        if out.shape[1] > a.shape[1]:
            out[:, a.shape[1]:] = 0.0
        return out

    def get(self, a):
        return a[:, : a.shape[0]].cpu().numpy() if self.torch else a

    def mm(self, a, b):
        """a . b (square, same size)."""
        if not self.torch:
            return a @ b
        n, ld = a.shape
        bt = self._new(n)
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mope_b200_gen_transpose(b.data_ptr(), bt.data_ptr(), n, n, ld, ld, self._stream()), "gen_transpose")
            out = self._new(n)
            self._check(self.lib.mope_b200_gen_dgemm_tn(a.data_ptr(), bt.data_ptr(), out.data_ptr(), n, n, n, ld, ld, ld, self._stream()),
                        "gen_dgemm_tn")
        self.gemms += 1
        return out

    def _binom(self, n_states: int, rows_valid: int, n_trials: int, p: np.ndarray):
        out = self._new(n_states)
        p = np.ascontiguousarray(p, dtype=np.float64)
        cache = self.__dict__.setdefault("_lc_cache", {})
        if n_trials not in cache:   # ln C(n, k) exactly as scipy.stats.binom._logpmf forms it (the reference's pmf)
            k = np.arange(n_trials + 1, dtype=np.float64)
            cache[n_trials] = np.ascontiguousarray(gammaln(n_trials + 1.0) - (gammaln(k + 1.0) + gammaln(n_trials - k + 1.0)))
        lc = cache[n_trials]
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mope_b200_gen_binom_matrix(out.data_ptr(), out.shape[1], n_states, n_states, rows_valid, n_trials,
                                                            p.ctypes.data, lc.ctypes.data, self._stream()), "gen_binom_matrix")
        return out

    def wright_fisher(self, N: int, s: float, u: float, v: float):
        """wright_fisher_matrix (generate_transitions.py:57-67); on the device by binom_matrix_kernel."""
        if not self.torch:
            return wright_fisher_matrix(N, s, u, v)
        p = np.arange(N + 1) / N
        pmut = (1 - u) * p + v * (1 - p)
        pstar = np.minimum(pmut * (1 + s) / (pmut * (1 + s) + 1 - pmut), 1.0)
        return self._binom(N + 1, N + 1, N, pstar)

    def binner(self, n_states: int, breaks: np.ndarray):
        """bin_matrix for device matrices (bin_matrix_kernel); only the F x F result leaves the device."""
        if not self.torch:
            return lambda P: bin_matrix(P, breaks)
        t = self.torch
        F = int(np.asarray(breaks).shape[0])
        brk = t.as_tensor(np.asarray(breaks, dtype=np.int32), device=self.device)

        def run(P):
            out = t.empty((F, F), dtype=t.float64, device=self.device)
            with t.cuda.device(self.device):
                self._check(self.lib.mope_b200_gen_bin_matrix(P.data_ptr(), P.shape[1], n_states, brk.data_ptr(), F, out.data_ptr(),
                                                              self._stream()), "gen_bin_matrix")
            return out.cpu().numpy()
        return run

    def var_n(self, N1: int, N2: int, N0: int, mu: float):
        """var_n_matrix (_transition.pyx:6-33) on the device."""
        if not self.torch:
            return var_n_matrix(N1, N2, N0, mu)
        f = np.arange(N1 + 1) / N1
        f = (1 - mu) * f + (1 - f) * mu
        return self._binom(N0 + 1, N1 + 1, N2, f)


def _power_from_cache(mat: _Mat, squares: list, n: int):
    """P^n from cached P^(2^k)."""
    out = None
    k = 0
    while n:
        if n & 1:
            out = squares[k] if out is None else mat.mm(out, squares[k])
        n >>= 1
        k += 1
        if n and k >= len(squares):
            squares.append(mat.mm(squares[-1], squares[-1]))
    return out


def generate_tables(N: int = 1000, breaks=(0.5, 0.01), gens: Sequence[int] = DEFAULT_GENS, us: Sequence[float] = DEFAULT_MUTS,
                    nbs: Optional[Sequence[int]] = None, bot_us: Optional[Sequence[float]] = None, s: float = 0.0,
                    device: Optional[str] = None, progress=None) -> TransitionTables:
    """Drift grid [len(gens), len(us)] (+ bottleneck grid [len(nbs), len(bot_us)]) at WF size N with the reference's
    symmetric binning `--breaks uniform_weight min_bin_size`."""
    mat = _Mat(device)
    gens = [int(g) for g in gens]
    if sorted(gens) != gens or len(set(gens)) != len(gens):
        raise ValueError("generations must be strictly ascending")
    br = get_breaks_symmetric(N, breaks[0], breaks[1])
    freqs = get_binned_frequencies(N, br)
    F = br.shape[0]
    drift = np.empty((len(gens), len(us), F, F))
    binned = mat.binner(N + 1, br)
    for j, u in enumerate(us):
        P = mat.wright_fisher(N, s, u, u)
        squares = [P]
        cur = _power_from_cache(mat, squares, gens[0])
        drift[0, j] = binned(cur)
        for i in range(1, len(gens)):
            step = _power_from_cache(mat, squares, gens[i] - gens[i - 1])
            cur = mat.mm(cur, step)
            drift[i, j] = binned(cur)
        if progress:
            progress("drift u=%g done" % u)
    kw = {}
    if nbs is not None:
        bot_us = list(us if bot_us is None else bot_us)
        bot = np.empty((len(nbs), len(bot_us), F, F))
        for j, u in enumerate(bot_us):
            for i, Nb in enumerate(nbs):
                bot[i, j] = binned(bottleneck_matrix(N, int(Nb), u, mat, on_device=True))
            if progress:
                progress("bottleneck u=%g done" % u)
        kw = dict(bot=bot, nbs=np.array(nbs, dtype=np.float64), bot_us=np.array(bot_us, dtype=np.float64))
    return TransitionTables(drift=drift, gens=np.array(gens, dtype=np.float64), us=np.array(us, dtype=np.float64), N=float(N),
                            freqs=freqs, breaks=br, **kw)


def generate_selection_tables(N: int = 1000, breaks=(0.5, 0.01), gens: Sequence[int] = DEFAULT_GENS, ss: Sequence[float] = (0.0,),
                              u: float = 0.0, v: float = 0.0, device: Optional[str] = None, progress=None) -> TransitionTables:
    """Selection table set (the files `mope generate-commands --selection` asks for, generate_transitions.py:545-556: one
    chain per selection coefficient, u = v = 0): drift grid [len(gens), len(ss)] whose second axis (`us` of the returned
    tables) holds the selection coefficients -- what TransitionData(selection=True) builds from the `s` attribute.  The
    selection model needs negative coefficients and 0 on the grid (Inference.loglike maps |x| < 10 to alpha = 0)."""
    mat = _Mat(device)
    gens = [int(g) for g in gens]
    ss = [float(x) for x in ss]
    if sorted(gens) != gens or len(set(gens)) != len(gens) or sorted(ss) != ss or len(set(ss)) != len(ss):
        raise ValueError("generations and selection coefficients must be strictly ascending")
    br = get_breaks_symmetric(N, breaks[0], breaks[1])
    freqs = get_binned_frequencies(N, br)
    F = br.shape[0]
    drift = np.empty((len(gens), len(ss), F, F))
    binned = mat.binner(N + 1, br)
    for j, sc in enumerate(ss):
        P = mat.wright_fisher(N, sc, u, v)
        squares = [P]
        cur = _power_from_cache(mat, squares, gens[0])
        drift[0, j] = binned(cur)
        for i in range(1, len(gens)):
            cur = mat.mm(cur, _power_from_cache(mat, squares, gens[i] - gens[i - 1]))
            drift[i, j] = binned(cur)
        if progress:
            progress("selection s=%g done" % sc)
    return TransitionTables(drift=drift, gens=np.array(gens, dtype=np.float64), us=np.array(ss, dtype=np.float64),
                            N=float(N), freqs=freqs, breaks=br)


def bottleneck_matrix(N: int, Nb: int, mu: float, mat: Optional[_Mat] = None, on_device: bool = False):
    """Instantaneous bottleneck to Nb followed by doubling back to N (generate_transitions.py:29-54)."""
    if Nb >= N:
        raise ValueError("Nb must be < N")
    mat = mat or _Mat("cpu")
    num_doublings = int(np.floor(np.log2(N / Nb)))
    num_gens = int(np.ceil(np.log2(N / Nb)))
    Ns = [N, Nb] + [Nb * 2 ** d for d in range(1, num_doublings + 1)]
    if num_gens > num_doublings:
        Ns.append(N)
    fac = lambda a, b: mat.var_n(a, b, N, mu)   # noqa: E731
    P = fac(Ns[0], Ns[1])
    for N1, N2 in zip(Ns[1:-1], Ns[2:]):
        P = mat.mm(P, fac(N1, N2))
    return P if on_device else mat.get(P)

"""CPU oracle for the mope ontogenetic-phylogeny log-likelihood.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
`--impl reference` legs may import this module; nothing in mope_b200/ does, and the
product path never falls back to it.

This is an independent NumPy restatement of the algorithm in ammodramus/mope
(reference paths are relative to /root/reference).  It is pinned against outputs of the
reference itself executed in the build container (tests/golden/make_golden.py generates
tests/golden/*.npz from the patched reference install in baseline/_ref; tests/
test_oracle_golden.py checks this module against them).  The reference's own tests hold
no golden values for the non-selection likelihood, so those generated
vectors are the pin.

Numerical stand-ins: the reference's binomial pmf is GSL's gsl_ran_binomial_pdf (GSL is a
dependency with no pinned version, setup.py:30-38); its published algorithm
(exp(lnchoose + k log p + (n-k) log1p(-p)) with the k>n, p==0, p==1 short-cuts) is restated
in `binom_pmf`.  The root prior uses scipy.stats.beta.cdf exactly as the reference does.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import scipy.stats as st
from scipy.special import gammaln

# ----------------------------------------------------------------------------------------
# transition tables
# ----------------------------------------------------------------------------------------


def check_transition_matrix(P: np.ndarray) -> None:
    """In-place clamp to [0,1] and renormalise rows whose (sequential) sum exceeds 1.

    mope/_util.pyx:129-148.  Vectorised except for the row sum, which must be the same
    left-to-right accumulation as the reference loop (np.cumsum is sequential).
    """
    np.clip(P, 0.0, 1.0, out=P)
    rowsum = np.cumsum(P, axis=1)[:, -1]
    big = rowsum > 1
    if big.any():
        P[big, :] /= rowsum[big, None]


def bilinear_blend(q11, q12, q21, q22, w11, w12, w21, w22):
    """Element-wise blend in the reference's evaluation order (mope/_interp.pyx:30-35):
    ((q11*w11 + q21*w21) + q12*w12) + q22*w22."""
    return q11 * w11 + q21 * w21 + q12 * w12 + q22 * w22


class _LRU(OrderedDict):
    """Same eviction behaviour as lru-dict's LRU as used by the reference."""

    def __init__(self, size):
        super().__init__()
        self.size = size

    def get_touch(self, key):
        val = self[key]
        self.move_to_end(key)
        return val

    def put(self, key, val):
        self[key] = val
        self.move_to_end(key)
        while len(self) > self.size:
            self.popitem(last=False)


class TransitionData:
    """Drift tables keyed (generation, mutation rate).  mope/transition_data_2d.py:22-368.

    `mats[g, u]` is the F x F matrix stored for gens[g], us[u] (both sorted ascending).
    """

    def __init__(self, mats, gens, us, N, freqs, breaks):
        self.mats = np.asarray(mats, dtype=np.float64)
        self.gens = np.asarray(gens, dtype=np.float64)
        self.us = np.asarray(us, dtype=np.float64)
        assert np.all(np.diff(self.gens) > 0) and np.all(np.diff(self.us) > 0)
        self.N = N
        self.freqs = np.asarray(freqs, dtype=np.float64)
        self.breaks = np.asarray(breaks)
        self.F = self.mats.shape[-1]
        self.cache = _LRU(500)  # transition_data_2d.py:39
        # transition_data_2d.py:105-108
        self.min_coal_time = self.gens.min() / N
        self.max_coal_time = self.gens.max() / N
        self.min_mutsel_rate = self.us.min() * 2 * N
        self.max_mutsel_rate = self.us.max() * 2 * N

    def clear_cache(self):
        self.cache.clear()

    def get_transition_probabilities_2d(self, scaled_time, scaled_mut):
        """transition_data_2d.py:133-146: LRU lookup, then the check runs on EVERY return
        (also on cache hits, mutating the cached array again)."""
        if not np.isfinite(scaled_time):
            raise ValueError("scaled_time is not finite")
        key = (float(scaled_time), float(scaled_mut))
        if key in self.cache:
            val = self.cache.get_touch(key)
        else:
            val = self.interpolate(scaled_time, scaled_mut)
            self.cache.put(key, val)
        check_transition_matrix(val)
        return val

    def interpolate(self, scaled_time, scaled_mut):
        """transition_data_2d.py:148-238 (range checks, identity short-cut, left searchsorted,
        bilinear weights; index 0 wraps to -1 with a zero weight exactly like the reference)."""
        if not np.isfinite(scaled_time):
            raise ValueError("scaled_time is not finite")
        if not np.isfinite(scaled_mut):
            raise ValueError("scaled_mutsel is not finite")
        if scaled_time > self.max_coal_time:
            raise ValueError("drift time too large: {} > {} (max)".format(scaled_time, self.max_coal_time))
        if scaled_mut < self.min_mutsel_rate:
            raise ValueError("mutsel rate too small: {} < {} (min)".format(scaled_mut, self.min_mutsel_rate))
        if scaled_mut > self.max_mutsel_rate:
            raise ValueError("mutsel rate too large: {} > {} (max)".format(scaled_mut, self.max_mutsel_rate))
        t = self.N * scaled_time
        if t < self.gens[0]:
            return np.diag(np.ones(self.F))
        gi = int(np.searchsorted(self.gens, t))
        x = scaled_mut / (2.0 * self.N)
        xi = int(np.searchsorted(self.us, x))
        t1, t2 = self.gens[gi - 1], self.gens[gi]
        x1, x2 = self.us[xi - 1], self.us[xi]
        denom = (t2 - t1) * (x2 - x1)
        w11 = (t2 - t) * (x2 - x) / denom
        w21 = (t - t1) * (x2 - x) / denom
        w12 = (t2 - t) * (x - x1) / denom
        w22 = (t - t1) * (x - x1) / denom
        return bilinear_blend(self.mats[gi - 1, xi - 1], self.mats[gi - 1, xi],
                              self.mats[gi, xi - 1], self.mats[gi, xi], w11, w12, w21, w22)


class TransitionDataBottleneck:
    """Bottleneck tables keyed (Nb, u).  mope/transition_data_bottleneck.py:20-230.
    No identity short-cut, hard range errors on both axes, and NO check_transition_matrix."""

    def __init__(self, mats, nbs, us, N):
        self.mats = np.asarray(mats, dtype=np.float64)
        self.nbs = np.asarray(nbs, dtype=np.int64)
        self.us = np.asarray(us, dtype=np.float64)
        self.N = N
        self.cache = _LRU(10000)
        self.min_nb, self.max_nb = self.nbs.min(), self.nbs.max()
        self.min_mut = self.us.min() * 2 * N
        self.max_mut = self.us.max() * 2 * N

    def get_transition_probabilities_2d(self, nb, scaled_mut):
        key = (float(nb), float(scaled_mut))
        if key in self.cache:
            return self.cache.get_touch(key)
        val = self.interpolate(nb, scaled_mut)
        self.cache.put(key, val)
        return val

    def interpolate(self, nb, scaled_mut):
        if nb < self.min_nb:
            raise ValueError("bottleneck size too small: {} < {} (min)".format(nb, self.min_nb))
        if nb > self.max_nb:
            raise ValueError("bottleneck size too large: {} > {} (max)".format(nb, self.max_nb))
        if scaled_mut < self.min_mut:
            raise ValueError("mutation rate too small: {} < {} (min)".format(scaled_mut, self.min_mut))
        if scaled_mut > self.max_mut:
            raise ValueError("mutation rate too large: {} > {} (max)".format(scaled_mut, self.max_mut))
        bi = int(np.searchsorted(self.nbs, nb))
        u = scaled_mut / (2.0 * self.N)
        ui = int(np.searchsorted(self.us, u))
        t = nb
        t1, t2 = self.nbs[bi - 1], self.nbs[bi]
        u1, u2 = self.us[ui - 1], self.us[ui]
        denom = (t2 - t1) * (u2 - u1)
        w11 = (t2 - t) * (u2 - u) / denom
        w21 = (t - t1) * (u2 - u) / denom
        w12 = (t2 - t) * (u - u1) / denom
        w22 = (t - t1) * (u - u1) / denom
        return bilinear_blend(self.mats[bi - 1, ui - 1], self.mats[bi - 1, ui],
                              self.mats[bi, ui - 1], self.mats[bi, ui], w11, w12, w21, w22)


# ----------------------------------------------------------------------------------------
# leaf emissions
# ----------------------------------------------------------------------------------------


def _libm_lgamma():
    """C libm lgamma (what the reference build's GSL stand-in calls).  CPython's math.lgamma is its own
    Lanczos code and differs by an ulp or two of lgamma(n+1) ~ 1e-11 absolute at n = 5000, which shows up
    as a relative offset of the whole pmf row."""
    import ctypes
    import ctypes.util
    try:
        lib = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        fn = lib.lgamma
        fn.restype = ctypes.c_double
        fn.argtypes = [ctypes.c_double]
        return fn
    except OSError:  # pragma: no cover
        return math.lgamma


_lgamma = _libm_lgamma()


def lnchoose(n: int, m: int) -> float:
    """GSL gsl_sf_lnchoose (published algorithm), as bound at mope/_binom.pyx:10-11."""
    if m == n or m == 0:
        return 0.0
    if m * 2 > n:
        m = n - m
    return _lgamma(n + 1.0) - _lgamma(m + 1.0) - _lgamma(n - m + 1.0)


def binom_pmf(k: float, n: float, p: float) -> float:
    """mope/_binom.pyx:20-23 -> gsl_ran_binomial_pdf(uint k, p, uint n)."""
    uk, un = int(k), int(n)  # C cast double -> unsigned int truncates
    if uk > un:
        return 0.0
    if p == 0.0:
        return 1.0 if uk == 0 else 0.0
    if p == 1.0:
        return 1.0 if uk == un else 0.0
    return math.exp(lnchoose(un, uk) + uk * math.log(p) + (un - uk) * math.log1p(-p))


def get_binom_likelihoods(X, n, freqs, min_phred_score=-1.0):
    """mope/_binom.pyx:49-92.  X, n: [L, S] (NaN count = missing tissue -> all ones).
    Vectorised over the frequency axis; scalar special cases match binom_pmf."""
    X = np.asarray(X, dtype=np.float64)
    n = np.asarray(n, dtype=np.float64)
    freqs = np.asarray(freqs, dtype=np.float64)
    L, S = X.shape
    F = freqs.shape[0]
    perr = 10.0 ** (-1.0 * min_phred_score / 10.0) if min_phred_score != -1.0 else 0.0
    p = (1.0 - perr) * freqs + (perr / 3.0) * (1.0 - freqs)
    out = np.zeros((L, S, F))
    is0 = p == 0.0
    is1 = p == 1.0
    mid = ~(is0 | is1)
    # libm log/log1p per frequency (what the reference's C code calls): k*log(p) magnifies a 1-ulp
    # difference in log(p) by the read depth, so NumPy's SIMD log is not interchangeable here
    logp = np.array([math.log(v) if 0.0 < v else -math.inf for v in p])
    log1mp = np.array([math.log1p(-v) if v < 1.0 else -math.inf for v in p])
    for i in range(L):
        for j in range(S):
            if math.isnan(X[i, j]):
                out[i, j, :] = 1.0
                continue
            uk, un = int(X[i, j]), int(n[i, j])
            if uk > un:
                continue
            row = out[i, j]
            row[is0] = 1.0 if uk == 0 else 0.0
            row[is1] = 1.0 if uk == un else 0.0
            lc = lnchoose(un, uk)
            ex = lc + uk * logp[mid] + (un - uk) * log1mp[mid]
            row[mid] = [math.exp(v) for v in ex]
    return out


def get_non_asc_probs(qualities, freqs, min_empirical_freq):
    """mope/_poisson_binom.pyx:18-90: per true frequency f, Poisson-binomial DP over reads with
    per-read alt probability a_j = f(1-e_j) + (1-f)e_j, e_j = 10^(-q_j/10); returns
    P(alt count < ceil(min_freq * nreads)).  Only columns k < K are ever read back, so the DP
    is carried on K columns (identical values; the reference's full table also writes one
    element past its row end at j = nreads-1, which is never read)."""
    q = np.asarray(qualities, dtype=np.float64)
    freqs = np.asarray(freqs, dtype=np.float64)
    nq = q.shape[0]
    K = int(math.ceil(min_empirical_freq * nq))
    err = 10.0 ** (-q / 10.0)
    out = np.zeros(freqs.shape[0])
    for i, f in enumerate(freqs):
        a = f * (1.0 - err) + (1 - f) * err
        prev = np.zeros(K + 1)
        prev[0] = 1.0 - a[0]
        if K >= 1:
            prev[1] = a[0]
        for j in range(1, nq):
            cur = np.empty_like(prev)
            cur[0] = (1 - a[j]) * prev[0]
            cur[1:] = a[j] * prev[:-1] + (1 - a[j]) * prev[1:]
            prev = cur
        acc = 0.0
        for k in range(K):
            acc += prev[k]
        out[i] = acc
    return out


# ----------------------------------------------------------------------------------------
# root prior
# ----------------------------------------------------------------------------------------


def discretize_beta_distn(a, b, N):
    """mope/likelihoods.py:27-41."""
    lower = np.concatenate(([0], np.arange(1, 2 * N, 2))) / (2 * N)
    upper = np.concatenate((np.arange(1, 2 * N + 1, 2), [2 * N])) / (2 * N)
    return st.beta.cdf(upper, a, b) - st.beta.cdf(lower, a, b)


def get_stationary_distribution_double_beta(breaks, N, ab, prob_poly):
    """mope/likelihoods.py:43-60: point masses (1-ppoly)/2 at 0 and 1, ppoly * Beta(ab,ab)
    cell masses in between, then summed into the frequency bins."""
    distn = np.zeros(N + 1)
    distn[[0, -1]] = (1.0 - prob_poly) / 2
    distn[1:-1] = discretize_beta_distn(ab, ab, N - 2) * prob_poly
    return np.add.reduceat(distn, breaks)


# ----------------------------------------------------------------------------------------
# tree
# ----------------------------------------------------------------------------------------


class Node:
    """mope/newick.py:42-102: `name:var`, `name:var*multiplier`, `name:var^` (bottleneck)."""

    def __init__(self, name, length):
        self.name = name
        self.children = []
        self.parent = None
        self.varname = None
        self.multipliername = None
        self.is_bottleneck = False
        self.cond = None
        if length is not None:
            length = length.strip()
            if "*" in length:
                parts = length.split("*")
                if len(parts) != 2:
                    raise ValueError("invalid length in tree file")
                self.varname, self.multipliername = parts
            else:
                self.varname = length
            if self.varname.endswith("^"):
                self.is_bottleneck = True
                self.varname = self.varname[:-1]

    @property
    def is_leaf(self):
        return not self.children

    def postorder(self):
        for c in self.children:
            yield from c.postorder()
        yield self


def parse_newick(s: str) -> Node:
    """Recursive-descent parser for the dialect above (first tree of the string)."""
    s = s.strip()
    if ";" in s:
        s = s.split(";")[0].strip()
    pos = 0

    def label(stop):
        nonlocal pos
        j = pos
        while j < len(s) and s[j] not in stop:
            j += 1
        tok = s[pos:j]
        pos = j
        return tok

    def node():
        nonlocal pos
        kids = []
        if pos < len(s) and s[pos] == "(":
            pos += 1
            while True:
                kids.append(node())
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError("unmatched braces")
        lab = label(",()")
        name, length = (lab.split(":", 1) + [None])[:2] if ":" in lab else (lab, None)
        nd = Node(name.strip() or None, length)
        for k in kids:
            k.parent = nd
            nd.children.append(k)
        return nd

    return node()


# ----------------------------------------------------------------------------------------
# pruning pass (mope/_likes.pyx, mope/likelihoods.py:308-403, mope/ascertainment.py:132-236)
# ----------------------------------------------------------------------------------------


def compute_rate_transition_likelihood(child, parent, lengths, mut, transitions):
    """_likes.pyx:42-56: one matrix per row (LRU de-duplicates equal times)."""
    for i in range(child.shape[0]):
        P = transitions.get_transition_probabilities_2d(lengths[i], mut)
        parent[i, :] *= np.dot(P, child[i, :])


def compute_length_transition_likelihood(child, parent, length, mut, transitions):
    """_likes.pyx:139-152 and :232-248 (bottleneck twin): one matrix, a mat-vec per row."""
    P = transitions.get_transition_probabilities_2d(length, mut)
    for i in range(child.shape[0]):
        parent[i] *= np.dot(P, child[i])


def compute_root_log_like(root, stat):
    """_likes.pyx:312-322."""
    ll = 0.0
    for i in range(root.shape[0]):
        d = np.dot(root[i], stat)
        ll += math.log(d) if d > 0 else _log_nonpos(d)
    return ll


def _log_nonpos(x):
    # C log(): log(0) = -inf, log(<0) = nan
    if x == 0:
        return -math.inf
    return math.nan


def get_individual_locus_asc_logprobs(root, stat):
    """_likes.pyx:426-442: log(1 - stat.root[2k] - stat.root[2k+1]), sequential subtraction."""
    n = root.shape[0] // 2
    out = np.zeros(n)
    for k in range(n):
        pp = 1.0
        pp -= np.dot(stat, root[2 * k])
        pp -= np.dot(stat, root[2 * k + 1])
        out[k] = math.log(pp) if pp > 0 else _log_nonpos(pp)
    return out


def _prune(tree, leaf_vectors, nrows, F, branch_lengths, mut_rates, branch_index, row_multipliers,
           transitions, bottlenecks):
    """Shared postorder walk.  leaf_vectors(name) -> [nrows, F] copy; row_multipliers(mult) ->
    [nrows] multiplier values."""
    for nd in tree.postorder():
        nd.cond = np.ones((nrows, F))
    for nd in tree.postorder():
        if nd is tree:
            break
        bi = branch_index[nd.name]
        mut = mut_rates[bi]
        child = leaf_vectors(nd.name) if nd.is_leaf else nd.cond
        parent = nd.parent.cond
        if nd.multipliername is not None:
            lengths = row_multipliers(nd.multipliername) * branch_lengths[bi]
            compute_rate_transition_likelihood(child, parent, lengths, mut, transitions)
        elif nd.is_bottleneck:
            if bottlenecks is None:
                raise ValueError("pre-computed bottleneck data needed")
            compute_length_transition_likelihood(child, parent, branch_lengths[bi], mut, bottlenecks)
        else:
            compute_length_transition_likelihood(child, parent, branch_lengths[bi], mut, transitions)
    return tree.cond


# ----------------------------------------------------------------------------------------
# Inference (mope/inference.py)
# ----------------------------------------------------------------------------------------


class Pedigree:
    """One (data, tree, ages) triple after the reference's init-time processing
    (inference.py:331-352, 385-459, 504-584)."""

    def __init__(self, data, tree_str, ages, freqs, min_phred_score=None):
        """data: dict column -> float array (plus 'family' -> array); ages: dict column -> array
        with 'family'.  Columns follow the reference TSVs (leaf, leaf_n)."""
        self.tree = parse_newick(tree_str)
        self.asc_tree = parse_newick(tree_str)
        nodes = [nd for nd in self.tree.postorder() if nd is not self.tree]
        self.leaf_names = [nd.name for nd in nodes if nd.is_leaf]
        self.branch_names = [nd.name for nd in nodes]
        if len(set(self.branch_names)) != len(self.branch_names):
            raise ValueError("node names must be unique in each tree")
        self.varname = {nd.name: nd.varname for nd in nodes}
        self.multiplier = {nd.name: nd.multipliername for nd in nodes}
        self.is_bottleneck = {nd.varname: nd.is_bottleneck for nd in nodes}
        self.multipliernames = sorted({m for m in self.multiplier.values() if m is not None})
        self.branch_index = {b: i for i, b in enumerate(self.branch_names)}

        # drop fixed loci (data.py:219-228, inference.py:456-459)
        X = np.column_stack([np.asarray(data[l], dtype=np.float64) for l in self.leaf_names])
        n = np.column_stack([np.asarray(data[l + "_n"], dtype=np.float64) for l in self.leaf_names])
        fsum = np.nansum(X / n, axis=1).astype(np.float64)
        fixed = np.isclose(fsum, X.shape[1]) | (fsum == 0.0)
        keep = ~fixed
        self.X, self.n = X[keep], n[keep]
        self.family = np.asarray(data["family"])[keep]
        self.num_loci = int(keep.sum())

        # ages / family counts / multiplier columns (inference.py:504-525)
        self.ages = {k: np.asarray(v) for k, v in ages.items()}
        fam = self.ages["family"]
        self.counts = np.array([(self.family == f).sum() for f in fam], dtype=np.int64)
        self.n_fam = fam.shape[0]
        self.row_mult = {}
        for m in self.multipliernames:
            if m in data:
                self.row_mult[m] = np.asarray(data[m], dtype=np.float64)[keep]
            else:
                vals = np.empty(self.num_loci)
                for i, f in enumerate(self.family):
                    vals[i] = self.ages[m][np.nonzero(fam == f)[0][0]]
                self.row_mult[m] = vals

        phred = -1.0 if min_phred_score is None else min_phred_score
        likes = get_binom_likelihoods(self.X, self.n, freqs, phred)  # inference.py:564-584
        self.leaf_likes = {l: likes[:, i, :].copy() for i, l in enumerate(self.leaf_names)}


class Inference:
    """Restatement of mope.inference.Inference for allele-count data, non-selection model."""

    def __init__(self, pedigrees_in, drift, bottleneck, genome_size=16569, poisson_like_penalty=1.0,
                 min_freq=0.001, log_unif_drift=False, inverse_bot_priors=False, post_is_prior=False,
                 lower_drift_limit=1e-3, upper_drift_limit=3, min_phred_score=None):
        """pedigrees_in: list of (data dict, tree string, ages dict)."""
        self.transition_data = drift
        self.bottleneck_data = bottleneck
        self.freqs = drift.freqs
        self.breaks = drift.breaks
        self.num_freqs = self.freqs.shape[0]
        self.transition_N = drift.N
        self.genome_size = genome_size
        self.poisson_like_penalty = poisson_like_penalty
        self.min_freq = min_freq
        self.log_unif_drift = log_unif_drift
        self.inverse_bot_priors = inverse_bot_priors
        self.post_is_prior = post_is_prior
        self.lower_drift_limit = lower_drift_limit
        self.upper_drift_limit = upper_drift_limit
        self.peds = [Pedigree(d, t, a, self.freqs, min_phred_score) for d, t, a in pedigrees_in]

        names = set()
        self.is_bottleneck = {}
        for p in self.peds:
            for v, b in p.is_bottleneck.items():
                if v in self.is_bottleneck and self.is_bottleneck[v] != b:
                    raise ValueError("nodes with the same variable name must have the same status "
                                     "as bottlenecks in all trees")
                self.is_bottleneck[v] = b
                names.add(v)
        self.varnames = sorted(names)
        self.num_varnames = nv = len(self.varnames)
        vidx = {v: i for i, v in enumerate(self.varnames)}
        # inference.py:434-451
        self.translate_indices = []
        for p in self.peds:
            li = [vidx[p.varname[b]] for b in p.branch_names]
            self.translate_indices.append(np.array(li + [nv + i for i in li] + [2 * nv, 2 * nv + 1], dtype=np.int32))
        self.lower, self.upper, self.ndim = self._limits()

    # inference.py:1018-1052
    def _min_max_mults(self):
        mins = {v: [] for v in self.varnames}
        maxs = {v: [] for v in self.varnames}
        for p in self.peds:
            for b in p.branch_names:
                v, m = p.varname[b], p.multiplier[b]
                if m:
                    mins[v].append(np.nanmin(p.ages[m]))
                    maxs[v].append(np.nanmax(p.ages[m]))
                else:
                    mins[v].append(1)
                    maxs[v].append(1)
        return (np.array([min(mins[v]) for v in self.varnames]),
                np.array([max(maxs[v]) for v in self.varnames]))

    # inference.py:154-258 (non-selection branch)
    def _limits(self):
        nv = self.num_varnames
        td = self.transition_data
        min_len = 0.1 * max(td.min_coal_time, self.lower_drift_limit)
        max_len = min(td.max_coal_time, self.upper_drift_limit)
        min_mults, max_mults = self._min_max_mults()
        lower_len = min_len / min_mults
        upper_len = max_len / max_mults
        isb = np.array([self.is_bottleneck[v] for v in self.varnames], dtype=bool)
        self.is_bottleneck_arr = isb
        lower_len[isb] = 2
        upper_len[isb] = 500
        lower_len = np.log10(lower_len)
        upper_len = np.log10(upper_len)
        min_mut = max(-8, np.log10(td.min_mutsel_rate))
        max_mut = min(-1, np.log10(td.max_mutsel_rate))
        lower = np.concatenate((lower_len, np.repeat(min_mut, nv), (-9, -9)))
        upper = np.concatenate((upper_len, np.repeat(max_mut, nv), (0, 0)))
        return lower, upper, 2 * nv + 2

    # inference.py:1055-1110
    def logprior(self, x):
        nv = self.num_varnames
        if np.any(x < self.lower) or np.any(x > self.upper):
            return -np.inf
        if self.log_unif_drift:
            return np.sum(np.log(1.0 / (self.upper - self.lower)))
        u, l = self.upper, self.lower
        drift_x = x[:nv]
        # the reference forms log(10**u - 10**l) over the whole vector; only the drift part can
        # broadcast against drift_x, i.e. this branch only works when sliced as below
        drift_part = drift_x * 2.3025850929940459 + 0.834032445247956 - np.log(10 ** u[:nv] - 10 ** l[:nv])
        mut_part = np.sum(np.log(1.0 / (u[nv:2 * nv] - l[nv:2 * nv])))
        root_part = np.sum(np.log(1.0 / (u[2 * nv:] - l[2 * nv:])))
        if self.inverse_bot_priors:
            isb = self.is_bottleneck_arr
            l_b, u_b, x_b = l[:nv][isb], u[:nv][isb], drift_x[isb]
            drift_part[isb] = np.log(l_b) + np.log(u_b) + -np.log(u_b - l_b) - 2 * np.log(x_b)
        return np.sum(drift_part) + mut_part + root_part

    # inference.py:1113-1129
    def poisson_log_like(self, logasc, ped):
        loglams = logasc + np.log(self.genome_size)
        lams = np.exp(loglams)
        c = ped.counts
        return np.mean(logasc), (-lams + c * loglams - gammaln(c + 1)).sum()

    def ped_log_likelihood(self, ped, branch_lengths, mut_rates, stat):
        """likelihoods.py:308-403."""
        root = _prune(ped.tree, lambda name: ped.leaf_likes[name].copy(), ped.num_loci, self.num_freqs,
                      branch_lengths, mut_rates, ped.branch_index, lambda m: ped.row_mult[m],
                      self.transition_data, self.bottleneck_data)
        return compute_root_log_like(root, stat)

    def ped_asc_logprobs(self, ped, branch_lengths, mut_rates, stat):
        """ascertainment.py:132-236."""
        F = self.num_freqs
        e0 = np.zeros(F)
        e0[self.freqs < self.min_freq] = 1.0
        eN = np.zeros(F)
        eN[self.freqs > 1 - self.min_freq] = 1.0
        eboth = np.vstack((e0, eN) * ped.n_fam)
        root = _prune(ped.asc_tree, lambda name: eboth.copy(), 2 * ped.n_fam, F, branch_lengths, mut_rates,
                      ped.branch_index, lambda m: np.repeat(np.asarray(ped.ages[m], dtype=np.float64), 2),
                      self.transition_data, self.bottleneck_data)
        return get_individual_locus_asc_logprobs(root, stat)

    # inference.py:770-910 (non-selection branch)
    def loglike(self, x, components=None):
        ll = 0.0
        ab, ppoly = 10 ** x[-2:]
        stat = get_stationary_distribution_double_beta(self.breaks, self.transition_N, ab, ppoly)
        for i, ped in enumerate(self.peds):
            params = x[self.translate_indices[i]]
            nb = len(ped.branch_names)
            lens = 10 ** params[:nb]
            muts = 10 ** params[nb:]
            ped_ll = self.ped_log_likelihood(ped, lens, muts, stat)
            logasc = self.ped_asc_logprobs(ped, lens, muts, stat)
            if components is not None:
                components.append({"root_ll": ped_ll, "log_asc": logasc.copy()})
            ped_ll -= (logasc * ped.counts).sum()
            if self.poisson_like_penalty > 0:
                _, lp = self.poisson_log_like(logasc, ped)
                ped_ll += lp * self.poisson_like_penalty
            ll += ped_ll
        return ll

    # inference.py:1132-1161
    def log_posterior(self, x):
        nv = self.num_varnames
        x = np.array(x, dtype=np.float64)
        x[nv:2 * nv] = -1.0 * np.abs(x[nv:2 * nv])
        x[-2:] = -np.abs(x[-2:])
        pr = self.logprior(x)
        if self.post_is_prior:
            v = pr
        else:
            like = -np.inf if not np.isfinite(pr) else self.loglike(x)
            v = pr + like
        if np.isnan(v):
            v = -np.inf
        return v

    def clear_cache(self):
        self.transition_data.clear_cache()

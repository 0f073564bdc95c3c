"""Host-side mirror of the reference's inference API for the log-likelihood path.

Same names, argument meaning and error behaviour as mope/inference.py for: the module-level trampolines
`logl` / `logp` / `post_clone` reading the global `inf_data` (:45-60), `Inference(...)` (:263-738, count data,
non-selection model), `Inference.loglike` (:770-910), `like_obj` (:766), `inf_bound_like_obj` (:1008-1015),
`logprior` (:1055-1110), `log_posterior` (:1132-1161) and the prior box of `_get_likelihood_limits` (:154-258).
The likelihood itself runs on the GPU through mope_b200.engine.Engine; the batched entry points
(`loglike_batch`, `log_posterior_batch`) are what `GpuPool.map` uses so that one emcee half-step is one
device call.
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Sequence

import numpy as np

from . import model as _model
from . import tables as _tables
from .engine import Engine, OutOfTableError, status_message
from .stationary import stationary_distribution_batch

inf_data = None


def logl(x):
    global inf_data
    return -1.0 * inf_data.inf_bound_like_obj(x)


def logp(x):
    global inf_data
    return inf_data.logprior(x)


def post_clone(x):
    global inf_data
    return inf_data.log_posterior(x)


def target(x):
    """mope/inference.py:62-65: the optimiser's objective (optimize_posterior's PSO / L-BFGS-B stages)."""
    return -1.0 * inf_data.penalty_bound_log_posterior(x)


def inf_bound_target(x):
    """mope/inference.py:67-70: the Nelder-Mead stage's objective."""
    return -1.0 * inf_data.log_posterior(x)


def _get_valid_start_from(sf_str):
    for v in ("prior", "true", "map"):
        if sf_str == v:
            return v
    warnings.warn("getting starting position from heuristic guess")
    return "initguess"


class Inference(object):
    def __init__(self, data_files, tree_files, age_files, transitions_file, true_parameters=None, start_from="prior",
                 data_are_freqs=False, genome_size=16569, bottleneck_file=None, poisson_like_penalty=1.0, min_freq=0.001,
                 transition_copy=None, transition_buf=None, transition_shape=None, print_debug=False,
                 log_unif_drift=False, inverse_bot_priors=False, post_is_prior=False, lower_drift_limit=1e-3,
                 upper_drift_limit=3, min_phred_score=None, selection_model=False, device=0, workspace_bytes=None,
                 _inputs_are_text=False, mult_bounds=None, root_prior=None):
        if data_are_freqs:
            raise NotImplementedError("frequency-format data (get_nearest_likelihoods_cython) is outside the GPU "
                                      "log-likelihood path; use allele counts")
        if true_parameters:
            raise NotImplementedError("--true-parameters is not supported by the GPU path")
        self.data_files, self.tree_files, self.age_files = list(data_files), list(tree_files), list(age_files)
        if not (len(self.data_files) == len(self.tree_files) == len(self.age_files)) or not self.data_files:
            raise ValueError("need one data file, one tree file and one ages file per pedigree type")
        self.data_are_freqs = data_are_freqs
        self.genome_size = genome_size
        self.transitions_file = transitions_file
        self.start_from = _get_valid_start_from(start_from)
        self.bottleneck_file = bottleneck_file
        self.poisson_like_penalty = poisson_like_penalty
        self.min_freq = min_freq
        self.print_debug = print_debug
        self.log_unif_drift = log_unif_drift
        self.inverse_bot_priors = inverse_bot_priors
        self.post_is_prior = post_is_prior
        self.lower_drift_limit = lower_drift_limit
        self.upper_drift_limit = upper_drift_limit
        self.min_phred_score = min_phred_score
        self.selection_model = bool(selection_model)
        # where the root (stationary) distribution is computed: "auto" (default) = on the device for walkers with
        # ab >= ROOT_PRIOR_HOST_BELOW, SciPy's betainc on the host below it (there the reference's CDF differences are
        # dominated by rounding noise that only the same betainc reproduces); "device" = always the device kernel (the more
        # accurate values; differs from the reference by up to ~3e-10 relative in the log-likelihood at ab ~ 1e-9);
        # "scipy" = always the host
        import os as _os
        self.root_prior = root_prior or _os.environ.get("MOPE_B200_ROOT_PRIOR", "auto")
        if self.root_prior not in ("auto", "device", "scipy"):
            raise ValueError("root_prior must be 'auto', 'device' or 'scipy'")
        self.true_params = None
        self.true_loglike = None

        # tables (inference.py:741-763)
        self.transition_data = _tables.load(transitions_file, bottleneck_file, selection=self.selection_model)
        self.bottleneck_data = self.transition_data if self.transition_data.bot is not None else None
        self.freqs = self.transition_data.get_bin_frequencies()
        self.num_freqs = self.freqs.shape[0]
        self.breaks = self.transition_data.get_breaks()
        self.transition_N = self.transition_data.get_N()

        # pedigrees
        self.tree_strings, self.peds = [], []
        for dat_fn, tree_fn, age_fn in zip(self.data_files, self.tree_files, self.age_files):
            if _inputs_are_text:
                tree_str = tree_fn.strip()
            else:
                with open(tree_fn) as fin:
                    tree_str = fin.read().strip()
            self.tree_strings.append(tree_str)
            datp = _model.read_table(dat_fn, _inputs_are_text)
            if self.selection_model:
                if "position" not in datp.columns:
                    raise ValueError('The selection model requires a "position" column in each dataframe. Could not find '
                                     'one in {}'.format("<text>" if _inputs_are_text else dat_fn))
                datp = datp.sort_values(["family", "position"], kind="stable").reset_index(drop=True)   # inference.py:341-342
            agep = _model.read_table(age_fn, _inputs_are_text)
            self.peds.append(_model.compile_pedigree(datp, tree_str, agep))

        self.leaf_names = [p.leaf_names for p in self.peds]
        self.branch_names = [p.branch_names for p in self.peds]
        self.multiplierdict = [p.multiplier_of for p in self.peds]
        self.varnamedict = [p.varname_of for p in self.peds]
        self.multipliernames = [p.multipliernames for p in self.peds]
        self.is_bottleneck = {}
        for p in self.peds:
            for v, b in p.bottleneck_of_var.items():
                if v in self.is_bottleneck and self.is_bottleneck[v] != b:
                    raise ValueError("nodes with the same variable name must have the same status as bottlenecks in all trees")
                self.is_bottleneck[v] = b
        self.varnames = sorted(self.is_bottleneck.keys())
        self.num_varnames = len(self.varnames)
        self.var_indices = {v: i for i, v in enumerate(self.varnames)}
        self.num_branches = [len(b) for b in self.branch_names]
        self.branch_indices = [{br: i for i, br in enumerate(b)} for b in self.branch_names]
        self.leaf_indices = [{br: i for i, br in enumerate(b)} for b in self.leaf_names]
        nv = self.num_varnames
        self.translate_indices = []
        for p in self.peds:
            li = [self.var_indices[p.varname_of[b]] for b in p.branch_names]
            self.translate_indices.append(np.array(li + [nv + i for i in li] + [2 * nv, 2 * nv + 1], dtype=np.int32))
        self.asc_ages = [p.ages for p in self.peds]
        self.num_asc_combs = [p.n_fam for p in self.peds]
        self.num_loci = [p.num_loci for p in self.peds]
        if any(self.is_bottleneck.values()) and self.bottleneck_data is None:
            # the reference raises at the first evaluation (likelihoods.py:381-382)
            warnings.warn("tree has bottleneck branches but no bottleneck tables were given")

        # (locus-sharded runs pass the multiplier range of ALL families so that every shard has the same prior box)
        self.min_mults, self.max_mults = self.get_min_max_mults() if mult_bounds is None else \
            (np.asarray(mult_bounds[0], dtype=np.float64), np.asarray(mult_bounds[1], dtype=np.float64))
        self.header = "\t".join(["ll"] + [v + "_l" for v in self.varnames] + [v + "_m" for v in self.varnames]
                                + ["log10ab", "log10polyprob"])
        self.lower, self.upper, self.ndim = self._get_likelihood_limits()
        if self.selection_model:
            # per-position selection coefficients, DFEs, the selection prior box and header (mope_b200/selection.py)
            from . import selection as _selection
            _selection.init_selection(self)

        self.engine = Engine(self.transition_data, self.peds, [self.var_indices] * len(self.peds), nv, min_phred_score,
                             min_freq, genome_size, poisson_like_penalty, device=device, workspace_bytes=workspace_bytes)

    # ------------------------------------------------------------------------------------------
    def get_min_max_mults(self):
        """inference.py:1018-1052."""
        mins = {v: [] for v in self.varnames}
        maxes = {v: [] for v in self.varnames}
        for p in self.peds:
            for br in p.branch_names:
                v, m = p.varname_of[br], p.multiplier_of[br]
                if m:
                    vals = p.ages[m].values
                    mins[v].append(np.nanmin(vals))
                    maxes[v].append(np.nanmax(vals))
                else:
                    mins[v].append(1)
                    maxes[v].append(1)
        return (np.array([min(mins[v]) for v in self.varnames]), np.array([max(maxes[v]) for v in self.varnames]))

    def _get_likelihood_limits(self):
        """inference.py:154-258, non-selection branch."""
        nv = self.num_varnames
        td = self.transition_data
        min_allowed_len = 0.1 * max(td.get_min_coal_time(), self.lower_drift_limit)
        max_allowed_len = min(td.get_max_coal_time(), self.upper_drift_limit)
        lower_len = min_allowed_len / self.min_mults
        upper_len = max_allowed_len / self.max_mults
        isb = np.array([self.is_bottleneck[v] for v in self.varnames], dtype=bool)
        self.is_bottleneck_arr = isb
        lower_len[isb] = 2
        upper_len[isb] = 500
        lower_len, upper_len = np.log10(lower_len), np.log10(upper_len)
        with np.errstate(invalid="ignore", divide="ignore"):   # selection tables hold coefficients <= 0; the selection box replaces these
            min_mut = max(-8, np.log10(td.get_min_mutsel_rate()))
            max_mut = min(-1, np.log10(td.get_max_mutsel_rate()))
        lower = np.concatenate((lower_len, np.repeat(min_mut, nv), (-9, -9)))
        upper = np.concatenate((upper_len, np.repeat(max_mut, nv), (0, 0)))
        return lower, upper, 2 * nv + 2

    # ------------------------------------------------------------------------------------------
    # likelihood
    # ------------------------------------------------------------------------------------------
    ROOT_PRIOR_HOST_BELOW = 1e-6   # alphabeta below which "auto" keeps the reference's own betainc (see __init__)

    def root_prior_host_mask(self, X: np.ndarray) -> np.ndarray:
        """Walkers whose root distribution comes from SciPy's betainc on the host (see `root_prior` in __init__)."""
        if self.root_prior == "scipy":
            return np.ones(X.shape[0], dtype=bool)
        if self.root_prior == "device":
            return np.zeros(X.shape[0], dtype=bool)
        return X[:, -2] < np.log10(self.ROOT_PRIOR_HOST_BELOW)

    def stationary_distributions(self, X: np.ndarray) -> np.ndarray:
        ab = 10 ** X[:, -2]
        pp = 10 ** X[:, -1]
        return stationary_distribution_batch(self.breaks, self.transition_N, ab, pp)

    def loglike_batch(self, X, raise_on_error: bool = True, return_status: bool = False):
        """Inference.loglike for every row of X [W, ndim] in one device call."""
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        if X.shape[1] != self.ndim:
            raise ValueError("expected parameter vectors of length {}".format(self.ndim))
        if self.selection_model:   # per-position matrices: one device evaluation per vector
            return np.array([self.loglike(x.copy()) for x in X])
        if self.root_prior != "scipy":
            ll, status = self.engine.loglike_hybrid(X, self.root_prior_host_mask(X), self.stationary_distributions)
        elif X.shape[0] >= 32:
            # large batches: root distributions (host, SciPy) for the next chunk overlap the device work of the current one
            ll, status = self.engine.loglike_pipelined(X, self.stationary_distributions)
        else:
            ll, status = self.engine.loglike(X, self.stationary_distributions(X))
        if raise_on_error and status.any():
            w = int(np.nonzero(status)[0][0])
            raise OutOfTableError("walker {}: {}".format(w, status_message(int(status[w]))))
        return (ll, status) if return_status else ll

    def loglike_components(self, X):
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        stat = self.stationary_distributions(X)
        ll, status, root, lasc = self.engine.loglike(X, stat, want_components=True)
        out, o = [], 0
        for k, p in enumerate(self.peds):
            u = lasc[:, o:o + p.n_asc]
            o += p.n_asc
            out.append(dict(root_ll=root[:, k], log_asc=u[:, p.fam_asc]))
        return ll, status, out

    def loglike(self, orig_params):
        if self.selection_model:
            from . import selection as _selection
            return _selection.loglike(self, orig_params)
        return float(self.loglike_batch(np.asarray(orig_params, dtype=np.float64)[None, :])[0])

    def like_obj(self, varparams):
        return -self.loglike(varparams)

    def inf_bound_like_obj(self, x):
        if self.print_debug and np.any(x < self.lower):
            print("inf:", x)
            return np.inf
        if self.print_debug and np.any(x > self.upper):
            print("inf:", x)
            return np.inf
        return self.like_obj(x)

    def _get_bounds_penalty(self, x):
        """inference.py:1170-1192: parameters clipped into the prior box and a positive quadratic penalty for what was
        clipped, in units of the box width."""
        x = np.asarray(x, dtype=np.float64)
        scaling = self.upper - self.lower
        lower_penalty = ((np.maximum(self.lower - x, 0.0) / scaling) ** 2).sum()
        upper_penalty = ((np.maximum(x - self.upper, 0.0) / scaling) ** 2).sum()
        good_params = np.minimum(np.maximum(x, self.lower), self.upper)
        return good_params, lower_penalty + upper_penalty

    def penalty_bound_like_obj(self, x):
        """inference.py:1018-1020."""
        good_x, penalty = self._get_bounds_penalty(x)
        return self.like_obj(good_x) + penalty

    def penalty_bound_log_posterior(self, x):
        """inference.py:1163-1167."""
        good_x, penalty = self._get_bounds_penalty(x)
        return self.log_posterior(good_x) - penalty

    def debug_loglike_components(self, varparams):
        """inference.py:925-972: per pedigree [tree log-likelihood, per-family log ascertainment probabilities, family counts,
        (mean log ascertainment probability, Poisson term when the penalty is on,) total, log prior] -- one device call."""
        varparams = np.asarray(varparams, dtype=np.float64)
        _ll, _status, comps = self.loglike_components(varparams[None, :])
        out = []
        for i, ped in enumerate(self.peds):
            tll = float(comps[i]["root_ll"][0])
            log_asc = np.array(comps[i]["log_asc"][0])
            counts = np.array(ped.ages["count"].values)
            tcomps = [tll, log_asc, counts.copy()]
            tll -= (log_asc * counts).sum()
            if self.poisson_like_penalty > 0:
                from scipy.special import gammaln
                loglams = log_asc + np.log(self.genome_size)
                logpoisson = (-np.exp(loglams) + counts * loglams - gammaln(counts + 1)).sum()   # inference.py:1113-1129
                tll += logpoisson * self.poisson_like_penalty
                tcomps += [np.mean(log_asc), logpoisson]
            tcomps += [tll, self.logprior(varparams)]
            out.append(tcomps)
        return out

    # ------------------------------------------------------------------------------------------
    # prior / posterior
    # ------------------------------------------------------------------------------------------
    def logprior(self, x):
        """inference.py:1055-1110.  (The reference's uniform-drift branch cannot run as written: it subtracts a
        length-ndim vector from the length-n_var drift part, :1079-1080; the drift slice is used here.)"""
        x = np.asarray(x, dtype=np.float64)
        nv = self.num_varnames
        if np.any(x < self.lower):
            return -np.inf
        if np.any(x > self.upper):
            return -np.inf
        if self.log_unif_drift:
            return np.sum(np.log(1.0 / (self.upper - self.lower)))
        u, l = self.upper, self.lower
        drift_x = x[:nv]
        drift_part_arr = drift_x * 2.3025850929940459 + 0.834032445247956 - np.log(10 ** u[:nv] - 10 ** l[:nv])
        mut_part = np.sum(np.log(1.0 / (u[nv:2 * nv] - l[nv:2 * nv])))
        root_part = np.sum(np.log(1.0 / (u[2 * nv:] - l[2 * nv:])))
        if self.inverse_bot_priors:
            isb = self.is_bottleneck_arr
            l_b, u_b, x_b = l[:nv][isb], u[:nv][isb], drift_x[isb]
            drift_part_arr[isb] = np.log(l_b) + np.log(u_b) + -np.log(u_b - l_b) - 2 * np.log(x_b)
        return np.sum(drift_part_arr) + mut_part + root_part

    def fold(self, X):
        """Sign folding of log_posterior (inference.py:1140-1145)."""
        X = np.array(X, dtype=np.float64, copy=True)
        nv = self.num_varnames
        X[..., nv:2 * nv] = -1.0 * np.abs(X[..., nv:2 * nv])
        X[..., -2:] = -np.abs(X[..., -2:])
        return X

    def log_posterior_batch(self, X):
        X = self.fold(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        pr = np.array([self.logprior(x) for x in X])
        v = pr.copy()
        if not self.post_is_prior:
            ok = np.isfinite(pr)
            like = np.full(X.shape[0], -np.inf)
            if ok.any():
                like[ok] = self.loglike_batch(X[ok])
            v = pr + like
        v[np.isnan(v)] = -np.inf
        return v

    def log_posterior(self, x):
        v = float(self.log_posterior_batch(np.asarray(x, dtype=np.float64)[None, :])[0])
        if self.print_debug:
            print("@@", v, self.logprior(self.fold(x)))
        return v

    # ------------------------------------------------------------------------------------------
    # MCMC driver (inference.py:1248-1363) on top of the batched posterior
    # ------------------------------------------------------------------------------------------
    def translate_positions(self, pos):
        """mope/_util.pyx:29-53: log10 -> natural scale for the chain file; drift within one decade of its lower bound is
        reported as 0 (point mass) with a NaN mutation rate; root parameters stay in log10."""
        import math
        pos = np.asarray(pos, dtype=np.float64)
        ret = pos.copy()
        nv = self.num_varnames
        zero = pos[:, :nv] <= self.lower[:nv] + 1
        # the reference's 10**x is C pow() on scalars; NumPy's vectorised power can differ from it in the last bit, which
        # would show in the printed chain -- so: libm pow per element
        for i in range(pos.shape[0]):
            for j in range(nv):
                if zero[i, j]:
                    ret[i, j] = 0.0
                    ret[i, j + nv] = np.nan
                else:
                    ret[i, j] = math.pow(10.0, pos[i, j])
                    ret[i, j + nv] = math.pow(10.0, pos[i, j + nv])
        return ret

    @staticmethod
    def print_csv_lines(arr, lnprobs, file=None):
        """mope/_util.pyx:17-26 (same print calls, hence the same number formatting)."""
        import sys
        file = file or sys.stdout
        npos, ncols = arr.shape
        for i in range(npos):
            print(lnprobs[i], "\t", sep="", end="", file=file)
            for j in range(ncols - 1):
                print(arr[i, j], "\t", sep="", end="", file=file)
            print(arr[i, ncols - 1], file=file)

    def _get_initial_mcmc_position(self, num_walkers, prev_chain, rng, file=None):
        """inference.py:1248-1319.  Fresh start: uniform in the prior box (log10 units) until every walker has a finite
        likelihood and prior.  `prev_chain`: last num_walkers rows of a chain file; the file holds translated values, which
        are mapped back to log10 (zero drift -> half a decade above its lower bound, NaN mutation -> redrawn); the reference
        feeds the translated values back unchanged, which restarts in natural units."""
        import sys
        import pandas as pd
        nv = self.num_varnames
        if prev_chain is not None:
            try:
                prev = pd.read_csv(prev_chain, sep="\t", header=None, dtype=np.float64, comment="#")
            except Exception:
                prev = pd.read_csv(prev_chain, sep="\t", header=0, dtype=np.float64, comment="#")
            vals = prev.iloc[-num_walkers:, 1:].values
            if vals.shape != (num_walkers, self.ndim):
                raise ValueError("previous chain does not hold {} walkers of dimension {}".format(num_walkers, self.ndim))
            pos = np.empty_like(vals)
            drift, mut = vals[:, :nv], vals[:, nv:2 * nv]
            zero = ~(drift > 0)
            with np.errstate(divide="ignore", invalid="ignore"):
                pos[:, :nv] = np.where(zero, self.lower[:nv] + 0.5, np.log10(drift))
                redraw = rng.uniform(np.tile(self.lower[nv:2 * nv], (num_walkers, 1)), np.tile(self.upper[nv:2 * nv], (num_walkers, 1)))
                pos[:, nv:2 * nv] = np.where(np.isfinite(mut) & (mut > 0), np.log10(mut), redraw)
            pos[:, 2 * nv:] = vals[:, 2 * nv:]
            return pos
        while True:
            pos = rng.uniform(np.tile(self.lower, (num_walkers, 1)), np.tile(self.upper, (num_walkers, 1)))
            if not self.log_unif_drift:
                pos[:, :nv] = np.log10(rng.uniform(np.tile(10 ** self.lower[:nv], (num_walkers, 1)),
                                                   np.tile(10 ** self.upper[:nv], (num_walkers, 1))))
            logl_val = self.loglike_batch(pos, raise_on_error=False)
            logp_val = np.array([self.logprior(x) for x in pos])
            if np.all(np.isfinite(logl_val)) and np.all(np.isfinite(logp_val)):
                return pos
            if not np.all(np.isfinite(logl_val)):
                print("# warning: attempted start at initial position where log-likelihood is not finite", file=file or sys.stdout)
            if not np.all(np.isfinite(logp_val)):
                print("# warning: attempted start at initial position where prior is not finite", file=file or sys.stdout)

    def run_mcmc(self, num_iter, num_walkers, num_processes=1, mpi=False, prev_chain=None, start_from="prior", init_norm_sd=0.2,
                 chain_alpha=2.0, init_pos=None, pool=None, seed=None, file=None, device_sampler=False):
        """inference.py:1322-1363 with the built-in stretch-move sampler; every half-step is one device call.  `pool`,
        `num_processes` and `mpi` are accepted for signature compatibility and ignored (the GPU is the pool).
        device_sampler=True keeps the whole step on the GPU (ensemble.DeviceEnsembleSampler): only the printed ensembles
        come back."""
        import sys
        from .ensemble import DeviceEnsembleSampler, EnsembleSampler
        global inf_data
        inf_data = self
        file = file or sys.stdout
        print("# " + " ".join(sys.argv), file=file)
        rng = np.random.RandomState(seed)
        if init_pos is None:
            init_pos = self._get_initial_mcmc_position(num_walkers, prev_chain, rng, file=file)
        print(self.header, file=file)
        if device_sampler:
            sampler = DeviceEnsembleSampler(self, num_walkers, a=chain_alpha, seed=None if seed is None else seed + 1)
        else:
            sampler = EnsembleSampler(num_walkers, self.ndim, self.log_posterior_batch, a=chain_alpha,
                                      seed=None if seed is None else seed + 1)
        for ps, lnprobs, _state in sampler.sample(init_pos, iterations=num_iter):
            self.print_csv_lines(self.translate_positions(ps), lnprobs, file=file)
        print("# mean acceptance across chains:", np.mean(sampler.acceptance_fraction), file=file)
        return sampler

    def close(self):
        self.engine.close()


class GpuPool(object):
    """Drop-in for the `pool` argument of Inference.run_mcmc / emcee.EnsembleSampler (inference.py:1322-1356):
    `map(func, iterable)` recognises the module-level trampolines by identity and evaluates the whole iterable
    in one device call; anything else is mapped serially."""

    def __init__(self, inf: Inference):
        self.inf = inf

    def map(self, func, iterable):
        xs = [np.asarray(x, dtype=np.float64) for x in iterable]
        if not xs:
            return []
        # emcee never hands over the bare function: EnsembleSampler wraps it in emcee.ensemble._function_wrapper (2.x) /
        # _FunctionWrapper (3.x), an object with .f, .args, .kwargs whose __call__ is f(x, *args, **kwargs)
        inner = func
        if not hasattr(func, "__name__") and callable(getattr(func, "f", None)) \
                and not getattr(func, "args", None) and not getattr(func, "kwargs", None):
            inner = func.f
        name = getattr(inner, "__name__", None)
        mod = getattr(inner, "__module__", "") or ""
        if mod.endswith("inference") and name in ("post_clone", "logl", "logp"):
            X = np.stack(xs)
            if name == "post_clone":
                return list(self.inf.log_posterior_batch(X))
            if name == "logp":
                return [self.inf.logprior(x) for x in X]
            return list(self.inf.loglike_batch(X))  # logl = -inf_bound_like_obj = +loglike
        return [func(x) for x in xs]

    def close(self):
        pass

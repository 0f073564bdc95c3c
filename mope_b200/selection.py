"""Selection model: the reference's `Inference(selection_model=True)` likelihood with every
branch update on the GPU.

Mirrors mope/inference.py:770-820,896-922 (parameter unpacking, per-position scaled selection coefficients, DFE term),
mope/likelihoods.py:404-525 (`get_log_l_and_asc_prob_selection`: three rows per locus through the tree, ascertainment from
the root's child branches), mope/dfe.py (`CygnusDistribution`) and the selection branch of `_get_likelihood_limits`
(inference.py:190-247).  The per-row matrix fetches and mat-vecs of mope/_likes.pyx:59-79,155-175 run in
`mope_b200_sel_loglike` (csrc/selection.cuh); the masked-matrix operators of the posterior-mutation-location code
(_likes.pyx:82-136,178-229) are `masked_branch_update`.  The table set is a selection grid: TransitionTables whose `us`
axis holds selection coefficients s (TransitionData(selection=True), transition_data_2d.py:52-55).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.special import expit

from . import _lib
from .engine import OutOfTableError, status_message


class CygnusDistribution(object):
    """mope/dfe.py:17-120: point mass at zero + exponential tails for negative / positive coefficients."""

    def __init__(self):
        self.lower_limits = np.array([-6, -6, -4, -6])
        self.upper_limits = -self.lower_limits
        self.upper_limits[-1] = 0
        self.param_names = ["logitprobzero", "logitprobpos", "log10ratiopostoneg", "log10focalmeanalphaneg"]
        self.nparams = self.lower_limits.shape[0]

    def get_loglike(self, dfe_params, alphas):
        logitprobzero, logitprobpos, log10ratiopostoneg, log10focalmeanalphaneg = dfe_params[:4]
        prob_zero = expit(logitprobzero)
        prob_nonzero = 1 - prob_zero
        log_prob_zero = np.log(prob_zero)
        log_prob_nonzero = np.log(prob_nonzero)
        focal_mean_alpha_neg = 10 ** log10focalmeanalphaneg
        log_prob_pos = np.log(expit(logitprobpos))
        log_prob_neg = np.log(1 - np.exp(log_prob_pos))
        focal_mean_alpha_pos = 10 ** log10ratiopostoneg * focal_mean_alpha_neg
        alphas = np.asarray(alphas, dtype=np.float64)
        logprobs = np.zeros_like(alphas)
        logprobs[alphas == 0.0] = log_prob_zero
        pos = alphas > 0
        logprobs[pos] = log_prob_nonzero + log_prob_pos + np.log(focal_mean_alpha_pos) - focal_mean_alpha_pos * alphas[pos]
        neg = alphas < 0
        logprobs[neg] = log_prob_nonzero + log_prob_neg + np.log(focal_mean_alpha_neg) - focal_mean_alpha_neg * alphas[neg]
        return np.sum(logprobs)


def init_selection(inf):
    """The selection part of Inference.__init__ (inference.py:486-498, :594-668, :703-713)."""
    for p in inf.peds:
        if p.position is None:
            raise ValueError('The selection model requires a "position" column in the data')
    uniq = sorted(set().union(*[set(p.position.tolist()) for p in inf.peds]))
    inf.unique_positions = np.array(uniq)
    inf.num_unique_positions = len(uniq)
    index_of = {pos: i for i, pos in enumerate(uniq)}
    inf.position_idx = [np.array([index_of[x] for x in p.position.tolist()], dtype=np.int32) for p in inf.peds]
    inf.focal_branch = inf.varnames[0]
    inf.focal_branch_idx = 0
    inf.focal_branch_varpar_idx = inf.num_varnames
    inf.abs_sel_zero_value = 10
    # DFE labels (inference.py:612-668): one CygnusDistribution per distinct label, a position keeps a single label
    labels = set()
    have = [p.dfe is not None for p in inf.peds]   # rows of fixed loci are already gone, as in the reference (:456-459)
    if any(have):
        if not all(have):
            raise ValueError("dfe must be specified in all dataframes if it is specified in any.")
        pos_dfes = {}
        for p in inf.peds:
            for pos, d in zip(p.position.tolist(), p.dfe.tolist()):
                if pos_dfes.setdefault(pos, d) != d:
                    raise ValueError("each position must correspond to just a single DFE")
                labels.add(d)
        inf.pos_dfes = pos_dfes
    else:
        labels.add(0)
        inf.pos_dfes = {pos: 0 for pos in uniq}
    inf.num_dfes = len(labels)
    inf.dfe_translation = {el: i for i, el in enumerate(sorted(labels))}
    inf.dfes = [CygnusDistribution() for _ in range(inf.num_dfes)]
    inf.total_num_dfe_params = sum(d.nparams for d in inf.dfes)
    inf.dfe_indices = np.array([inf.pos_dfes[pos] for pos in inf.unique_positions])
    # header (inference.py:137-152)
    dfe_names = []
    for i in range(inf.num_dfes):
        dfe_names.extend([el + "_{}".format(i) for el in inf.dfes[i].param_names])
    inf.header = "\t".join(["ll"] + [v + "_l" for v in inf.varnames] + [v + "_a" for v in inf.varnames] + dfe_names
                           + ["log10ab", "log10polyprob"] + ["alpha_{}".format(bp) for bp in inf.unique_positions])
    inf.lower, inf.upper, inf.ndim = selection_limits(inf)
    inf.num_nonsel_params = inf.ndim - inf.num_unique_positions
    inf.max_alpha = inf.transition_data.get_max_mutsel_rate()


def selection_limits(inf):
    """inference.py:154-247, selection branch, as written (the upper limit of the per-position coefficients is formed from
    the MINIMUM table rate, :238-239)."""
    nv = inf.num_varnames
    td = inf.transition_data
    min_allowed_len = 0.1 * max(td.get_min_coal_time(), inf.lower_drift_limit)
    max_allowed_len = min(td.get_max_coal_time(), inf.upper_drift_limit)
    lower_len = min_allowed_len / inf.min_mults
    upper_len = max_allowed_len / inf.max_mults
    isb = np.array([inf.is_bottleneck[v] for v in inf.varnames], dtype=bool)
    inf.is_bottleneck_arr = isb
    lower_len[isb] = 2
    upper_len[isb] = 500
    lower_len, upper_len = np.log10(lower_len), np.log10(upper_len)
    dfe_lowers, dfe_uppers = [], []
    for dfe in inf.dfes:
        dfe_lowers.extend(list(dfe.lower_limits))
        dfe_uppers.extend(list(dfe.upper_limits))
    lower_sel = np.array([-2] * nv + dfe_lowers)
    upper_sel = np.array([2] * nv + dfe_uppers)
    min_mutsel_rate = 10 ** -2 * td.get_min_mutsel_rate()
    min_sel_param = min_mutsel_rate - inf.abs_sel_zero_value
    max_sel_param = min_mutsel_rate + inf.abs_sel_zero_value
    lower_alpha = np.repeat(min_sel_param, inf.num_unique_positions)
    upper_alpha = np.repeat(max_sel_param, inf.num_unique_positions)
    lower = np.concatenate((lower_len, lower_sel, (-9, -9), lower_alpha))
    upper = np.concatenate((upper_len, upper_sel, (0, 0), upper_alpha))
    ndim = 2 * nv + 2 + sum(d.nparams for d in inf.dfes) + inf.num_unique_positions
    return lower, upper, ndim


def unpack_selection_params(inf, orig_params):
    """inference.py:774-813 -> (varparams, sel_params, dfe_params, locus_alpha_values [n_var, n_pos]).  Like the reference
    it writes 0 into the focal branch's relative-size entry of the caller's vector."""
    orig_params[inf.focal_branch_varpar_idx] = 0.0
    npos = inf.num_unique_positions
    varparams = orig_params[:-npos]
    sel_params = orig_params[-npos:].copy()
    zero_filt = np.abs(sel_params) < inf.abs_sel_zero_value
    sel_params[zero_filt] = 0
    neg_filt = sel_params < 0
    sel_params[(~zero_filt) & neg_filt] += inf.abs_sel_zero_value
    pos_filt = sel_params >= 0
    sel_params[(~zero_filt) & pos_filt] -= inf.abs_sel_zero_value
    nvn = inf.num_varnames
    dfe_params = varparams[2 * nvn:2 * nvn + inf.total_num_dfe_params]
    popsizes_varpar = varparams[nvn:2 * nvn]
    popsizes_varpar[inf.focal_branch_idx] = 0.0
    relative_eff_pop_sizes = 10 ** popsizes_varpar / 1.0
    locus_alpha_values = sel_params * relative_eff_pop_sizes[:, np.newaxis]
    return varparams, sel_params, dfe_params, locus_alpha_values


def get_log_l_and_asc_prob_selection(inf, branch_lengths, locus_alpha_values, stationary_distribution, tree_idx):
    """likelihoods.get_log_l_and_asc_prob_selection (likelihoods.py:404-525) on the device -> (loglikes, log_asc_probs),
    one entry per locus in the order of the data file."""
    eng = inf.engine
    bl = np.ascontiguousarray(branch_lengths, dtype=np.float64)
    la = np.ascontiguousarray(locus_alpha_values, dtype=np.float64)
    stat = np.ascontiguousarray(stationary_distribution, dtype=np.float64)
    pidx = np.ascontiguousarray(inf.position_idx[tree_idx], dtype=np.int32)
    L = int(pidx.shape[0])
    if la.shape != (bl.shape[0], inf.num_unique_positions) or bl.shape[0] != inf.num_branches[tree_idx]:
        raise ValueError("locus_alpha_values must be [n_branches, n_unique_positions]")
    ll, la_out = np.empty(L), np.empty(L)
    status = C.c_int32(0)
    _lib.check(eng.lib.mope_b200_sel_loglike(eng.ctx, int(tree_idx), _lib.dptr(bl), _lib.dptr(la), int(la.shape[1]), _lib.iptr(pidx),
                                             _lib.dptr(stat), eng._stream(), _lib.dptr(ll), _lib.dptr(la_out), C.byref(status)),
               "mope_b200_sel_loglike")
    if status.value:
        raise OutOfTableError(status_message(int(status.value)))   # the reference's ValueError from the matrix fetch
    order = inf.peds[tree_idx].row_order   # device rows -> rows of the data file (after removal of fixed loci)
    ll_f, la_f = np.empty(L), np.empty(L)
    ll_f[order] = ll
    la_f[order] = la_out
    return ll_f, la_f


def get_dfe_loglikes(inf, all_dfe_params, all_sel_params):
    """inference.py:913-922."""
    total, cur = 0.0, 0
    for i, dfe in enumerate(inf.dfes):
        total += dfe.get_loglike(all_dfe_params[cur:cur + dfe.nparams], all_sel_params[inf.dfe_indices == i])
        cur += dfe.nparams
    return total


def loglike(inf, orig_params):
    """Inference.loglike, selection branch (inference.py:770-909)."""
    from .stationary import stationary_distribution_batch
    orig_params = np.asarray(orig_params, dtype=np.float64)
    if orig_params.shape != (inf.ndim,):
        raise ValueError("expected a parameter vector of length {}".format(inf.ndim))
    varparams, sel_params, dfe_params, locus_alpha_values = unpack_selection_params(inf, orig_params)
    if np.any(locus_alpha_values > inf.max_alpha):
        return -np.inf
    alphabeta, polyprob = 10 ** varparams[-2:]
    stat_dist = stationary_distribution_batch(inf.breaks, inf.transition_N, np.array([alphabeta]), np.array([polyprob]))[0]
    ll = 0.0
    for i in range(len(inf.peds)):
        trans_idxs = inf.translate_indices[i]
        nb = inf.num_branches[i]
        params = varparams[trans_idxs[:nb]]
        branch_lengths = 10 ** params
        locus_alpha_values_p = locus_alpha_values[trans_idxs[:nb], :]
        ped_ll, ped_log_asc_prob = get_log_l_and_asc_prob_selection(inf, branch_lengths, locus_alpha_values_p, stat_dist, i)
        ll += (ped_ll - ped_log_asc_prob).sum()
    ll += get_dfe_loglikes(inf, dfe_params, sel_params)
    return float(ll)


MASK_NONE, MASK_ZERO, MASK_FOCAL_NODE, MASK_DESCENDANT = 0, 1, 2, 3
MASK_BOTTLENECK = 4


def masked_branch_update(engine, mask: int, child, parent, times, alphas, want_products: bool = False, bottleneck: bool = False):
    """parent[r] *= Pmask_r . child[r] in place (mope_b200_sel_branch_update).  times: scalar (fixed-length versions) or
    [n_rows] (rate versions); alphas: scalar or [n_rows].  mask 0 = the selection operators, 1 / 2 / 3 = the `zero`,
    `focal node zero` and `descendant` masks of the posterior-mutation-location operators; bottleneck: the matrix comes
    from the bottleneck tables, `times` is the bottleneck size (_likes.pyx:251-307)."""
    if bottleneck:
        mask = int(mask) | MASK_BOTTLENECK
    child = np.ascontiguousarray(child, dtype=np.float64)
    if not (isinstance(parent, np.ndarray) and parent.dtype == np.float64 and parent.flags.c_contiguous):
        raise ValueError("parent must be a C-contiguous float64 array (it is updated in place)")
    n = child.shape[0]
    t = np.ascontiguousarray(np.atleast_1d(times), dtype=np.float64)
    per_row = 1 if t.shape[0] == n and np.ndim(times) > 0 else 0
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(alphas, dtype=np.float64), (n,)))
    prod = np.empty_like(child) if want_products else None
    status = C.c_int32(0)
    _lib.check(engine.lib.mope_b200_sel_branch_update(engine.ctx, int(mask), _lib.dptr(child), _lib.dptr(parent), int(n), _lib.dptr(t), per_row,
                                                      _lib.dptr(a), _lib.dptr(prod) if prod is not None else None, engine._stream(),
                                                      C.byref(status)), "mope_b200_sel_branch_update")
    if status.value:
        raise OutOfTableError(status_message(int(status.value)))
    return prod

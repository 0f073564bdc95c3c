"""Posterior mutation location: the likelihood part of mope/posteriormutloc.py with every
branch update on the GPU.

`fill_cond_probs` (posteriormutloc.py:48-133): per-locus log-likelihoods of the pruning pass and, optionally, the same
with the root distribution restricted to the two fixed states ("mutation only").
`fill_cond_probs_mutation` (posteriormutloc.py:136-271): per-locus log-likelihood that the root was fixed for the
ancestral allele and the heteroplasmy arose by mutation on the branch above a focal node -- leaves below the focal node
keep only their polymorphic bins, the others only the fixed ones, and the transition matrices are masked accordingly
(`Pzero` of _likes.pyx:82-136, 178-229, 251-307: `zero` beside the focal node, `focal node` on its own branch, `descendant`
below it).

The branch updates are the C-ABI operators mope_b200_branch_update / mope_b200_sel_branch_update; the tree walk is the
reference's postorder, one operator call per branch (this is an offline analysis, not the sampler's hot loop).  Results
are per locus in the order of the data file after removal of fixed loci, like the reference's.
"""
from __future__ import annotations

import numpy as np

from . import selection as _sel
from .stationary import stationary_distribution_batch


def _unpack(inf, tree_idx, varparams):
    varparams = np.asarray(varparams, dtype=np.float64)
    trans_idxs = inf.translate_indices[tree_idx]
    params = varparams[trans_idxs]
    nb = inf.num_branches[tree_idx]
    branch_lengths = 10 ** params[:nb]
    mutation_rates = 10 ** params[nb:]
    alphabeta, polyprob = 10 ** varparams[-2:]
    stat = stationary_distribution_batch(inf.breaks, inf.transition_N, np.array([alphabeta]), np.array([polyprob]))[0]
    return branch_lengths, mutation_rates, stat


def _leaf_likes(inf, tree_idx):
    """[L, S, F] binomial leaf likelihoods of the pedigree (device rows), computed once and cached on the Inference."""
    cache = inf.__dict__.setdefault("_pml_leaf_likes", {})
    if tree_idx not in cache:
        ped = inf.peds[tree_idx]
        phred = -1.0 if inf.min_phred_score is None else float(inf.min_phred_score)
        cache[tree_idx] = inf.engine.leaf_likelihoods(ped.counts, ped.coverage, inf.freqs, phred)
    return cache[tree_idx]


def _descendants(ped, focal):
    """branch indices strictly below branch `focal` (postorder numbering: children precede their parent)."""
    below = set()
    for b in range(focal - 1, -1, -1):
        if ped.br_parent[b] == focal or ped.br_parent[b] in below:
            below.add(b)
    return below


def _walk(inf, tree_idx, varparams, focal=None):
    ped = inf.peds[tree_idx]
    eng = inf.engine
    branch_lengths, mutation_rates, stat = _unpack(inf, tree_idx, varparams)
    L, F, B = ped.num_loci, inf.num_freqs, len(ped.branch_names)
    likes = _leaf_likes(inf, tree_idx)
    cond = [np.ones((L, F)) for _ in range(B + 1)]          # reset_likes_ones; node B is the root
    below = _descendants(ped, focal) if focal is not None else set()
    for b in range(B):
        leaf, mc, bott = int(ped.br_leaf[b]), int(ped.br_mult[b]), bool(ped.br_bottleneck[b])
        if leaf >= 0:
            child = np.ascontiguousarray(likes[:, leaf, :]).copy()
            if focal is not None:
                if b in below:
                    child[:, [0, -1]] = 0.0       # descendants of the focal node: polymorphic bins only
                else:
                    child[:, 1:-1] = 0.0          # everybody else: fixed bins only
        else:
            child = cond[b]
        parent = cond[int(ped.br_parent[b])]
        mut = float(mutation_rates[b])
        if mc >= 0:
            times = ped.row_mult[mc] * branch_lengths[b]
        else:
            times = float(branch_lengths[b])
        if focal is None:
            eng.branch_update(1 if mc >= 0 else (2 if bott else 0), child, parent, times, mut)
        else:
            mask = _sel.MASK_DESCENDANT if b in below else (_sel.MASK_FOCAL_NODE if b == focal else _sel.MASK_ZERO)
            _sel.masked_branch_update(eng, mask, child, parent, times, mut, bottleneck=(bott and mc < 0))
    return cond[B], stat, ped.row_order


def _to_file_order(values, order):
    out = np.empty_like(values)
    out[order] = values
    return out


def fill_cond_probs(inf, tree_idx, varparams, include_mutation_only=False):
    """posteriormutloc.fill_cond_probs -> overall_loglikes [L] (, mut_only_loglikes [L])."""
    root, stat, order = _walk(inf, tree_idx, varparams)
    with np.errstate(divide="ignore"):
        overall = _to_file_order(np.log(np.dot(root, stat)), order)
        if not include_mutation_only:
            return overall
        stat = stat.copy()
        stat[1:-1] = 0
        return overall, _to_file_order(np.log(np.dot(root, stat)), order)


def fill_cond_probs_mutation(focalnode, inf, tree_idx, varparams):
    """posteriormutloc.fill_cond_probs_mutation; `focalnode` is a node name (or an object with a `.name`)."""
    name = getattr(focalnode, "name", focalnode)
    ped = inf.peds[tree_idx]
    if name not in ped.branch_names:
        raise ValueError("focal node {!r} is not a non-root node of tree {}".format(name, tree_idx))
    root, stat, order = _walk(inf, tree_idx, varparams, focal=ped.branch_names.index(name))
    with np.errstate(divide="ignore"):
        return _to_file_order(np.log(root[:, 0] * stat[0]), order)

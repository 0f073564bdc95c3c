"""CPU oracle, posterior-mutation-location likelihoods (TEST INFRASTRUCTURE; same rules as oracle/mope_oracle.py).

NumPy restatement of mope/posteriormutloc.py:48-133 (`fill_cond_probs`) and :136-271 (`fill_cond_probs_mutation`) on the
oracle's own tree / table / operator classes.  Pinned against outputs of the reference itself
(tests/golden/make_golden_pml.py -> tests/golden/outputs_pml.npz; tests/test_posteriormutloc.py)."""
from __future__ import annotations

import numpy as np

from . import mope_oracle as orc
from . import mope_oracle_selection as osel


def _setup(I, k, x):
    ped = I.peds[k]
    params = x[I.translate_indices[k]]
    nb = len(ped.branch_names)
    bl, mr = 10 ** params[:nb], 10 ** params[nb:]
    ab, pp = 10 ** x[-2:]
    stat = orc.get_stationary_distribution_double_beta(I.breaks, I.transition_N, ab, pp)
    for nd in ped.tree.postorder():
        nd.cond = np.ones((ped.num_loci, I.num_freqs))
    return ped, bl, mr, stat


def fill_cond_probs(I, k, x, include_mutation_only=False):
    ped, bl, mr, stat = _setup(I, k, x)
    tree = ped.tree
    for nd in tree.postorder():
        if nd is tree:
            break
        bi = ped.branch_index[nd.name]
        child = ped.leaf_likes[nd.name].copy() if nd.is_leaf else nd.cond
        if nd.multipliername is not None:
            orc.compute_rate_transition_likelihood(child, nd.parent.cond, ped.row_mult[nd.multipliername] * bl[bi], mr[bi], I.transition_data)
        elif nd.is_bottleneck:
            orc.compute_length_transition_likelihood(child, nd.parent.cond, bl[bi], mr[bi], I.bottleneck_data)
        else:
            orc.compute_length_transition_likelihood(child, nd.parent.cond, bl[bi], mr[bi], I.transition_data)
    with np.errstate(divide="ignore"):
        overall = np.log(np.dot(tree.cond, stat))
        if not include_mutation_only:
            return overall
        s = stat.copy()
        s[1:-1] = 0
        return overall, np.log(np.dot(tree.cond, s))


def fill_cond_probs_mutation(focal_name, I, k, x):
    ped, bl, mr, stat = _setup(I, k, x)
    tree = ped.tree
    focal = [nd for nd in tree.postorder() if nd.name == focal_name][0]
    below = {nd.name for nd in focal.postorder() if nd is not focal}
    for nd in tree.postorder():
        if nd is tree:
            break
        bi = ped.branch_index[nd.name]
        if nd.is_leaf:
            child = ped.leaf_likes[nd.name].copy()
            if nd.name in below:
                child[:, [0, -1]] = 0
            else:
                child[:, 1:-1] = 0
        else:
            child = nd.cond
        mask = 3 if nd.name in below else (2 if nd is focal else 1)
        if nd.multipliername is not None:
            osel.compute_masked_transition_likelihood(child, nd.parent.cond, ped.row_mult[nd.multipliername] * bl[bi], mr[bi], I.transition_data, mask)
        elif nd.is_bottleneck:
            osel.compute_masked_transition_likelihood(child, nd.parent.cond, bl[bi], mr[bi], I.bottleneck_data, mask)
        else:
            osel.compute_masked_transition_likelihood(child, nd.parent.cond, bl[bi], mr[bi], I.transition_data, mask)
    with np.errstate(divide="ignore"):
        return np.log(tree.cond[:, 0] * stat[0])

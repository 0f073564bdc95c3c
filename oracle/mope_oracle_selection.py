"""CPU oracle, selection model and masked-matrix operators.

TEST INFRASTRUCTURE (see oracle/mope_oracle.py: the same rules apply -- only tests/, smoke() and bench.py's CPU legs
may import this; the product path never does).

NumPy restatement of the reference's selection-model likelihood (paths relative to /root/reference):
mope/inference.py:770-820,896-922 (parameter unpacking, DFE term), mope/likelihoods.py:404-525
(get_log_l_and_asc_prob_selection), mope/_likes.pyx:59-79,155-175 (per-row matrices), :82-136,178-229,251-307 (masked
"zero / focal node / descendant" operators), mope/dfe.py:17-110.  Pinned against outputs of the reference itself
(tests/golden/make_golden_selection.py -> tests/golden/*_sel.*; tests/test_selection.py checks this module against them)
and, like mope/test_selection.py does for the reference, against a hand recomputation with np.dot.
"""
from __future__ import annotations

import numpy as np
from scipy.special import expit

from . import mope_oracle as orc


class CygnusDistribution:
    """dfe.py:17-110."""

    def __init__(self):
        self.lower_limits = np.array([-6, -6, -4, -6])
        self.upper_limits = np.array([6, 6, 4, 0])
        self.param_names = ["logitprobzero", "logitprobpos", "log10ratiopostoneg", "log10focalmeanalphaneg"]
        self.nparams = 4

    def get_loglike(self, q, alphas):
        p0 = expit(q[0])
        lp_pos = np.log(expit(q[1]))
        lp_neg = np.log(1 - np.exp(lp_pos))
        mean_neg = 10 ** q[3]
        mean_pos = 10 ** q[2] * mean_neg
        a = np.asarray(alphas, dtype=np.float64)
        out = np.zeros_like(a)
        out[a == 0.0] = np.log(p0)
        out[a > 0] = np.log(1 - p0) + lp_pos + np.log(mean_pos) - mean_pos * a[a > 0]
        out[a < 0] = np.log(1 - p0) + lp_neg + np.log(mean_neg) - mean_neg * a[a < 0]
        return np.sum(out)


def compute_transition_likelihood_selection(child, parent, lengths, alphas, transitions):
    """_likes.pyx:59-79 (lengths [n]) and :155-175 (scalar length): one matrix per row; returns the products."""
    n = child.shape[0]
    lengths = np.broadcast_to(np.asarray(lengths, dtype=np.float64), (n,))
    ret = np.zeros_like(parent)
    for i in range(n):
        P = transitions.get_transition_probabilities_2d(lengths[i], alphas[i])
        tmp = np.dot(P, child[i])
        parent[i] *= tmp
        ret[i] = tmp
    return ret


def masked(P, mask):
    """Pzero of _likes.pyx:82-136: 1 = [0,0]; 2 = [0,1:-1]; 3 = [1:-1,1:-1]."""
    Z = np.zeros_like(P)
    if mask == 1:
        Z[0, 0] = P[0, 0]
    elif mask == 2:
        Z[0, 1:-1] = P[0, 1:-1]
    elif mask == 3:
        Z[1:-1, 1:-1] = P[1:-1, 1:-1]
    else:
        Z[:] = P
    return Z


def compute_masked_transition_likelihood(child, parent, lengths, rate, transitions, mask):
    """_likes.pyx:82-136 (lengths [n]), :178-229 and :251-307 (scalar length / bottleneck size)."""
    n = child.shape[0]
    lengths = np.broadcast_to(np.asarray(lengths, dtype=np.float64), (n,))
    for i in range(n):
        P = transitions.get_transition_probabilities_2d(lengths[i], rate)
        parent[i] *= np.dot(masked(P, mask), child[i])


class SelectionInference(orc.Inference):
    """mope.inference.Inference(selection_model=True)."""

    def __init__(self, pedigrees_in, drift, **kw):
        # inference.py:341-342: rows sorted by (family, position) (pandas sorts several keys with a stable lexsort)
        srt = []
        for data, t, a in pedigrees_in:
            order = np.lexsort((np.asarray(data["position"]), np.asarray(data["family"])))
            srt.append(({k: np.asarray(v)[order] for k, v in data.items()}, t, a))
        pedigrees_in = srt
        super().__init__(pedigrees_in, drift, None, **kw)
        self.positions, self.dfe_labels = [], []
        for (data, _t, _a), ped in zip(pedigrees_in, self.peds):
            X = np.column_stack([np.asarray(data[l], dtype=np.float64) for l in ped.leaf_names])
            n = np.column_stack([np.asarray(data[l + "_n"], dtype=np.float64) for l in ped.leaf_names])
            fsum = np.nansum(X / n, axis=1)
            keep = ~(np.isclose(fsum, X.shape[1]) | (fsum == 0.0))
            self.positions.append(np.asarray(data["position"])[keep])
            self.dfe_labels.append(np.asarray(data["dfe"])[keep] if "dfe" in data else None)
        # inference.py:486-498
        uniq = sorted(set().union(*[set(p.tolist()) for p in self.positions]))
        self.unique_positions = np.array(uniq)
        self.num_unique_positions = len(uniq)
        self.position_idx = [np.array([uniq.index(x) for x in p.tolist()]) for p in self.positions]
        self.abs_sel_zero_value = 10
        # inference.py:612-668
        labels = set()
        if any(d is not None for d in self.dfe_labels):
            pos_dfes = {}
            for p, d in zip(self.positions, self.dfe_labels):
                for pos, lab in zip(p.tolist(), d.tolist()):
                    pos_dfes.setdefault(pos, lab)
                    labels.add(lab)
        else:
            labels.add(0)
            pos_dfes = {pos: 0 for pos in uniq}
        self.dfes = [CygnusDistribution() for _ in labels]
        self.total_num_dfe_params = 4 * len(self.dfes)
        self.dfe_indices = np.array([pos_dfes[pos] for pos in uniq])
        self.lower, self.upper, self.ndim = self._selection_limits()
        self.max_alpha = self.transition_data.max_mutsel_rate

    # inference.py:154-247, selection branch
    def _selection_limits(self):
        nv = self.num_varnames
        lower_n, upper_n, _ = self._limits()
        dl = [v for d in self.dfes for v in d.lower_limits]
        du = [v for d in self.dfes for v in d.upper_limits]
        m = 10 ** -2 * self.transition_data.min_mutsel_rate
        npos = self.num_unique_positions
        lower = np.concatenate((lower_n[:nv], [-2] * nv + dl, (-9, -9), np.repeat(m - self.abs_sel_zero_value, npos)))
        upper = np.concatenate((upper_n[:nv], [2] * nv + du, (0, 0), np.repeat(m + self.abs_sel_zero_value, npos)))
        return lower, upper, 2 * nv + 2 + self.total_num_dfe_params + npos

    def unpack(self, x):
        """inference.py:774-813; writes 0 into x[num_varnames] like the reference."""
        nv, npos = self.num_varnames, self.num_unique_positions
        x[nv] = 0.0
        varparams = x[:-npos]
        sel = x[-npos:].copy()
        zero = np.abs(sel) < self.abs_sel_zero_value
        sel[zero] = 0
        neg = sel < 0
        sel[(~zero) & neg] += self.abs_sel_zero_value
        pos = sel >= 0
        sel[(~zero) & pos] -= self.abs_sel_zero_value
        dfe_params = varparams[2 * nv:2 * nv + self.total_num_dfe_params]
        lav = sel * (10 ** varparams[nv:2 * nv])[:, None]
        return varparams, sel, dfe_params, lav

    def ped_selection(self, k, branch_lengths, lav, stat):
        """likelihoods.py:404-525 -> (loglikes [L], log_asc_probs [L])."""
        ped = self.peds[k]
        L, F = ped.num_loci, self.num_freqs
        e0 = (self.freqs <= self.min_freq).astype(np.float64)
        eN = (self.freqs >= 1 - self.min_freq).astype(np.float64)
        eboth = np.vstack((e0,) * L + (eN,) * L)
        pidx = self.position_idx[k]
        tree = ped.tree
        for nd in tree.postorder():
            nd.cond = np.ones((3 * L, F))
        asc = np.ones((L, F))
        for nd in tree.postorder():
            if nd is tree:
                break
            bi = ped.branch_index[nd.name]
            alphas = np.tile(lav[bi, :][pidx], 3)
            child = np.concatenate((ped.leaf_likes[nd.name], eboth)) if nd.is_leaf else nd.cond
            if nd.multipliername is not None:
                lengths = np.tile(ped.row_mult[nd.multipliername] * branch_lengths[bi], 3)
            else:
                lengths = branch_lengths[bi]
            trans = compute_transition_likelihood_selection(child, nd.parent.cond, lengths, alphas, self.transition_data)
            if nd.parent is tree:
                asc *= 1.0 - trans[L:2 * L] - trans[2 * L:]
        with np.errstate(divide="ignore", invalid="ignore"):
            loglikes = np.log(np.dot(tree.cond[:L], stat))
            log_asc = np.log(np.dot(asc, stat))
        return loglikes, log_asc

    def get_dfe_loglikes(self, dfe_params, sel):
        tot, cur = 0.0, 0
        for i, d in enumerate(self.dfes):
            tot += d.get_loglike(dfe_params[cur:cur + 4], sel[self.dfe_indices == i])
            cur += 4
        return tot

    def loglike(self, x, components=None):
        """inference.py:770-909, selection branch."""
        varparams, sel, dfe_params, lav = self.unpack(x)
        if np.any(lav > self.max_alpha):
            return -np.inf
        ab, pp = 10 ** varparams[-2:]
        stat = orc.get_stationary_distribution_double_beta(self.breaks, self.transition_N, ab, pp)
        ll = 0.0
        for k, ped in enumerate(self.peds):
            tr = self.translate_indices[k]
            nb = len(ped.branch_names)
            bl = 10 ** varparams[tr[:nb]]
            a, b = self.ped_selection(k, bl, lav[tr[:nb], :], stat)
            if components is not None:
                components.append((a, b))
            ll += (a - b).sum()
        return ll + self.get_dfe_loglikes(dfe_params, sel)

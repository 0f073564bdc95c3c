"""Host model compiler: data + tree + ages -> flat arrays for the C ABI.

Mirrors the init-time logic of the reference's Inference.__init__ (mope/inference.py:331-352 file reading,
:385-451 branch/variable index maps, :456-459 removal of fixed loci, :504-525 family counts and multiplier
columns) and flattens the result into the mope_pedigree_desc layout of include/mope_b200.h.
"""
from __future__ import annotations

import io
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
from scipy.special import gammaln

from . import newick


def read_table(path_or_text, is_text: bool = False) -> pd.DataFrame:
    """TSV reader with the reference's options (inference.py:332-333, :510)."""
    if isinstance(path_or_text, pd.DataFrame):   # already parsed (synthetic data sets are built as tables)
        return path_or_text
    src = io.StringIO(path_or_text) if is_text else path_or_text
    return pd.read_csv(src, sep="\t", comment="#", na_values=["NA"], header=0)


def is_fixed(X: np.ndarray, n: np.ndarray) -> np.ndarray:
    """mope/data.py:219-228 for count data."""
    with np.errstate(divide="ignore", invalid="ignore"):
        ps = X / n
    freqsums = np.nansum(ps, axis=1).astype(np.float64)
    return np.isclose(freqsums, X.shape[1]) | (freqsums == 0.0)


@dataclass
class PedigreeModel:
    tree: newick.Node
    leaf_names: List[str]
    branch_names: List[str]
    varname_of: Dict[str, str]
    multiplier_of: Dict[str, Optional[str]]
    bottleneck_of_var: Dict[str, bool]
    multipliernames: List[str]
    # data after init-time processing
    counts: np.ndarray        # [L, S]
    coverage: np.ndarray      # [L, S]
    family: np.ndarray        # [L]
    ages: pd.DataFrame        # the ages table with a 'count' column (asc_ages in the reference)
    row_mult: np.ndarray      # [M, L]
    asc_mult: np.ndarray      # [M, U]
    fam_asc: np.ndarray       # [n_fam] int32
    fam_count: np.ndarray     # [n_fam] float64
    # schedule
    br_parent: np.ndarray
    br_leaf: np.ndarray
    br_mult: np.ndarray
    br_bottleneck: np.ndarray
    position: Optional[np.ndarray] = None    # [L] base-pair position of every locus row (selection model), row order of `counts`
    row_order: Optional[np.ndarray] = None   # [L] index of every row in the data file after removal of fixed loci
    dfe: Optional[np.ndarray] = None         # [L] DFE label of every locus row (selection model, optional "dfe" column)

    @property
    def num_loci(self) -> int:
        return int(self.counts.shape[0])

    @property
    def n_fam(self) -> int:
        return int(self.fam_asc.shape[0])

    @property
    def n_asc(self) -> int:
        return int(self.asc_mult.shape[1])


def compile_pedigree(data: pd.DataFrame, tree_str: str, ages: pd.DataFrame, sort_rows: bool = True) -> PedigreeModel:
    tree = newick.loads(tree_str)[0]
    nodes = [nd for nd in tree.postorder() if nd is not tree]
    leaf_names = [nd.name for nd in nodes if nd.is_leaf]
    branch_names = [nd.name for nd in nodes]
    if len(set(branch_names)) != len(branch_names):
        raise ValueError("node names must be unique in each tree")
    for nd in nodes:
        if nd.varname is None:
            raise ValueError("every non-root node needs a length parameter name (node {})".format(nd.name))
    varname_of = {nd.name: nd.varname for nd in nodes}
    multiplier_of = {nd.name: nd.multipliername for nd in nodes}
    bott = {}
    for nd in nodes:
        if nd.varname in bott and bott[nd.varname] != nd.is_bottleneck:
            raise ValueError("nodes with the same variable name must have the same status as bottlenecks in all trees")
        bott[nd.varname] = nd.is_bottleneck
    multipliernames = sorted({m for m in multiplier_of.values() if m is not None})

    data = data.copy()
    for col in data.columns:
        if col not in ("family", "position", "dfe"):
            data[col] = data[col].astype(np.float64)
    missing = [c for c in leaf_names + [l + "_n" for l in leaf_names] if c not in data.columns]
    if missing:
        raise ValueError("data file lacks columns for tree leaves: {}".format(missing))
    X = data.loc[:, leaf_names].values.astype(np.float64)
    n = data.loc[:, [l + "_n" for l in leaf_names]].values.astype(np.float64)
    keep = ~is_fixed(X, n)
    data = data.loc[keep, :].reset_index(drop=True)
    X, n = np.ascontiguousarray(X[keep]), np.ascontiguousarray(n[keep])
    family = data["family"].values
    position = data["position"].values if "position" in data.columns else None
    dfe = data["dfe"].values if "dfe" in data.columns else None
    row_order = np.arange(X.shape[0])

    ages = ages.copy()
    fam_ids = ages["family"].values
    # loci per family (ascertainment.py:173 counts rows per family id); one pass instead of one scan per family
    counts = pd.Series(family).value_counts().reindex(fam_ids, fill_value=0).values.astype(np.int64)
    ages["count"] = counts
    for m in multipliernames:
        if m not in ages.columns:
            raise ValueError("multiplier column {} not found in the ages file".format(m))
    M, L = len(multipliernames), X.shape[0]
    row_mult = np.zeros((M, L))
    first_row_of = {}
    for i, f in enumerate(fam_ids):
        first_row_of.setdefault(f, i)
    for j, m in enumerate(multipliernames):
        if m in data.columns:
            row_mult[j] = data[m].values.astype(np.float64)
        else:
            col = ages[m].values.astype(np.float64)
            first_idx = np.array([first_row_of[f] for f in pd.unique(family)], dtype=np.int64)
            row_mult[j] = col[first_idx[pd.factorize(family)[0]]]   # .iloc[0] of the matching rows
    # distinct ascertainment row pairs: the pruning of the pseudo-rows depends on the family only through its
    # multiplier values, so equal multiplier tuples share one pair (bit-identical results, fewer rows)
    fam_mult = np.zeros((M, len(fam_ids)))
    for j, m in enumerate(multipliernames):
        fam_mult[j] = ages[m].values.astype(np.float64)
    if M == 0:
        asc_mult = np.zeros((0, 1))
        fam_asc = np.zeros(len(fam_ids), dtype=np.int32)
    else:
        keys, fam_asc_l, uniq = {}, [], []
        for i in range(len(fam_ids)):
            k = fam_mult[:, i].tobytes()
            if k not in keys:
                keys[k] = len(uniq)
                uniq.append(fam_mult[:, i])
            fam_asc_l.append(keys[k])
        asc_mult = np.ascontiguousarray(np.array(uniq).T) if uniq else np.zeros((M, 0))
        fam_asc = np.array(fam_asc_l, dtype=np.int32)

    if M > 0 and sort_rows and len(fam_ids) > 1:
        # Age-scaled branches blend the two generation-grid matrices around N*age*rate, so a 64-row tile only needs
        # the grid cells its rows fall into.  Order the families along a k-d tree of their (log) multiplier values:
        # tiles become narrow in EVERY multiplier column at once.  Row order does not change any result beyond the
        # association order of the final sum over loci.
        rank_of_fam_pos = np.empty(len(fam_ids), dtype=np.int64)
        rank_of_fam_pos[_kd_order(np.log(np.maximum(fam_mult.T, 1e-300)), leaf_size=8)] = np.arange(len(fam_ids))
        pos_of_fam = {}
        for i, f in enumerate(fam_ids):
            pos_of_fam.setdefault(f, i)
        locus_rank = np.array([rank_of_fam_pos[pos_of_fam[f]] if f in pos_of_fam else len(fam_ids) for f in family])
        order = np.argsort(locus_rank, kind="stable")
        X, n, family = np.ascontiguousarray(X[order]), np.ascontiguousarray(n[order]), family[order]
        row_mult = np.ascontiguousarray(row_mult[:, order])
        row_order = row_order[order]
        position = position[order] if position is not None else None
        dfe = dfe[order] if dfe is not None else None
        # distinct ascertainment pairs in the same order
        uniq_rank = np.full(asc_mult.shape[1], len(fam_ids), dtype=np.int64)
        np.minimum.at(uniq_rank, fam_asc, rank_of_fam_pos)
        uorder = np.argsort(uniq_rank, kind="stable")
        new_index = np.empty_like(uorder)
        new_index[uorder] = np.arange(len(uorder))
        asc_mult = np.ascontiguousarray(asc_mult[:, uorder])
        fam_asc = new_index[fam_asc].astype(np.int32)

    if sort_rows and L > 8:
        # A binomial emission is an exact 0.0 far from the observed frequency, and the pruning kernels skip the K chunks in
        # which the 8 rows a warp owns are all zero (emission_mask_kernel), so rows with similar allele frequencies go next
        # to each other; rows with a missing tissue (emission all ones for that leaf) are grouped by the tissue they miss.
        # Without age-scaled branches the row order is free (global sort); with them the rows stay in the k-d order above
        # (a tile's rows decide the generation segments of its GEMMs) inside a few frequency classes / inside large blocks.
        # Only the association order of the final sum over loci changes.
        with np.errstate(invalid="ignore", divide="ignore"):
            fr = X / n
        miss = np.isnan(fr)
        all_miss = miss.all(axis=1)
        mean_fr = np.where(all_miss, 0.0, np.nanmean(np.where(all_miss[:, None], 0.0, fr), axis=1))
        any_miss = miss.any(axis=1)
        first_miss = np.where(any_miss, miss.argmax(axis=1), -1)
        if M == 0:
            order = np.lexsort((mean_fr, first_miss, any_miss))
        else:
            # Measured on B200 (profiles/r02_row_order.md): one multiplier column (the k-d order is a plain sort by age, whole
            # runs of tiles share their generation segments): frequency order inside blocks of 32768 rows, cfg4 4228 -> 3868 ms
            # per step; several columns (small k-d cells): two frequency classes, each in k-d order, cfg3 499 -> 483 ms.
            classes = int(os.environ.get("MOPE_FREQ_CLASSES", "2" if (M > 1 and L >= 8192) else "1"))   # small data: keep the k-d cells
            block = int(os.environ.get("MOPE_FREQ_BLOCK", "32768" if M == 1 else "0"))
            order = None
            if classes > 1 or block > 0:
                cls = np.zeros(L, dtype=np.int64)
                if classes > 1:
                    cls = np.searchsorted(np.quantile(mean_fr, np.arange(1, classes) / classes), mean_fr)
                o1 = np.lexsort((np.arange(L), cls))
                order = o1
                if block > 0:
                    o2 = np.lexsort((mean_fr[o1], first_miss[o1], any_miss[o1], np.arange(L) // block))
                    order = o1[o2]
        if order is not None:
            X, n, family = np.ascontiguousarray(X[order]), np.ascontiguousarray(n[order]), family[order]
            row_mult = np.ascontiguousarray(row_mult[:, order])
            row_order = row_order[order]
            position = position[order] if position is not None else None
            dfe = dfe[order] if dfe is not None else None

    idx_of = {id(nd): i for i, nd in enumerate(nodes)}
    idx_of[id(tree)] = len(nodes)
    br_parent = np.array([idx_of[id(nd.parent)] for nd in nodes], dtype=np.int32)
    br_leaf = np.array([leaf_names.index(nd.name) if nd.is_leaf else -1 for nd in nodes], dtype=np.int32)
    br_mult = np.array([multipliernames.index(nd.multipliername) if nd.multipliername is not None else -1 for nd in nodes],
                       dtype=np.int32)
    br_bott = np.array([1 if nd.is_bottleneck else 0 for nd in nodes], dtype=np.int32)
    return PedigreeModel(tree=tree, leaf_names=leaf_names, branch_names=branch_names, varname_of=varname_of,
                         multiplier_of=multiplier_of, bottleneck_of_var=bott, multipliernames=multipliernames,
                         counts=X, coverage=n, family=family, ages=ages, row_mult=np.ascontiguousarray(row_mult),
                         asc_mult=asc_mult, fam_asc=fam_asc, fam_count=counts.astype(np.float64),
                         br_parent=br_parent, br_leaf=br_leaf, br_mult=br_mult, br_bottleneck=br_bott, position=position,
                         row_order=row_order, dfe=dfe)


def _kd_order(P: np.ndarray, leaf_size: int = 8) -> np.ndarray:
    """Indices of the rows of P [n, d] in k-d tree order (recursive median split along the widest coordinate)."""
    out = []
    stack = [np.arange(P.shape[0])]
    while stack:
        idx = stack.pop()
        if idx.shape[0] <= leaf_size:
            out.append(idx)
            continue
        sub = P[idx]
        col = int(np.argmax(sub.max(axis=0) - sub.min(axis=0)))
        o = np.argsort(sub[:, col], kind="stable")
        half = idx.shape[0] // 2
        # LIFO: push the upper half first so that the lower half comes out first
        stack.append(idx[o[half:]])
        stack.append(idx[o[:half]])
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


def fam_lgamma(counts: np.ndarray) -> np.ndarray:
    """gammaln(count + 1) of inference.py:1126."""
    return np.ascontiguousarray(gammaln(np.asarray(counts, dtype=np.float64) + 1), dtype=np.float64)

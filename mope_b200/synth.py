"""Synthetic heteroplasmy datasets of the BASELINE.json shapes.

Counts follow the reference simulator's sampling scheme (mope/simulate.py:240-262: coverage ~ Poisson(mean),
alt count ~ Binomial(coverage, leaf frequency)); leaf frequencies are a heteroplasmic root frequency plus
per-tissue noise rather than a forward Wright-Fisher walk (the likelihood cost does not depend on it).
Everything is written in the reference's own TSV / Newick formats, so the same text feeds both
implementations.
"""
from __future__ import annotations

import io
from typing import List

import numpy as np


class SyntheticPedigree:
    """One synthetic pedigree type: the data and ages tables (as the DataFrames the reference's pd.read_csv would give)
    and the tree.  `data_tsv` / `ages_tsv` render the reference's TSV text on first use (the CPU reference arm and the
    parity tests read files; a million-locus table is never rendered)."""

    def __init__(self, data, tree: str, ages, leaf_names: List[str], n_loci: int = None, n_fam: int = None):
        import pandas as pd
        from . import model
        self._data_tsv = data if isinstance(data, str) else None
        self._ages_tsv = ages if isinstance(ages, str) else None
        self.data = model.read_table(data, True) if isinstance(data, str) else data
        self.ages = model.read_table(ages, True) if isinstance(ages, str) else ages
        assert isinstance(self.data, pd.DataFrame) and isinstance(self.ages, pd.DataFrame)
        self.tree = tree
        self.leaf_names = list(leaf_names)
        self.n_loci = int(self.data.shape[0] if n_loci is None else n_loci)
        self.n_fam = int(self.ages.shape[0] if n_fam is None else n_fam)

    @property
    def data_tsv(self) -> str:
        if self._data_tsv is None:
            leaves = self.leaf_names
            buf = io.StringIO()
            cols = list(self.data.columns)
            buf.write("\t".join(cols) + "\n")
            x = self.data[leaves].values
            cov = self.data[[l + "_n" for l in leaves]].values
            fam = self.data["family"].values
            for i in range(x.shape[0]):
                row = [("NA" if np.isnan(x[i, j]) else "%d" % x[i, j]) for j in range(len(leaves))]
                row += ["%d" % cov[i, j] for j in range(len(leaves))]
                row.append("%d" % fam[i])
                buf.write("\t".join(row) + "\n")
            self._data_tsv = buf.getvalue()
        return self._data_tsv

    @property
    def ages_tsv(self) -> str:
        if self._ages_tsv is None:
            a = self.ages
            buf = io.StringIO()
            buf.write("family\tmother_birth_age\tchild_age\tmother_age\n")
            fam, mba, ca, ma = a["family"].values, a["mother_birth_age"].values, a["child_age"].values, a["mother_age"].values
            for k in range(a.shape[0]):
                buf.write("%d\t%.4f\t%.4f\t%.4f\n" % (fam[k], mba[k], ca[k], ma[k]))
            self._ages_tsv = buf.getvalue()
        return self._ages_tsv

    def subset(self, n_loci: int) -> "SyntheticPedigree":
        """The first n_loci loci and the families they belong to, same tree."""
        n_loci = min(int(n_loci), self.data.shape[0])
        d = self.data.iloc[:n_loci].reset_index(drop=True)
        fams = set(d["family"].values.tolist())
        a = self.ages[self.ages["family"].isin(fams)].reset_index(drop=True)
        return SyntheticPedigree(d, self.tree, a, self.leaf_names)


def balanced_tree(n_leaves: int = 16, rate_leaves: bool = False, rate_internal: bool = False, bottleneck: bool = False) -> str:
    """Balanced binary tree.  Branch parameter names are shared by depth and side (d<depth><a|b>), i.e. 8 names
    for 16 leaves.  rate_leaves: leaf branches become `r<a|b>*mother_age` / `*child_age` alternating;
    rate_internal: one depth-1 branch becomes `loo*mother_birth_age`; bottleneck: the other depth-1 branch is `eoo^`."""
    depth = int(round(np.log2(n_leaves)))
    assert 2 ** depth == n_leaves
    counter = [0]

    def build(d, side):
        # returns newick text of the subtree whose root is at depth d
        if d == depth:
            i = counter[0]
            counter[0] += 1
            name = "t%d" % i
            if rate_leaves:
                return "%s:r%s*%s" % (name, side, "mother_age" if i % 2 == 0 else "child_age")
            return "%s:d%d%s" % (name, d, side)
        kids = build(d + 1, "a") + "," + build(d + 1, "b")
        name = "n%d_%d" % (d, counter[0])
        if d == 0:
            return "(%s)root" % kids
        var = "d%d%s" % (d, side)
        if d == 1 and side == "a" and rate_internal:
            var = "loo*mother_birth_age"
        if d == 1 and side == "b" and bottleneck:
            var = "eoo^"
        return "(%s)%s:%s" % (kids, name, var)

    return build(0, "") + ";"


def make_dataset(n_loci: int, n_leaves: int = 16, tree: str = None, loci_per_family: int = 3, mean_coverage: float = 1000.0,
                 missing_frac: float = 0.02, seed: int = 20261019) -> SyntheticPedigree:
    rng = np.random.RandomState(seed)
    tree = tree or balanced_tree(n_leaves)
    leaves = ["t%d" % i for i in range(n_leaves)]
    n_fam = (n_loci + loci_per_family - 1) // loci_per_family
    # ages: normals of examples/params/mope_map.params truncated at 0.4 (simulate.py:98-99)
    mba = np.maximum(rng.normal(30.30, 5.11, n_fam), 0.4)
    ca = np.maximum(rng.normal(10.16, 5.0, n_fam), 0.4)
    ma = ca + mba
    fam_of_locus = np.arange(n_loci) // loci_per_family
    # heteroplasmic root frequency, U-shaped like the stationary distribution but kept off the boundaries so
    # that no locus is dropped as fixed
    root = rng.beta(0.3, 0.3, n_loci) * 0.96 + 0.02
    f = np.clip(root[:, None] + rng.normal(0.0, 0.05, (n_loci, n_leaves)), 0.0, 1.0)
    cov = np.maximum(rng.poisson(mean_coverage, (n_loci, n_leaves)), 1)
    x = rng.binomial(cov, f).astype(np.float64)
    # every locus keeps at least one informative tissue (otherwise the reference drops it as fixed)
    dead = (x.sum(1) == 0) | np.all(x == cov, axis=1)
    x[dead, 0] = np.floor(cov[dead, 0] / 2)
    miss = rng.rand(n_loci, n_leaves) < missing_frac
    miss[:, 0] = False
    import pandas as pd
    xs = np.where(miss, np.nan, x)
    data = pd.DataFrame(np.concatenate((xs, cov.astype(np.float64)), axis=1), columns=leaves + [l + "_n" for l in leaves])
    data["family"] = fam_of_locus.astype(np.int64)

    def as_written(v):   # the value a reader of the TSV gets back ("%.4f")
        return np.char.mod("%.4f", v).astype(np.float64)
    ages = pd.DataFrame({"family": np.arange(n_fam, dtype=np.int64), "mother_birth_age": as_written(mba),
                         "child_age": as_written(ca), "mother_age": as_written(ma)})
    return SyntheticPedigree(data, tree, ages, leaves, n_loci, n_fam)


def walkers_in_prior_box(lower: np.ndarray, upper: np.ndarray, n_walkers: int, seed: int = 1, shrink: float = 0.0) -> np.ndarray:
    """Uniform draws inside the prior box (optionally shrunk towards its centre by `shrink` on each side)."""
    rng = np.random.RandomState(seed)
    lo = lower + shrink * (upper - lower)
    up = upper - shrink * (upper - lower)
    return rng.uniform(lo, up, (n_walkers, lower.shape[0]))

#!/usr/bin/env python
"""Second set of golden fixtures, again produced by RUNNING THE REFERENCE (ammodramus/mope): the grid sizes and
shapes the benchmark uses, on tables that reach the end of the reference's default generation grid.

    python tests/golden/make_golden_wide.py        (build container only: needs /root/reference)

Writes
  tables_f115.npz    reference generator, WF N=1000, --breaks 0.5 0.01 (F=115, the published default binning);
                     generations 1..10000 (row sums of the stored matrices exceed 1 the most at the long end, which is what
                     the clamp / renormalisation of _util.check_transition_matrix acts on), 3 rates; 4 bottleneck sizes
  tables_f201.npz    the same at --breaks 0.5 0.005 (F=201, the benchmark's grid size); 3 generations x 2 rates; 2 x 2 bottleneck
  cases_wide.json    inputs per case (synthetic data sets in the reference's TSV / Newick formats)
  outputs_wide.npz   reference outputs per case (as make_golden.py) + the binomial emission tensor at coverage ~1000 on
                     the F=201 grid
Cases: the reference's example data on the long grid; a cfg3-shaped data set (age-scaled leaf branches, 200 families with
distinct ages); a cfg4-shaped one (bottleneck branch, ascertainment, 200 families); F=201 versions of cfg2 / cfg3 shapes.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import install_ref  # noqa: E402
import make_golden as mg  # noqa: E402

warnings.filterwarnings("ignore")

GRIDS = {
    "f115": dict(breaks=(0.5, 0.01), gens=[1, 10, 100, 1000, 5000, 10000], us=[1e-8, 1e-5, 1e-3], nbs=[2, 20, 200, 500]),
    "f201": dict(breaks=(0.5, 0.005), gens=[1, 500, 10000], us=[1e-6, 1e-3], nbs=[2, 500]),
}


def tables(tag, workdir):
    g = GRIDS[tag]
    mg.BREAKS, mg.GENS, mg.US, mg.NBS = g["breaks"], g["gens"], g["us"], g["nbs"]
    sub = os.path.join(workdir, tag)
    os.makedirs(sub)
    saved = os.path.join(mg.HERE, "tables_small.npz")
    keep = saved + ".keep"
    os.rename(saved, keep)      # generate_tables writes tables_small.npz; park the real one
    try:
        dm, bm = mg.generate_tables(sub)
        os.rename(saved, os.path.join(HERE, "tables_%s.npz" % tag))
    finally:
        os.rename(keep, saved)
    return dm, bm


def main():
    install_ref.install()
    assert install_ref.activate()
    from mope_b200 import synth
    examples = os.path.join(install_ref.DEST, "examples")
    base = mg.build_cases(examples)
    out, cases = {}, {}
    with tempfile.TemporaryDirectory() as wd:
        masters = {tag: tables(tag, wd) for tag in GRIDS}

        def synth_case(n_loci, n_leaves, seed, **tree_kw):
            ds = synth.make_dataset(n_loci, n_leaves, tree=synth.balanced_tree(n_leaves, **tree_kw), seed=seed)
            return dict(peds=[dict(data=ds.data_tsv, tree=ds.tree, ages=ds.ages_tsv)], opts={})

        cases["f115/example_both"] = dict(peds=base["example_both"]["peds"], nvec=6, opts={})
        cases["f115/cfg3_shaped"] = dict(nvec=3, **synth_case(600, 8, 21, rate_leaves=True, rate_internal=True))
        cases["f115/cfg4_shaped"] = dict(nvec=3, **synth_case(600, 8, 22, rate_internal=True, bottleneck=True))
        cases["f201/cfg2_shaped"] = dict(nvec=3, **synth_case(150, 16, 23))
        cases["f201/cfg3_shaped"] = dict(nvec=2, **synth_case(150, 8, 24, rate_leaves=True, rate_internal=True))
        first = {}
        for name, case in cases.items():
            tag = name.split("/")[0]
            case["tables"] = "tables_%s.npz" % tag
            mg.GENS, mg.US, mg.NBS = GRIDS[tag]["gens"], GRIDS[tag]["us"], GRIDS[tag]["nbs"]
            I = mg.run_case(name.replace("/", "_"), case, masters[tag][0], masters[tag][1], wd, out)
            first.setdefault(tag, I)
        # binomial emissions at the benchmark's coverage (n ~ Poisson(1000)) on the F = 201 grid
        from mope import _binom
        rng = np.random.RandomState(31)
        n = rng.poisson(1000, (24, 4)).astype(np.float64)
        f = np.clip(rng.beta(0.3, 0.3, (24, 1)) * 0.96 + 0.02 + rng.normal(0, 0.05, (24, 4)), 0, 1)
        x = rng.binomial(n.astype(np.int64), f).astype(np.float64)
        x[3, 1] = np.nan
        x[4, 2] = 0
        x[5, 0] = n[5, 0]
        I201 = first["f201"]
        out["unit/f201_binom_X"], out["unit/f201_binom_n"] = x, n
        out["unit/f201_binom_like"] = _binom.get_binom_likelihoods_cython(x, n, I201.freqs, -1.0)
        out["unit/f201_binom_like_phred30"] = _binom.get_binom_likelihoods_cython(x, n, I201.freqs, 30.0)
        # matrices on the long end of the grid (row sums > 1 -> renormalisation)
        for tag, I in first.items():
            td = I.transition_data
            N = I.transition_N
            ts = [9999.5 / N, 7000.0 / N, 10000.0 / N, 2500.0 / N]
            ms = [2 * N * GRIDS[tag]["us"][-1] * 0.7, 2 * N * GRIDS[tag]["us"][0] * 1.5, 2 * N * GRIDS[tag]["us"][-1], 2 * N * GRIDS[tag]["us"][1]]
            P = []
            for t, m in zip(ts, ms):
                td.clear_cache()
                P.append(np.array(td.get_transition_probabilities_2d(t, m)))
            out["unit/%s_drift_t" % tag], out["unit/%s_drift_m" % tag], out["unit/%s_drift_P" % tag] = np.array(ts), np.array(ms), np.array(P)
    with open(os.path.join(HERE, "cases_wide.json"), "w") as fjs:
        json.dump(cases, fjs, indent=1)
    np.savez_compressed(os.path.join(HERE, "outputs_wide.npz"), **out)
    for tag in GRIDS:
        z = np.load(os.path.join(HERE, "tables_%s.npz" % tag))
        rs = z["drift"].sum(-1)
        print(tag, "F", z["freqs"].shape[0], "drift row sums in [%.3e, 1 + %.3e]" % (rs.min(), rs.max() - 1), "negatives", int((z["drift"] < 0).sum()))
    print("wrote", HERE)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden fixtures of the SELECTION model and of the masked-matrix operators, produced by RUNNING THE REFERENCE
(ammodramus/mope) in the build container:

    python tests/golden/make_golden_selection.py        (needs /root/reference)

The reference's own tests (mope/test_selection.py) read `selection_transitions.h5`, a download that is not in the repository;
the table set is therefore generated here by the reference's own generator (generate_transitions._run_generate, the recipe
`mope generate-commands --selection` prints: u = v = 0, one file per selection coefficient), on a grid that -- unlike the
printed recipe -- holds negative coefficients and 0, which Inference.loglike needs (|x| < 10 means alpha = 0, and
TransitionData raises below the smallest tabulated coefficient).

Writes
  tables_sel.npz         WF N=1000, --breaks 0.5 0.02 (F = 65); 7 generations x 7 selection coefficients (`us` holds s)
  cases_sel.json         inputs per case: the four data sets of mope/test_selection.py (one locus, one locus twice, two loci in
                         two families on the age-scaled tree, the same with two DFEs) + the example data with synthetic positions
                         on a six-branch tree + a two-pedigree case
  outputs_sel.npz        per case: X, Inference.loglike(X) (nan where the reference raises ValueError), per pedigree the
                         per-locus loglikes / log_asc_probs of likelihoods.get_log_l_and_asc_prob_selection, DFE term, prior box,
                         header; operator-level vectors of the selection and masked ("zero", "focal node", "descendant")
                         branch updates of _likes.pyx, drift and bottleneck tables
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import warnings
from argparse import Namespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import install_ref  # noqa: E402

warnings.filterwarnings("ignore")

N_WF = 1000
BREAKS = (0.5, 0.02)
GENS = [1, 5, 20, 100, 500, 1500, 3000]
SS = [-5e-3, -1e-3, -1e-4, 0.0, 1e-4, 1e-3, 5e-3]

SIX_BRANCH_TREE = ("((mother_blood:blo*mother_age,mother_cheek:buc*mother_age)somM:som,"
                   "((child_blood:blo*child_age,child_cheek:buc*child_age)som1:som)loo:loo*mother_birth_age)emb;")


def write_master(path, mats, axis_name, axis, xs_name, xs, N, freqs, breaks):
    import h5py  # the pickle-backed stand-in of oracle/ref_shims
    with h5py.File(path, "w") as mf:
        k = 0
        for i, g in enumerate(axis):
            for j, x in enumerate(xs):
                ds = mf.create_dataset("P%d" % k, data=mats[i, j])
                ds.attrs.update({"N": N, axis_name: g, xs_name: x})
                if xs_name == "s":
                    ds.attrs.update({"u": 0.0, "v": 0.0})
                k += 1
        mf.attrs["frequencies"] = freqs
        mf.attrs["breaks"] = breaks
        mf.attrs["N"] = N
    return path


def generate_selection_tables(workdir):
    from mope import generate_transitions as gt
    import h5py
    gens_fn = os.path.join(workdir, "gens.txt")
    with open(gens_fn, "w") as f:
        f.write("\n".join(str(g) for g in GENS) + "\n")
    drift = None
    for j, s in enumerate(SS):
        fn = os.path.join(workdir, "sel_%d.h5" % j)
        gt._run_generate(Namespace(N=N_WF, s=s, u=0.0, v=0.0, inputfile=gens_fn, output=fn, bottlenecks=False, breaks=BREAKS,
                                   debug=False, log=False))
        with h5py.File(fn, "r") as f:
            freqs, breaks = np.array(f.attrs["frequencies"]), np.array(f.attrs["breaks"])
            if drift is None:
                drift = np.zeros((len(GENS), len(SS), freqs.shape[0], freqs.shape[0]))
            for name in f:
                ds = f[name]
                assert float(ds.attrs["s"]) == s
                drift[GENS.index(int(ds.attrs["gen"])), j] = ds[:, :]
    np.savez_compressed(os.path.join(HERE, "tables_sel.npz"), drift=drift, gens=np.array(GENS, dtype=np.float64), us=np.array(SS),
                        freqs=freqs, breaks=breaks, N=np.int64(N_WF))
    return drift, freqs, breaks


def read(path):
    with open(path) as f:
        return f.read()


def build_cases(examples):
    d, t = os.path.join(examples, "data"), os.path.join(examples, "trees")
    one = read(os.path.join(d, "data_test_allele_counts_with_position.txt"))
    one_ages = read(os.path.join(d, "data_test_allele_counts_with_position_ages.txt"))
    twice = read(os.path.join(d, "data_test_allele_counts_with_position_twice.txt"))
    two = read(os.path.join(d, "data_test_allele_counts_with_position_two_loci.txt"))
    two_ages = read(os.path.join(d, "data_test_allele_counts_with_position_two_loci_ages.txt"))
    two_dfes = read(os.path.join(d, "data_test_allele_counts_with_position_two_loci_with_dfes.txt"))
    simple = read(os.path.join(t, "test_mother_child_simple_tree.newick"))
    rate = read(os.path.join(t, "test_mother_child_rate_tree.newick"))
    # the example data with a synthetic position column: 40 positions, shared between families
    ex = read(os.path.join(d, "data_example_allele_counts.txt"))
    ex_ages = read(os.path.join(d, "data_example_allele_counts_ages.txt"))
    rng = np.random.RandomState(5)
    lines = [l for l in ex.splitlines() if not l.startswith("#")]
    out = [lines[0] + "\tposition"]
    for l in lines[1:]:
        out.append(l + "\t%d" % (16 * rng.randint(40) + 73))
    ex_pos = "\n".join(out) + "\n"
    ref_test = dict(min_freq=0.0, min_phred_score=1000)    # the options of mope/test_selection.py
    return {
        "sel_one_locus": dict(peds=[dict(data=one, tree=simple, ages=one_ages)], nvec=8, opts=ref_test),
        "sel_one_locus_twice": dict(peds=[dict(data=twice, tree=simple, ages=one_ages)], nvec=6, opts=ref_test),
        "sel_two_loci": dict(peds=[dict(data=two, tree=rate, ages=two_ages)], nvec=8, opts=ref_test),
        "sel_two_loci_dfes": dict(peds=[dict(data=two_dfes, tree=rate, ages=two_ages)], nvec=8, opts=ref_test),
        "sel_example": dict(peds=[dict(data=ex_pos, tree=SIX_BRANCH_TREE, ages=ex_ages)], nvec=6, opts=dict(min_freq=0.001)),
        "sel_two_pedigrees": dict(peds=[dict(data=ex_pos, tree=SIX_BRANCH_TREE, ages=ex_ages), dict(data=two, tree=rate, ages=two_ages)],
                                  nvec=4, opts=dict(min_freq=0.002, min_phred_score=30.0)),
    }


def draw_vectors(I, n, rng):
    """Parameter vectors that reach non-zero coefficients of both signs (the prior box itself maps almost every coefficient
    to zero: inference.py:232-241 forms both limits from the smallest tabulated rate)."""
    nv, npos = I.num_varnames, I.num_unique_positions
    lo, up = I.lower.copy(), I.upper.copy()
    X = rng.uniform(0.7 * lo + 0.3 * up, 0.3 * lo + 0.7 * up, (n, I.ndim))
    X[:, nv:2 * nv] = rng.uniform(-0.3, 0.3, (n, nv))          # relative population sizes 0.5 .. 2
    kind = rng.randint(3, size=(n, npos))
    mag = rng.uniform(0.05, 4.0, (n, npos))
    sel = np.where(kind == 0, rng.uniform(-9.5, 9.5, (n, npos)), np.where(kind == 1, 10.0 + mag, -10.0 - mag))
    X[:, -npos:] = sel
    X[0, -npos:] = 3.0                                         # all zero
    return X


def run_case(name, case, master, workdir, out):
    from mope import inference as refinf
    from mope import likelihoods as li
    dfs, tfs, afs = [], [], []
    for k, ped in enumerate(case["peds"]):
        for key, lst in (("data", dfs), ("tree", tfs), ("ages", afs)):
            fn = os.path.join(workdir, "%s_%d_%s.txt" % (name, k, key))
            with open(fn, "w") as f:
                f.write(ped[key])
            lst.append(fn)
    opts = dict(genome_size=16569, poisson_like_penalty=1.0, min_freq=0.001, min_phred_score=None)
    opts.update(case["opts"])
    I = refinf.Inference(data_files=dfs, tree_files=tfs, age_files=afs, transitions_file=master, true_parameters=None,
                         start_from="prior", data_are_freqs=False, bottleneck_file=None, log_unif_drift=True,
                         lower_drift_limit=1e-3, upper_drift_limit=3, selection_model=True, **opts)
    rng = np.random.RandomState(13)
    X = draw_vectors(I, case["nvec"], rng)
    npos = I.num_unique_positions
    # one vector over the largest tabulated rate (-inf, inference.py:816-820), one below the smallest (ValueError from the fetch)
    xa, xb = X[1].copy(), X[2].copy()
    xa[-npos] = 10.0 + 2.0 * I.max_alpha
    xb[-1] = -10.0 - 1.5 * abs(I.transition_data.get_min_mutsel_rate())
    X = np.vstack([X, xa, xb])
    n_ped = len(case["peds"])
    ll = np.zeros(X.shape[0])
    dfe = np.zeros(X.shape[0])
    per_ll = [np.full((X.shape[0], I.num_loci[k]), np.nan) for k in range(n_ped)]
    per_asc = [np.full((X.shape[0], I.num_loci[k]), np.nan) for k in range(n_ped)]
    X_after = X.copy()
    nvn = I.num_varnames
    for i in range(X.shape[0]):
        x = X[i].copy()
        I.transition_data.clear_cache()
        try:
            ll[i] = I.loglike(x)
        except ValueError as e:
            ll[i] = np.nan
            print("  ", name, i, "ValueError:", e)
        X_after[i] = x                                   # loglike writes 0 into the focal branch's population size
        if not np.isfinite(ll[i]):
            continue
        nvn = I.num_varnames
        varparams = x[:-npos]
        sel = x[-npos:].copy()
        zero = np.abs(sel) < I.abs_sel_zero_value
        sel[zero] = 0
        sel[(~zero) & (sel < 0)] += I.abs_sel_zero_value
        sel[(~zero) & (sel >= 0)] -= I.abs_sel_zero_value   # as inference.py:787-795 (the second filter is formed after the first update)
        lav = sel * (10 ** varparams[nvn:2 * nvn])[:, None]
        dfe[i] = I.get_dfe_loglikes(varparams[2 * nvn:2 * nvn + I.total_num_dfe_params], sel)
        ab, pp = 10 ** varparams[-2:]
        stat = li.get_stationary_distribution_double_beta(I.freqs, I.breaks, I.transition_N, ab, pp)
        tot = dfe[i]
        for k in range(n_ped):
            nb = I.num_branches[k]
            tr = I.translate_indices[k]
            bl = 10 ** varparams[tr][:nb]
            a, b = li.get_log_l_and_asc_prob_selection(bl, lav[tr[:nb], :], stat, I, k, I.min_freq)
            per_ll[k][i], per_asc[k][i] = a, b
            tot += (a - b).sum()
        assert abs(tot - ll[i]) <= 1e-9 * abs(ll[i]), (tot, ll[i])
    # log_posterior (sign folding as written, inference.py:1140-1145: it folds the relative population sizes and the LAST
    # two entries -- per-position coefficients in this layout) on vectors inside the prior box and one outside
    Xp = rng.uniform(I.lower, I.upper, (3, I.ndim))
    Xp[:, nvn:2 * nvn] *= rng.choice([-1.0, 1.0], (3, nvn))
    Xp[2, 0] = I.upper[0] + 0.1
    lp = np.zeros(3)
    for i in range(3):
        I.transition_data.clear_cache()
        lp[i] = I.log_posterior(Xp[i])
    out[name + "/Xp"] = Xp
    out[name + "/log_posterior"] = lp
    out[name + "/X"] = X
    out[name + "/X_after"] = X_after
    out[name + "/loglike"] = ll
    out[name + "/dfe_ll"] = dfe
    for k in range(n_ped):
        out[name + "/loglikes_%d" % k] = per_ll[k]
        out[name + "/log_asc_%d" % k] = per_asc[k]
    out[name + "/lower"], out[name + "/upper"] = I.lower, I.upper
    case["header"] = I.header
    case["varnames"] = list(I.varnames)
    case["unique_positions"] = [int(p) for p in I.unique_positions]
    case["dfe_indices"] = [int(v) for v in I.dfe_indices]
    print(name, "ndim", I.ndim, "loglike", ll)
    return I


def unit_vectors(I, out):
    """Operator-level vectors from mope/_likes.pyx: selection operators (:59-79, :155-175) on the selection tables, the masked
    operators (:82-136, :178-229) on the same tables read as (generation x rate) grids, bottleneck twins (:251-307)."""
    from mope import _likes
    from mope import transition_data_bottleneck as tdb
    from tests import helpers as H
    td = I.transition_data
    F = I.num_freqs
    rng = np.random.RandomState(17)
    L = 9
    child = rng.uniform(0, 1, (L, F))
    par = rng.uniform(0, 1, (L, F))
    lens = rng.uniform(0.002, 2.5, L)
    lens[3] = lens[2]
    lens[5] = 0.0004                                # N t < first generation: identity
    alphas = rng.uniform(-9.0, 9.0, L)
    alphas[3] = alphas[2]
    alphas[6] = 0.0
    alphas[7] = 2 * N_WF * SS[1]                    # exact grid value
    out["unit/child"], out["unit/parent_in"], out["unit/lengths"], out["unit/alphas"] = child, par, lens, alphas
    td.clear_cache()
    p = par.copy()
    out["unit/sel_rate_ret"] = _likes.compute_rate_transition_likelihood_selection(child, p, lens, alphas, td)
    out["unit/sel_rate_out"] = p
    p = par.copy()
    out["unit/sel_length_ret"] = _likes.compute_length_transition_likelihood_selection(child, p, 0.61, alphas, td)
    out["unit/sel_length_out"] = p
    rate = 3.7
    for tag, fn in (("zero", _likes.compute_leaf_zero_transition_likelihood), ("focal", _likes.compute_leaf_focal_node_zero_transition_likelihood),
                    ("desc", _likes.compute_leaf_zero_transition_likelihood_descendant)):
        p = par.copy()
        fn(child, p, lens, rate, td)
        out["unit/rate_%s_out" % tag] = p
    for tag, fn in (("zero", _likes.compute_length_transition_likelihood_zero), ("focal", _likes.compute_length_transition_likelihood_zero_focalnode),
                    ("desc", _likes.compute_length_transition_likelihood_zero_descendant)):
        p = par.copy()
        fn(child, p, 0.61, rate, td)
        out["unit/length_%s_out" % tag] = p
    out["unit/mask_rate"] = np.float64(rate)
    # bottleneck twins on the neutral fixture's bottleneck tables (tests/golden/tables_small.npz, made by the reference generator)
    tb = H.load_tables()
    with tempfile.TemporaryDirectory() as wd:
        bm = write_master(os.path.join(wd, "bot.h5"), tb["bot"], "Nb", [int(v) for v in tb["nbs"]], "u", [float(v) for v in tb["bot_us"]],
                          int(tb["N"]), tb["freqs"], tb["breaks"])
        bd = tdb.TransitionDataBottleneck(bm, memory=True)
    Fb = tb["freqs"].shape[0]
    childb, parb = rng.uniform(0, 1, (L, Fb)), rng.uniform(0, 1, (L, Fb))
    out["unit/bot_child"], out["unit/bot_parent_in"] = childb, parb
    out["unit/bot_nb"], out["unit/bot_rate"] = np.float64(33.3), np.float64(0.004)
    for tag, fn in (("zero", _likes.compute_bottleneck_transition_likelihood_zero),
                    ("focal", _likes.compute_bottleneck_transition_likelihood_zero_focalnode),
                    ("desc", _likes.compute_bottleneck_transition_likelihood_zero_descendant)):
        p = parb.copy()
        fn(childb, p, 33.3, 0.004, bd)
        out["unit/bot_%s_out" % tag] = p


def main():
    install_ref.install()
    assert install_ref.activate()
    examples = os.path.join(install_ref.DEST, "examples")
    cases = build_cases(examples)
    out = {}
    with tempfile.TemporaryDirectory() as wd:
        drift, freqs, breaks = generate_selection_tables(wd)
        master = write_master(os.path.join(wd, "selection_transitions.h5"), drift, "gen", GENS, "s", SS, N_WF, freqs, breaks)
        first = None
        for name, case in cases.items():
            I = run_case(name, case, master, wd, out)
            first = first or I
        unit_vectors(first, out)
    with open(os.path.join(HERE, "cases_sel.json"), "w") as fjs:
        json.dump(cases, fjs, indent=1)
    np.savez_compressed(os.path.join(HERE, "outputs_sel.npz"), **out)
    print("wrote", HERE)


if __name__ == "__main__":
    main()

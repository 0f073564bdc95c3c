#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE (ammodramus/mope).

Run in the build container only (needs /root/reference; installs the patched copy into
baseline/_ref through oracle/install_ref.py):

    python tests/golden/make_golden.py

Writes
  tables_small.npz   transition tables made by the reference's own generate_transitions._run_generate
                     (WF N=1000, --breaks 0.5 0.02 -> F=65; 12 generations x 4 rates; 8 bottleneck
                     sizes x 4 rates)
  cases.json         input texts (data TSV / newick / ages TSV) and options per case
  outputs.npz        reference outputs: per case the parameter vectors and loglike / logprior /
                     log_posterior / root log-likelihood / per-family log ascertainment probabilities;
                     unit-level vectors for interpolation, check_transition_matrix, binomial emissions,
                     stationary distribution, Poisson-binomial non-ascertainment probabilities.
"""
from __future__ import annotations

import io
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
from oracle import install_ref  # noqa: E402

warnings.filterwarnings("ignore")

N_WF = 1000
BREAKS = (0.5, 0.02)
GENS = [1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 3000]
US = [1e-9, 1e-7, 1e-5, 1e-3]
NBS = [2, 5, 10, 20, 50, 100, 200, 500]


def generate_tables(workdir):
    from mope import generate_transitions as gt
    import h5py  # the shim

    gens_fn = os.path.join(workdir, "gens.txt")
    bots_fn = os.path.join(workdir, "bots.txt")
    with open(gens_fn, "w") as f:
        f.write("\n".join(str(g) for g in GENS) + "\n")
    with open(bots_fn, "w") as f:
        f.write("\n".join(str(b) for b in NBS) + "\n")
    drift_files, bot_files = [], []
    for u in US:
        fn = os.path.join(workdir, "drift_%g.h5" % u)
        gt._run_generate(Namespace(N=N_WF, s=0.0, u=u, v=u, inputfile=gens_fn, output=fn, bottlenecks=False,
                                   breaks=BREAKS, debug=False, log=False))
        drift_files.append(fn)
        fn = os.path.join(workdir, "bot_%g.h5" % u)
        gt._run_generate(Namespace(N=N_WF, s=0.0, u=u, v=u, inputfile=bots_fn, output=fn, bottlenecks=True,
                                   breaks=BREAKS, debug=False, log=False))
        bot_files.append(fn)

    # master files: the shim has no ExternalLink, so datasets are copied under D<k> names
    def master(files, out):
        with h5py.File(out, "w") as mf:
            k = 0
            for fn in files:
                with h5py.File(fn, "r") as f:
                    for name in f:
                        ds = mf.create_dataset("D%d" % k, data=f[name][:, :])
                        ds.attrs.update(f[name].attrs)
                        k += 1
                    mf.attrs["frequencies"] = f.attrs["frequencies"]
                    mf.attrs["breaks"] = f.attrs["breaks"]
                    mf.attrs["N"] = f.attrs["N"]
        return out

    drift_master = master(drift_files, os.path.join(workdir, "drift_master.h5"))
    bot_master = master(bot_files, os.path.join(workdir, "bot_master.h5"))

    # dense arrays for the fixture
    with h5py.File(drift_master, "r") as f:
        F = f.attrs["frequencies"].shape[0]
        drift = np.zeros((len(GENS), len(US), F, F))
        for name in f:
            ds = f[name]
            drift[GENS.index(int(ds.attrs["gen"])), US.index(float(ds.attrs["u"]))] = ds[:, :]
        freqs = np.array(f.attrs["frequencies"])
        breaks = np.array(f.attrs["breaks"])
    with h5py.File(bot_master, "r") as f:
        bot = np.zeros((len(NBS), len(US), F, F))
        for name in f:
            ds = f[name]
            bot[NBS.index(int(ds.attrs["Nb"])), US.index(float(ds.attrs["u"]))] = ds[:, :]
    np.savez_compressed(os.path.join(HERE, "tables_small.npz"), drift=drift, gens=np.array(GENS, dtype=np.float64),
                        us=np.array(US), bot=bot, nbs=np.array(NBS, dtype=np.int64), bot_us=np.array(US),
                        freqs=freqs, breaks=breaks, N=np.int64(N_WF))
    return drift_master, bot_master


def read(path):
    with open(path) as f:
        return f.read()


def build_cases(examples):
    d, t = os.path.join(examples, "data"), os.path.join(examples, "trees")
    ex_data = read(os.path.join(d, "data_example_allele_counts.txt"))
    ex_ages = read(os.path.join(d, "data_example_allele_counts_ages.txt"))
    two_data = read(os.path.join(d, "data_test_allele_counts_with_position_two_loci.txt"))
    two_ages = read(os.path.join(d, "data_test_allele_counts_with_position_two_loci_ages.txt"))
    one_data = read(os.path.join(d, "data_test_allele_counts_with_position.txt"))
    one_ages = read(os.path.join(d, "data_test_allele_counts_with_position_ages.txt"))

    # a variant of the example data with missing tissues (NA counts) -- our own perturbation
    rng = np.random.RandomState(11)
    lines = ex_data.splitlines()
    hdr_idx = [i for i, l in enumerate(lines) if not l.startswith("#")][0]
    cols = lines[hdr_idx].split("\t")
    leafcols = [cols.index(c) for c in ("mother_blood", "mother_cheek", "child_blood", "child_cheek")]
    out = lines[: hdr_idx + 1]
    for l in lines[hdr_idx + 1:]:
        parts = l.split("\t")
        if rng.rand() < 0.25:
            parts[leafcols[rng.randint(4)]] = "NA"
        out.append("\t".join(parts))
    na_data = "\n".join(out) + "\n"

    cases = {
        "example_both": dict(peds=[dict(data=ex_data, tree=read(os.path.join(t, "mope_study_both.newick")), ages=ex_ages)],
                             nvec=64, opts={}),
        "example_fixed": dict(peds=[dict(data=ex_data, tree=read(os.path.join(t, "mope_study_fixed.newick")), ages=ex_ages)],
                              nvec=8, opts={}),
        "example_linear": dict(peds=[dict(data=ex_data, tree=read(os.path.join(t, "mope_study_linear.newick")), ages=ex_ages)],
                               nvec=8, opts={}),
        "example_both_na_phred": dict(peds=[dict(data=na_data, tree=read(os.path.join(t, "mope_study_both.newick")), ages=ex_ages)],
                                      nvec=8, opts=dict(min_phred_score=30.0, poisson_like_penalty=0.5, min_freq=0.002,
                                                        genome_size=10000)),
        "test_simple": dict(peds=[dict(data=one_data, tree=read(os.path.join(t, "test_mother_child_simple_tree.newick")), ages=one_ages)],
                            nvec=8, opts={}),
        "test_rate_two_loci": dict(peds=[dict(data=two_data, tree=read(os.path.join(t, "test_mother_child_rate_tree.newick")), ages=two_ages)],
                                   nvec=8, opts={}),
        "two_pedigrees": dict(peds=[dict(data=ex_data, tree=read(os.path.join(t, "mope_study_both.newick")), ages=ex_ages),
                                    dict(data=two_data, tree=read(os.path.join(t, "mope_study_linear.newick")), ages=two_ages)],
                              nvec=8, opts=dict(poisson_like_penalty=0.0)),
    }
    return cases


def run_case(name, case, drift_master, bot_master, workdir, out):
    from mope import inference as refinf
    from mope import likelihoods as li
    from mope import ascertainment as asc

    dfs, tfs, afs = [], [], []
    for k, ped in enumerate(case["peds"]):
        for key, lst in (("data", dfs), ("tree", tfs), ("ages", afs)):
            fn = os.path.join(workdir, "%s_%d_%s.txt" % (name, k, key))
            with open(fn, "w") as f:
                f.write(ped[key])
            lst.append(fn)
    opts = dict(genome_size=16569, poisson_like_penalty=1.0, min_freq=0.001, min_phred_score=None)
    opts.update(case["opts"])
    I = refinf.Inference(data_files=dfs, tree_files=tfs, age_files=afs, transitions_file=drift_master,
                         true_parameters=None, start_from="prior", data_are_freqs=False,
                         bottleneck_file=bot_master, log_unif_drift=True, lower_drift_limit=1e-3,
                         upper_drift_limit=3, **opts)
    rng = np.random.RandomState(7)
    lo, up = I.lower, I.upper
    X = rng.uniform(0.6 * lo + 0.4 * up, 0.4 * lo + 0.6 * up, (case["nvec"], I.ndim))
    # wider vectors: anywhere in the central 90% of the box (drives short times -> identity short-cut etc.)
    Xw = rng.uniform(0.95 * lo + 0.05 * up, 0.05 * lo + 0.95 * up, (max(4, case["nvec"] // 4), I.ndim))
    X = np.vstack([X, Xw])
    if name == "example_both":
        # the MAP point shipped in examples/params/mope_map.params
        mp = dict(blo=(2.104e-3, 1.035e-8), buc=(2.08e-3, 1.0635e-8), fblo=(2.104e-2, 1.035e-6),
                  fbuc=(2.08e-2, 1.0635e-6), som=(4.937e-2, 1.0632e-8), loo=(1.984e-4, 1.0265e-8),
                  eoo=(3.903e2, 1.01626e-8))
        N = I.transition_N
        xm = [np.log10(mp[v][0]) for v in I.varnames] + [np.log10(2 * N * mp[v][1]) for v in I.varnames]
        xm += [np.log10(3.14697e-3), np.log10(2.5575e-3)]
        xm = np.clip(np.array(xm), lo + 1e-9, up - 1e-9)
        X = np.vstack([X, xm])
    ll = np.zeros(X.shape[0])
    lp = np.zeros(X.shape[0])
    pr = np.zeros(X.shape[0])
    n_ped = len(case["peds"])
    root_ll = np.zeros((X.shape[0], n_ped))
    log_asc = [np.zeros((X.shape[0], I.num_asc_combs[k])) for k in range(n_ped)]
    for i, x in enumerate(X):
        I.transition_data.clear_cache()
        ll[i] = I.loglike(x.copy())
        pr[i] = I.logprior(x)
        # log_posterior folds signs; feed it a vector with flipped signs on the folded coordinates
        xs = x.copy()
        nv = I.num_varnames
        xs[nv:2 * nv] *= np.where(np.arange(nv) % 2 == 0, -1.0, 1.0)
        xs[-1] *= -1.0
        lp[i] = I.log_posterior(xs)
        ab, pp = 10 ** x[-2:]
        stat = li.get_stationary_distribution_double_beta(I.freqs, I.breaks, I.transition_N, ab, pp)
        for k in range(n_ped):
            params = x[I.translate_indices[k]]
            nb = I.num_branches[k]
            bl, mr = 10 ** params[:nb], 10 ** params[nb:]
            root_ll[i, k] = li.get_log_likelihood_somatic_newick(bl, mr, stat, I, k)
            log_asc[k][i] = asc.get_locus_asc_probs(bl, mr, stat, I, I.min_freq, k)
    out[name + "/X"] = X
    out[name + "/loglike"] = ll
    out[name + "/logprior"] = pr
    out[name + "/log_posterior"] = lp
    out[name + "/root_ll"] = root_ll
    for k in range(n_ped):
        out[name + "/log_asc_%d" % k] = log_asc[k]
        out[name + "/counts_%d" % k] = np.array(I.asc_ages[k]["count"].values, dtype=np.int64)
        out[name + "/num_loci_%d" % k] = np.int64(I.num_loci[k])
    out[name + "/lower"] = I.lower
    out[name + "/upper"] = I.upper
    case["varnames"] = list(I.varnames)
    case["branch_names"] = [list(b) for b in I.branch_names]
    case["leaf_names"] = [list(b) for b in I.leaf_names]
    case["header"] = I.header
    print(name, "ndim", I.ndim, "loglike[0]", ll[0], "finite", np.isfinite(ll).sum(), "/", ll.shape[0])
    return I


def unit_vectors(I, out):
    """Operator-level vectors from the reference's own functions."""
    from mope import _binom, _util, _interp, _likes
    from mope import likelihoods as li

    td, bd = I.transition_data, I.bottleneck_data
    N = I.transition_N
    rng = np.random.RandomState(3)
    # drift interpolation incl. exact grid hits, identity short-cut, table edges
    ts = list(rng.uniform(td.get_min_coal_time(), td.get_max_coal_time(), 12))
    ts += [GENS[0] / N, GENS[3] / N, GENS[-1] / N, 0.5 * GENS[0] / N, 1e-7]
    ms = list(10 ** rng.uniform(np.log10(td.get_min_mutsel_rate()), np.log10(td.get_max_mutsel_rate()), 12))
    ms += [2 * N * US[0], 2 * N * US[1], 2 * N * US[-1], 2 * N * 3e-6, 2 * N * 5e-8]
    P = []
    for t, m in zip(ts, ms):
        td.clear_cache()
        P.append(np.array(td.get_transition_probabilities_2d(t, m)))
    out["unit/drift_t"] = np.array(ts)
    out["unit/drift_m"] = np.array(ms)
    out["unit/drift_P"] = np.array(P)
    # second fetch of a cached key re-applies the check to the cached array
    td.clear_cache()
    a = np.array(td.get_transition_probabilities_2d(ts[0], ms[0]))
    b = np.array(td.get_transition_probabilities_2d(ts[0], ms[0]))
    out["unit/drift_P_second_fetch"] = b
    # bottleneck interpolation
    nbs = list(rng.uniform(2, 500, 10)) + [2.0, 500.0, 20.0, 37.5]
    bms = list(10 ** rng.uniform(np.log10(bd.get_min_mutation_rate()), np.log10(bd.get_max_mutation_rate()), 10))
    bms += [2 * N * US[0], 2 * N * US[-1], 2 * N * US[2], 2 * N * 4e-6]
    out["unit/bot_nb"] = np.array(nbs)
    out["unit/bot_m"] = np.array(bms)
    out["unit/bot_P"] = np.array([np.array(bd.get_transition_probabilities_2d(nb, m)) for nb, m in zip(nbs, bms)])
    # check_transition_matrix on a matrix that needs clamping and renormalising
    M = rng.uniform(-0.05, 0.2, (9, 9))
    M[2] *= 0.01
    M[4, 3] = 1.7
    out["unit/check_in"] = M.copy()
    _util.check_transition_matrix(M)
    out["unit/check_out"] = M
    # raw bilinear blend
    Q = rng.uniform(0, 1, (4, 7, 7))
    w = rng.dirichlet(np.ones(4))
    out["unit/blend_Q"] = Q
    out["unit/blend_w"] = w
    out["unit/blend_out"] = _interp.bilinear_interpolate(Q[0], Q[1], Q[2], Q[3], 1.0, w[0], w[1], w[2], w[3])
    # binomial emissions: edge cases k>n, k==0, k==n, p==0 and p==1 bins, NaN, phred mixing
    Xc = np.array([[0, 5, 1000, np.nan], [17, 0, 3, 2000], [1, 999, 1200, 40], [300, 64, 0, 1]], dtype=np.float64)
    nc = np.array([[1000, 1000, 1000, 900], [1000, 0, 3, 1999], [1, 1000, 2400, 5000], [1000, 128, 7, 1]], dtype=np.float64)
    out["unit/binom_X"] = Xc
    out["unit/binom_n"] = nc
    out["unit/binom_like"] = _binom.get_binom_likelihoods_cython(Xc, nc, I.freqs, -1.0)
    out["unit/binom_like_phred20"] = _binom.get_binom_likelihoods_cython(Xc, nc, I.freqs, 20.0)
    # stationary distribution
    abpp = np.array([[3.14697e-3, 2.5575e-3], [1e-9, 1e-9], [1.0, 1.0], [1e-5, 0.3], [0.2, 1e-4], [3e-7, 0.9]])
    out["unit/stat_abpp"] = abpp
    out["unit/stat_dist"] = np.array([li.get_stationary_distribution_double_beta(I.freqs, I.breaks, N, a, p) for a, p in abpp])
    # Poisson-binomial non-ascertainment probabilities.  The reference function writes one double past
    # the end of its scratch table (P[j, j+1] at j = nreads-1 with bounds checks off,
    # _poisson_binom.pyx:84-85), which corrupts the heap and aborts at free() time; each call is
    # therefore made in a child process that saves the result and leaves through os._exit.
    import subprocess
    for k, (nr, mf) in enumerate([(40, 0.1), (200, 0.02), (1000, 0.001), (333, 0.01)]):
        q = rng.choice([20.0, 30.0, 40.0], size=nr, p=[0.1, 0.6, 0.3])
        with tempfile.TemporaryDirectory() as td_:
            np.save(os.path.join(td_, "q.npy"), q)
            np.save(os.path.join(td_, "f.npy"), np.array(I.freqs))
            code = ("import sys, os, numpy as np; sys.path[:0] = %r; from mope import _poisson_binom as pb; "
                    "r = pb.get_non_asc_probs(np.load(%r), np.load(%r), %r); np.save(%r, r); "
                    "sys.stdout.flush(); os._exit(0)") % (
                [os.path.join(install_ref.DEST, "shims"), install_ref.DEST], os.path.join(td_, "q.npy"),
                os.path.join(td_, "f.npy"), mf, os.path.join(td_, "r.npy"))
            subprocess.call([sys.executable, "-c", code])
            res = np.load(os.path.join(td_, "r.npy"))
        out["unit/pb_q_%d" % k] = q
        out["unit/pb_minfreq_%d" % k] = np.float64(mf)
        out["unit/pb_out_%d" % k] = res
    # one branch update of each kind on random vectors
    L = 6
    child = rng.uniform(0, 1, (L, I.num_freqs))
    par = rng.uniform(0, 1, (L, I.num_freqs))
    lens = rng.uniform(0.002, 1.5, L)
    lens[1] = lens[0]
    out["unit/branch_child"] = child
    out["unit/branch_parent_in"] = par
    out["unit/branch_lengths"] = lens
    td.clear_cache()
    p1 = par.copy()
    _likes.compute_length_transition_likelihood(child, p1, 0.37, 0.004, td)
    out["unit/branch_length_out"] = p1
    p2 = par.copy()
    _likes.compute_rate_transition_likelihood(child, p2, lens, 0.004, td)
    out["unit/branch_rate_out"] = p2
    p3 = par.copy()
    _likes.compute_bottleneck_transition_likelihood(child, p3, 33.3, 0.004, bd)
    out["unit/branch_bottleneck_out"] = p3
    stat = li.get_stationary_distribution_double_beta(I.freqs, I.breaks, N, 3e-3, 2e-3)
    out["unit/root_stat"] = stat
    out["unit/root_ll"] = np.float64(_likes.compute_root_log_like(child, stat))
    small = child[:4] * 1e-3
    out["unit/asc_logprobs"] = _likes.get_individual_locus_asc_logprobs(small, stat, np.array([1, 2], dtype=np.int64))


def main():
    install_ref.install()
    assert install_ref.activate()
    examples = os.path.join(install_ref.DEST, "examples")
    out = {}
    with tempfile.TemporaryDirectory() as wd:
        drift_master, bot_master = generate_tables(wd)
        cases = build_cases(examples)
        first = None
        for name, case in cases.items():
            I = run_case(name, case, drift_master, bot_master, wd, out)
            if first is None:
                first = I
        unit_vectors(first, out)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(cases, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "outputs.npz"), **out)
    print("wrote", HERE)


if __name__ == "__main__":
    main()

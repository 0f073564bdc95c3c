"""Selection model and the masked-matrix operators of the posterior-mutation-location code.

CPU part: the oracle restatement (oracle/mope_oracle_selection.py) against the reference-generated fixtures
(tests/golden/make_golden_selection.py ran mope's Inference(selection_model=True) on selection tables made by mope's own
generator) and -- the method of the reference's own mope/test_selection.py -- against a hand recomputation with np.dot.
GPU part: mope_b200.Inference(selection_model=True) and selection.masked_branch_update, through the C ABI, against the same
fixtures (<= 1e-10 relative on the log-likelihood, 1e-12 per locus, 1e-13 on operator outputs)."""
import json
import os
import warnings

import numpy as np
import pytest

from tests import helpers as H
from oracle import mope_oracle as orc
from oracle import mope_oracle_selection as osel

TB = H.load_tables("tables_sel.npz")
with open(os.path.join(H.GOLDEN, "cases_sel.json")) as _f:
    CASES = json.load(_f)
_z = np.load(os.path.join(H.GOLDEN, "outputs_sel.npz"))
OUT = {k: _z[k] for k in _z.files}
NAMES = sorted(CASES)


def _oracle(case):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        drift = orc.TransitionData(TB["drift"], TB["gens"], TB["us"], int(TB["N"]), TB["freqs"], TB["breaks"])
        peds = [(H.read_tsv(p["data"]), p["tree"], H.read_tsv(p["ages"])) for p in case["peds"]]
        opts = dict(log_unif_drift=True, min_freq=0.001, min_phred_score=None)
        opts.update(case["opts"])
        return osel.SelectionInference(peds, drift, **opts)


# ------------------------------------------------------------------------------------------------ CPU: oracle pin
@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    I = _oracle(CASES[name])
    X, ll = OUT[name + "/X"], OUT[name + "/loglike"]
    assert I.ndim == X.shape[1]
    np.testing.assert_allclose(I.lower, OUT[name + "/lower"], rtol=1e-14)
    np.testing.assert_allclose(I.upper, OUT[name + "/upper"], rtol=1e-14)
    assert [int(p) for p in I.unique_positions] == CASES[name]["unique_positions"]
    assert [int(v) for v in I.dfe_indices] == CASES[name]["dfe_indices"]
    for i in range(X.shape[0]):
        x = X[i].copy()
        I.clear_cache()
        if np.isnan(ll[i]):
            with pytest.raises(ValueError):
                I.loglike(x)
            continue
        comp = []
        got = I.loglike(x, comp)
        np.testing.assert_array_equal(x, OUT[name + "/X_after"][i])     # the focal population size is overwritten in place
        if np.isneginf(ll[i]):
            assert np.isneginf(got)
            continue
        assert abs(got - ll[i]) <= 1e-12 * abs(ll[i])
        for k, (a, b) in enumerate(comp):
            np.testing.assert_allclose(a, OUT[name + "/loglikes_%d" % k][i], rtol=1e-12)
            np.testing.assert_allclose(b, OUT[name + "/log_asc_%d" % k][i], rtol=1e-12)
    for xp, lp in zip(OUT[name + "/Xp"], OUT[name + "/log_posterior"]):
        I.clear_cache()
        got = I.log_posterior(xp)
        assert (np.isneginf(lp) and np.isneginf(got)) or abs(got - lp) <= 1e-12 * abs(lp)


def test_oracle_operators_match_reference():
    drift = orc.TransitionData(TB["drift"], TB["gens"], TB["us"], int(TB["N"]), TB["freqs"], TB["breaks"])
    child, lens, alphas = OUT["unit/child"], OUT["unit/lengths"], OUT["unit/alphas"]
    p = OUT["unit/parent_in"].copy()
    ret = osel.compute_transition_likelihood_selection(child, p, lens, alphas, drift)
    np.testing.assert_allclose(ret, OUT["unit/sel_rate_ret"], rtol=1e-14)
    np.testing.assert_allclose(p, OUT["unit/sel_rate_out"], rtol=1e-14)
    p = OUT["unit/parent_in"].copy()
    ret = osel.compute_transition_likelihood_selection(child, p, 0.61, alphas, drift)
    np.testing.assert_allclose(ret, OUT["unit/sel_length_ret"], rtol=1e-14)
    np.testing.assert_allclose(p, OUT["unit/sel_length_out"], rtol=1e-14)
    rate = float(OUT["unit/mask_rate"])
    for mask, tag in ((1, "zero"), (2, "focal"), (3, "desc")):
        p = OUT["unit/parent_in"].copy()
        osel.compute_masked_transition_likelihood(child, p, lens, rate, drift, mask)
        np.testing.assert_allclose(p, OUT["unit/rate_%s_out" % tag], rtol=1e-14, atol=0)
        p = OUT["unit/parent_in"].copy()
        osel.compute_masked_transition_likelihood(child, p, 0.61, rate, drift, mask)
        np.testing.assert_allclose(p, OUT["unit/length_%s_out" % tag], rtol=1e-14, atol=0)
    tb = H.load_tables()
    bot = orc.TransitionDataBottleneck(tb["bot"], tb["nbs"], tb["bot_us"], int(tb["N"]))
    for mask, tag in ((1, "zero"), (2, "focal"), (3, "desc")):
        p = OUT["unit/bot_parent_in"].copy()
        osel.compute_masked_transition_likelihood(OUT["unit/bot_child"], p, float(OUT["unit/bot_nb"]), float(OUT["unit/bot_rate"]), bot, mask)
        np.testing.assert_allclose(p, OUT["unit/bot_%s_out" % tag], rtol=1e-14, atol=0)


def _hand_one_locus(I, x, twice=False):
    """The recomputation of mope/test_selection.py:80-135 (test_one_locus / test_one_locus_twice): one fixed-length branch
    per leaf, the same matrix for both, polymorphism probability from the first and last columns."""
    nv = 1
    length = 10 ** x[0]
    raw = x[-1]
    alpha = 0.0 if abs(raw) < 10 else np.sign(raw) * (abs(raw) - 10)
    root = 2 * nv + 4
    ab, pp = 10 ** x[root], 10 ** x[root + 1]
    I.clear_cache()
    P = np.array(I.transition_data.get_transition_probabilities_2d(length, alpha))
    stat = orc.get_stationary_distribution_double_beta(I.breaks, I.transition_N, ab, pp)
    ped = I.peds[0]
    tot = 0.0
    for i in range(ped.num_loci):
        like = np.dot(P, ped.leaf_likes["mother_blood"][i]) * np.dot(P, ped.leaf_likes["child_blood"][i])
        poly = 1.0 - P[:, 0] - P[:, -1]
        tot += np.log(np.dot(stat, like)) - np.log(np.dot(stat, poly * poly))
    dfe = osel.CygnusDistribution().get_loglike(x[2 * nv:2 * nv + 4], np.array([alpha]))
    return tot + dfe


@pytest.mark.parametrize("name", ["sel_one_locus", "sel_one_locus_twice"])
def test_reference_style_hand_recomputation(name):
    I = _oracle(CASES[name])
    X, ll = OUT[name + "/X"], OUT[name + "/loglike"]
    for i in range(X.shape[0]):
        if not np.isfinite(ll[i]):
            continue
        hand = _hand_one_locus(I, OUT[name + "/X_after"][i])
        assert abs(hand - ll[i]) <= 1e-11 * abs(ll[i])           # against the reference's value
        assert abs(hand - I.loglike(X[i].copy())) <= 1e-11 * abs(ll[i])


def test_product_host_pieces_without_gpu():
    """The host side of mope_b200/selection.py that needs no device: DFE density and parameter unpacking."""
    from types import SimpleNamespace
    from mope_b200 import selection as S
    rng = np.random.RandomState(2)
    q = rng.uniform([-6, -6, -4, -6], [6, 6, 4, 0])
    a = np.array([0.0, 1.3, -0.2, 0.0, 7.0, -3.3])
    assert abs(S.CygnusDistribution().get_loglike(q, a) - osel.CygnusDistribution().get_loglike(q, a)) <= 1e-13
    name = "sel_two_loci_dfes"
    O = _oracle(CASES[name])
    stub = SimpleNamespace(focal_branch_varpar_idx=O.num_varnames, num_unique_positions=O.num_unique_positions, abs_sel_zero_value=10,
                           num_varnames=O.num_varnames, focal_branch_idx=0, total_num_dfe_params=O.total_num_dfe_params)
    for x in OUT[name + "/X"]:
        x1, x2 = x.copy(), x.copy()
        _, sel, dfe, lav = S.unpack_selection_params(stub, x1)
        _, sel_o, dfe_o, lav_o = O.unpack(x2)
        np.testing.assert_array_equal(x1, x2)
        np.testing.assert_array_equal(sel, sel_o)
        np.testing.assert_array_equal(dfe, dfe_o)
        np.testing.assert_array_equal(lav, lav_o)
    # x == -10 exactly: the reference's two sequential filters move it to -10, not to 0 (inference.py:787-795)
    x = OUT[name + "/X"][0].copy()
    x[-1] = -10.0
    assert S.unpack_selection_params(stub, x)[1][-1] == -10.0


# ------------------------------------------------------------------------------------------------ GPU
def _tables():
    from mope_b200 import TransitionTables
    return TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=float(TB["N"]), freqs=TB["freqs"], breaks=TB["breaks"])


def _inference(case):
    from mope_b200 import Inference
    opts = dict(log_unif_drift=True, min_freq=0.001, min_phred_score=None)
    opts.update(case["opts"])
    peds = case["peds"]
    return Inference([p["data"] for p in peds], [p["tree"] for p in peds], [p["ages"] for p in peds], _tables(), selection_model=True,
                     _inputs_are_text=True, **opts)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_selection_loglike_matches_reference(name):
    from mope_b200 import OutOfTableError, selection as S
    case = CASES[name]
    I = _inference(case)
    try:
        X, ll = OUT[name + "/X"], OUT[name + "/loglike"]
        assert I.ndim == X.shape[1]
        assert I.header == case["header"]
        assert list(I.varnames) == case["varnames"]
        np.testing.assert_allclose(I.lower, OUT[name + "/lower"], rtol=1e-14)
        np.testing.assert_allclose(I.upper, OUT[name + "/upper"], rtol=1e-14)
        launches0 = I.engine.launch_count()
        for i in range(X.shape[0]):
            x = X[i].copy()
            if np.isnan(ll[i]):
                with pytest.raises((ValueError, OutOfTableError)):
                    I.loglike(x)
                continue
            got = I.loglike(x)
            np.testing.assert_array_equal(x, OUT[name + "/X_after"][i])
            if np.isneginf(ll[i]):
                assert np.isneginf(got)
                continue
            assert abs(got - ll[i]) <= 1e-10 * abs(ll[i]), (i, got, ll[i])
            # per-locus values of likelihoods.get_log_l_and_asc_prob_selection, in the order of the data file
            varparams, sel, dfe, lav = S.unpack_selection_params(I, x.copy())
            assert abs(S.get_dfe_loglikes(I, dfe, sel) - OUT[name + "/dfe_ll"][i]) <= 1e-12 * max(1.0, abs(OUT[name + "/dfe_ll"][i]))
            ab, pp = 10 ** varparams[-2:]
            stat = orc.get_stationary_distribution_double_beta(I.breaks, int(I.transition_N), ab, pp)
            for k in range(len(case["peds"])):
                tr = I.translate_indices[k]
                nb = I.num_branches[k]
                a, b = S.get_log_l_and_asc_prob_selection(I, 10 ** varparams[tr[:nb]], lav[tr[:nb], :], stat, k)
                np.testing.assert_allclose(a, OUT[name + "/loglikes_%d" % k][i], rtol=1e-12)
                np.testing.assert_allclose(b, OUT[name + "/log_asc_%d" % k][i], rtol=1e-12)
        assert I.engine.launch_count() > launches0
        # batch entry point and the posterior (box prior of the selection layout)
        fin = np.isfinite(ll)
        np.testing.assert_allclose(I.loglike_batch(X[fin][:3]), ll[fin][:3], rtol=1e-10)
        lp = OUT[name + "/log_posterior"]
        assert np.isfinite(lp[:2]).all() and np.isneginf(lp[2])
        for xp, v in zip(OUT[name + "/Xp"], lp):
            got = I.log_posterior(xp)
            assert (np.isneginf(v) and np.isneginf(got)) or abs(got - v) <= 1e-10 * abs(v)
        np.testing.assert_allclose(I.log_posterior_batch(OUT[name + "/Xp"])[:2], lp[:2], rtol=1e-10)
    finally:
        I.close()


@pytest.mark.gpu
def test_gpu_selection_operators_match_reference():
    from mope_b200 import selection as S
    I = _inference(CASES["sel_one_locus"])
    try:
        E = I.engine
        child, lens, alphas = OUT["unit/child"], OUT["unit/lengths"], OUT["unit/alphas"]
        p = OUT["unit/parent_in"].copy()
        ret = S.masked_branch_update(E, S.MASK_NONE, child, p, lens, alphas, want_products=True)
        np.testing.assert_allclose(ret, OUT["unit/sel_rate_ret"], rtol=1e-13)
        np.testing.assert_allclose(p, OUT["unit/sel_rate_out"], rtol=1e-13)
        p = OUT["unit/parent_in"].copy()
        ret = S.masked_branch_update(E, S.MASK_NONE, child, p, 0.61, alphas, want_products=True)
        np.testing.assert_allclose(ret, OUT["unit/sel_length_ret"], rtol=1e-13)
        np.testing.assert_allclose(p, OUT["unit/sel_length_out"], rtol=1e-13)
        rate = float(OUT["unit/mask_rate"])
        for mask, tag in ((S.MASK_ZERO, "zero"), (S.MASK_FOCAL_NODE, "focal"), (S.MASK_DESCENDANT, "desc")):
            p = OUT["unit/parent_in"].copy()
            S.masked_branch_update(E, mask, child, p, lens, rate)
            np.testing.assert_allclose(p, OUT["unit/rate_%s_out" % tag], rtol=1e-13, atol=0)
            p = OUT["unit/parent_in"].copy()
            S.masked_branch_update(E, mask, child, p, 0.61, rate)
            np.testing.assert_allclose(p, OUT["unit/length_%s_out" % tag], rtol=1e-13, atol=0)
        # out-of-table coefficient: the reference raises from the fetch and leaves the parent untouched
        from mope_b200 import OutOfTableError
        p = OUT["unit/parent_in"].copy()
        bad = alphas.copy()
        bad[4] = 2 * float(TB["N"]) * TB["us"][-1] * 1.01
        with pytest.raises(OutOfTableError):
            S.masked_branch_update(E, S.MASK_NONE, child, p, lens, bad)
        np.testing.assert_array_equal(p, OUT["unit/parent_in"])
    finally:
        I.close()


@pytest.mark.gpu
def test_gpu_masked_bottleneck_operators_match_reference():
    from mope_b200 import Inference, TransitionTables, selection as S
    tb = H.load_tables()
    tables = TransitionTables(drift=tb["drift"], gens=tb["gens"], us=tb["us"], N=float(tb["N"]), freqs=tb["freqs"], breaks=tb["breaks"],
                              bot=tb["bot"], nbs=tb["nbs"], bot_us=tb["bot_us"])
    ped = H.load_cases()["example_both"]["peds"][0]
    I = Inference([ped["data"]], [ped["tree"]], [ped["ages"]], tables, log_unif_drift=True, _inputs_are_text=True)
    try:
        for mask, tag in ((S.MASK_ZERO, "zero"), (S.MASK_FOCAL_NODE, "focal"), (S.MASK_DESCENDANT, "desc")):
            p = OUT["unit/bot_parent_in"].copy()
            S.masked_branch_update(I.engine, mask, OUT["unit/bot_child"], p, float(OUT["unit/bot_nb"]), float(OUT["unit/bot_rate"]), bottleneck=True)
            np.testing.assert_allclose(p, OUT["unit/bot_%s_out" % tag], rtol=1e-13, atol=0)
    finally:
        I.close()


@pytest.mark.gpu
def test_gpu_selection_matches_oracle_on_fresh_vectors():
    """Seeded vectors that are not in the fixtures, the larger data set."""
    name = "sel_example"
    I = _inference(CASES[name])
    O = _oracle(CASES[name])
    try:
        rng = np.random.RandomState(99)
        npos = I.num_unique_positions
        for _ in range(3):
            x = rng.uniform(0.75 * I.lower + 0.25 * I.upper, 0.25 * I.lower + 0.75 * I.upper)
            x[I.num_varnames:2 * I.num_varnames] = rng.uniform(-0.2, 0.2, I.num_varnames)
            x[-npos:] = rng.choice([-1.0, 1.0], npos) * (10.0 + rng.uniform(0, 5, npos)) * (rng.rand(npos) < 0.7)
            O.clear_cache()
            ref = O.loglike(x.copy())
            got = I.loglike(x.copy())
            assert abs(got - ref) <= 1e-10 * abs(ref)
    finally:
        I.close()


@pytest.mark.gpu
def test_gpu_selection_on_the_benchmark_grid_size_vs_oracle():
    """F = 201 (Fp = 208: the other interpolation-kernel staging variant, the k-permuted emission layout of the main path read
    back in natural order): the F = 201 fixture's drift matrices with their rate axis relabelled as selection coefficients
    (any complete grid of stochastic matrices is a valid table set for an oracle-vs-GPU comparison)."""
    from mope_b200 import Inference, TransitionTables
    tb = H.load_tables("tables_f201.npz")
    ss = np.array([-2e-3, 3e-3])
    assert tb["drift"].shape[1] == ss.shape[0]
    tables = TransitionTables(drift=tb["drift"], gens=tb["gens"], us=ss, N=float(tb["N"]), freqs=tb["freqs"], breaks=tb["breaks"])
    case = CASES["sel_example"]
    peds = case["peds"]
    I = Inference([p["data"] for p in peds], [p["tree"] for p in peds], [p["ages"] for p in peds], tables, selection_model=True,
                  log_unif_drift=True, min_freq=0.001, _inputs_are_text=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        drift = orc.TransitionData(tb["drift"], tb["gens"], ss, int(tb["N"]), tb["freqs"], tb["breaks"])
        O = osel.SelectionInference([(H.read_tsv(p["data"]), p["tree"], H.read_tsv(p["ages"])) for p in peds], drift, log_unif_drift=True,
                                    min_freq=0.001)
    try:
        rng = np.random.RandomState(4)
        npos = I.num_unique_positions
        for _ in range(2):
            x = rng.uniform(0.7 * I.lower + 0.3 * I.upper, 0.3 * I.lower + 0.7 * I.upper)
            x[I.num_varnames:2 * I.num_varnames] = rng.uniform(-0.2, 0.2, I.num_varnames)
            x[-npos:] = rng.choice([-1.0, 1.0], npos) * (10.0 + rng.uniform(0, 2.5, npos)) * (rng.rand(npos) < 0.7)
            O.clear_cache()
            ref = O.loglike(x.copy())
            got = I.loglike(x.copy())
            assert np.isfinite(ref) and abs(got - ref) <= 1e-10 * abs(ref), (got, ref)
    finally:
        I.close()

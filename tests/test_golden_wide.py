"""Second golden set (tests/golden/make_golden_wide.py, produced by running the reference): the benchmark's grid sizes
(F = 115, F = 201), tables reaching 10 000 generations (stored row sums up to 1 + 5.6e-9: the renormalisation branch of
_util.check_transition_matrix matters most here), cfg3- and cfg4-shaped data sets with 200 families of distinct ages,
and the binomial emission tensor at the benchmark's coverage (n ~ 1000).  CPU: the oracle against these vectors; GPU: the
CUDA path through the C ABI against them."""
import json
import os

import numpy as np
import pytest

from tests import helpers as H

with open(os.path.join(H.GOLDEN, "cases_wide.json")) as _f:
    CASES = json.load(_f)
OUT = {k: v for k, v in np.load(os.path.join(H.GOLDEN, "outputs_wide.npz")).items()}
_TABLES = {}


def _tb(name):
    if name not in _TABLES:
        _TABLES[name] = H.load_tables(name)
    return _TABLES[name]


def _key(name):
    return name.replace("/", "_")


# ------------------------------------------------------------------------------------------------ CPU: oracle
@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_oracle_matches_reference_on_wide_cases(name):
    case = CASES[name]
    tb = _tb(case["tables"])
    I = H.oracle_inference(case, tb)
    k = _key(name)
    np.testing.assert_allclose(I.lower, OUT[k + "/lower"], rtol=1e-15)
    np.testing.assert_allclose(I.upper, OUT[k + "/upper"], rtol=1e-15)
    X = OUT[k + "/X"]
    for i in (0, X.shape[0] - 1):          # one central and one wide vector (the oracle is a slow NumPy restatement)
        I.clear_cache()
        ll = I.loglike(X[i])
        assert H.relerr(ll, OUT[k + "/loglike"][i]) < 1e-12, (name, i)


def test_oracle_long_grid_matrices_and_emissions():
    from oracle import mope_oracle as orc
    for tag in ("f115", "f201"):
        tb = _tb("tables_%s.npz" % tag)
        drift, _ = H.oracle_tables(tb)
        for t, m, P in zip(OUT["unit/%s_drift_t" % tag], OUT["unit/%s_drift_m" % tag], OUT["unit/%s_drift_P" % tag]):
            drift.clear_cache()
            np.testing.assert_allclose(drift.get_transition_probabilities_2d(t, m), P, rtol=1e-15, atol=0)
            assert P.sum(1).max() <= 1.0 + 1e-13          # the reference's output is renormalised ...
        raw = tb["drift"][-1].sum(-1).max()
        assert raw > 1.0 + 1e-9                            # ... from stored matrices whose row sums exceed 1
    freqs = _tb("tables_f201.npz")["freqs"]
    for key, phred in (("unit/f201_binom_like", -1.0), ("unit/f201_binom_like_phred30", 30.0)):
        got = orc.get_binom_likelihoods(OUT["unit/f201_binom_X"], OUT["unit/f201_binom_n"], freqs, phred)
        np.testing.assert_allclose(got, OUT[key], rtol=2e-13, atol=1e-300)


# ------------------------------------------------------------------------------------------------ GPU: CUDA path
def _tables(tb):
    from mope_b200 import TransitionTables
    return TransitionTables(drift=tb["drift"], gens=tb["gens"], us=tb["us"], N=float(tb["N"]), freqs=tb["freqs"],
                            breaks=tb["breaks"], bot=tb["bot"], nbs=tb["nbs"], bot_us=tb["bot_us"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_gpu_matches_reference_on_wide_cases(name):
    from mope_b200 import Inference
    case = CASES[name]
    tb = _tb(case["tables"])
    peds = case["peds"]
    I = Inference([p["data"] for p in peds], [p["tree"] for p in peds], [p["ages"] for p in peds], _tables(tb),
                  _inputs_are_text=True, log_unif_drift=True)
    try:
        k = _key(name)
        np.testing.assert_allclose(I.lower, OUT[k + "/lower"], rtol=1e-15)
        X = OUT[k + "/X"]
        ll, status, comps = I.loglike_components(X)
        assert not status.any()
        rel = H.relerr(ll, OUT[k + "/loglike"])
        assert rel.max() <= 1e-10, (name, rel.max())
        assert H.relerr(comps[0]["root_ll"], OUT[k + "/root_ll"][:, 0]).max() <= 1e-10
        np.testing.assert_allclose(comps[0]["log_asc"], OUT[k + "/log_asc_0"], rtol=1e-9, atol=1e-14)
    finally:
        I.close()


@pytest.mark.gpu
def test_gpu_long_grid_matrices_and_emissions():
    from mope_b200 import Inference
    case = CASES["f201/cfg2_shaped"]
    tb = _tb("tables_f201.npz")
    p = case["peds"][0]
    I = Inference([p["data"]], [p["tree"]], [p["ages"]], _tables(tb), _inputs_are_text=True, log_unif_drift=True)
    try:
        E = I.engine
        for t, m, P in zip(OUT["unit/f201_drift_t"], OUT["unit/f201_drift_m"], OUT["unit/f201_drift_P"]):
            np.testing.assert_allclose(E.transition_matrix(t, m), P, rtol=1e-15, atol=0)
        # emissions at n ~ 1000: ln C(n, k) comes from the same libm lgamma calls the reference makes, so an element differs
        # from the reference's by the rounding of one exp: a few ulp, two decades inside the 1e-13 of SURVEY section 8c
        X, n = OUT["unit/f201_binom_X"], OUT["unit/f201_binom_n"]
        for key, phred in (("unit/f201_binom_like", -1.0), ("unit/f201_binom_like_phred30", 30.0)):
            got = E.leaf_likelihoods(X, n, tb["freqs"], phred)
            ref = OUT[key]
            np.testing.assert_allclose(got, ref, rtol=2e-15, atol=1e-300)
            assert np.array_equal(got == 0, ref == 0) and np.array_equal(got == 1, ref == 1)
    finally:
        I.close()


@pytest.mark.gpu
def test_gpu_table_generator_matches_reference_tables():
    """mope_b200.generate on the device against tables the reference's own generate_transitions wrote (F = 115 and 201,
    generations up to 10 000, bottleneck sizes 2..500)."""
    from mope_b200 import generate
    for tag, breaks in (("f115", (0.5, 0.01)), ("f201", (0.5, 0.005))):
        tb = _tb("tables_%s.npz" % tag)
        got = generate.generate_tables(N=int(tb["N"]), breaks=breaks, gens=[int(g) for g in tb["gens"]], us=list(tb["us"]),
                                       nbs=[int(b) for b in tb["nbs"]], bot_us=list(tb["bot_us"]), device="cuda")
        np.testing.assert_array_equal(got.breaks, tb["breaks"])
        np.testing.assert_allclose(got.freqs, tb["freqs"], rtol=1e-15)
        np.testing.assert_allclose(got.drift, tb["drift"], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(got.bot, tb["bot"], rtol=1e-9, atol=1e-300)

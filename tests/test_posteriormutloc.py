"""Posterior-mutation-location likelihoods (mope/posteriormutloc.py:48-271): the oracle restatement against
reference-generated fixtures (tests/golden/make_golden_pml.py) on the CPU, and mope_b200.posteriormutloc -- every branch
update a C-ABI operator call on the GPU -- against the same fixtures."""
import os

import numpy as np
import pytest

from tests import helpers as H
from oracle import mope_oracle_pml as opml

TB = H.load_tables()
CASES = H.load_cases()
_z = np.load(os.path.join(H.GOLDEN, "outputs_pml.npz"))
OUT = {k: _z[k] for k in _z.files}
NAMES = ["example_both", "example_fixed"]


def _same(got, ref, rtol):
    got, ref = np.asarray(got), np.asarray(ref)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=rtol)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    I = H.oracle_inference(CASES[name], TB)
    nodes = [str(n) for n in OUT[name + "/nodes"]]
    for i, x in enumerate(OUT[name + "/X"]):
        I.clear_cache()
        a, b = opml.fill_cond_probs(I, 0, x, include_mutation_only=True)
        _same(a, OUT[name + "/overall"][i], 1e-12)
        _same(b, OUT[name + "/mutation_only"][i], 1e-12)
        for j, nd in enumerate(nodes):
            I.clear_cache()
            _same(opml.fill_cond_probs_mutation(nd, I, 0, x), OUT[name + "/focal"][i, j], 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_reference(name):
    from mope_b200 import Inference, TransitionTables, posteriormutloc as pml
    tables = TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=float(TB["N"]), freqs=TB["freqs"], breaks=TB["breaks"],
                              bot=TB["bot"], nbs=TB["nbs"], bot_us=TB["bot_us"])
    ped = CASES[name]["peds"][0]
    I = Inference([ped["data"]], [ped["tree"]], [ped["ages"]], tables, log_unif_drift=True, _inputs_are_text=True)
    try:
        nodes = [str(n) for n in OUT[name + "/nodes"]]
        assert nodes == list(I.branch_names[0])
        launches0 = I.engine.launch_count()
        for i, x in enumerate(OUT[name + "/X"]):
            a, b = pml.fill_cond_probs(I, 0, x, include_mutation_only=True)
            _same(a, OUT[name + "/overall"][i], 1e-11)
            _same(b, OUT[name + "/mutation_only"][i], 1e-11)
            _same(pml.fill_cond_probs(I, 0, x), OUT[name + "/overall"][i], 1e-11)
            for j, nd in enumerate(nodes):
                _same(pml.fill_cond_probs_mutation(nd, I, 0, x), OUT[name + "/focal"][i, j], 1e-11)
        assert I.engine.launch_count() > launches0
        with pytest.raises(ValueError):
            pml.fill_cond_probs_mutation("no_such_node", I, 0, OUT[name + "/X"][0])
    finally:
        I.close()

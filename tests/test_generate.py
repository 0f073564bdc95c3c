"""The table generator against the tables the reference's own generate_transitions produced (golden fixture)."""
import numpy as np
import pytest

from mope_b200 import generate
from tests import helpers as H

TB = H.load_tables()


def test_breaks_and_frequencies_match_reference():
    br = generate.get_breaks_symmetric(1000, 0.5, 0.02)
    np.testing.assert_array_equal(br, TB["breaks"])
    np.testing.assert_allclose(generate.get_binned_frequencies(1000, br), TB["freqs"], rtol=1e-15)
    assert generate.get_breaks_symmetric(1000, 0.5, 0.005).shape[0] == 201
    assert generate.get_breaks_symmetric(1000, 0.5, 0.01).shape[0] == 115
    assert len(generate.DEFAULT_MUTS) == 44 and len(generate.DEFAULT_GENS) == 94 and len(generate.DEFAULT_BOTS) == 48


@pytest.mark.slow
def test_generated_tables_match_reference_generator():
    tb = generate.generate_tables(N=1000, breaks=(0.5, 0.02), gens=[1, 2, 5, 10, 20, 50], us=[1e-7, 1e-3],
                                  nbs=[2, 20, 500], bot_us=[1e-7, 1e-3], device="cpu")
    gi = [list(TB["gens"]).index(g) for g in (1, 2, 5, 10, 20, 50)]
    ui = [list(TB["us"]).index(u) for u in (1e-7, 1e-3)]
    ref = TB["drift"][np.ix_(gi, ui)]
    # different product order / BLAS than the reference run: agreement to rounding, not bitwise
    np.testing.assert_allclose(tb.drift, ref, rtol=1e-9, atol=1e-15)
    bi = [list(TB["nbs"]).index(b) for b in (2, 20, 500)]
    np.testing.assert_allclose(tb.bot, TB["bot"][np.ix_(bi, ui)], rtol=1e-9, atol=1e-15)


def test_device_generator_path_matches_host_path():
    """The device path of the generator (binomial rows from lgamma, binning as a GEMM with the bin-membership matrix)
    run on CPU tensors against the NumPy / SciPy path: same tables to rounding."""
    import torch

    class CpuTorchMat(generate._Mat):
        def __init__(self, device):
            self.torch, self.device = torch, "cpu"

    host = generate.generate_tables(N=1000, breaks=(0.5, 0.02), gens=[1, 3, 10], us=[1e-7, 1e-3], nbs=[2, 20, 500],
                                    bot_us=[1e-7, 1e-3], device="cpu")
    orig = generate._Mat
    generate._Mat = CpuTorchMat
    try:
        dev = generate.generate_tables(N=1000, breaks=(0.5, 0.02), gens=[1, 3, 10], us=[1e-7, 1e-3], nbs=[2, 20, 500],
                                       bot_us=[1e-7, 1e-3], device="cpu")
    finally:
        generate._Mat = orig
    np.testing.assert_allclose(dev.drift, host.drift, rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(dev.bot, host.bot, rtol=1e-9, atol=1e-15)
    np.testing.assert_array_equal(dev.breaks, host.breaks)
    # rows of a transition matrix are probability vectors
    assert np.abs(dev.drift.sum(axis=-1) - 1).max() < 1e-9 and np.abs(dev.bot.sum(axis=-1) - 1).max() < 1e-9

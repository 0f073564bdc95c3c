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


def test_generated_selection_tables_match_reference_generator():
    """generate_selection_tables (one Wright-Fisher chain per selection coefficient, u = v = 0) against the selection table set
    the reference's own generator produced (tests/golden/tables_sel.npz, make_golden_selection.py)."""
    ref = H.load_tables("tables_sel.npz")
    gens, ss = [1, 5, 20, 100], [-1e-3, 0.0, 5e-3]
    tb = generate.generate_selection_tables(N=1000, breaks=(0.5, 0.02), gens=gens, ss=ss, device="cpu")
    gi = [list(ref["gens"]).index(g) for g in gens]
    si = [list(ref["us"]).index(x) for x in ss]
    np.testing.assert_allclose(tb.drift, ref["drift"][np.ix_(gi, si)], rtol=1e-9, atol=1e-15)
    np.testing.assert_array_equal(tb.us, ss)
    np.testing.assert_array_equal(tb.breaks, ref["breaks"])
    with pytest.raises(ValueError):
        generate.generate_selection_tables(N=1000, breaks=(0.5, 0.02), gens=gens, ss=[0.0, -1e-3], device="cpu")


@pytest.mark.gpu
def test_gpu_selection_tables_match_reference_generator():
    """The same on the device (binom_matrix_kernel, dgemm_tn_kernel, bin_matrix_kernel), the whole fixture grid (to 3 000
    generations, negative and positive coefficients)."""
    ref = H.load_tables("tables_sel.npz")
    tb = generate.generate_selection_tables(N=1000, breaks=(0.5, 0.02), gens=[int(g) for g in ref["gens"]], ss=[float(x) for x in ref["us"]],
                                            device="cuda")
    np.testing.assert_allclose(tb.drift, ref["drift"], rtol=1e-8, atol=1e-14)
    assert tb.drift.min() >= 0.0 and tb.drift.max() <= 1.0


@pytest.mark.gpu
def test_gpu_generator_kernels_match_host():
    """The generator's own kernels through the C ABI (mope_b200_gen_*): binomial rows vs SciPy, dgemm_tn_kernel vs a NumPy
    fp64 product (odd sizes, rectangular tile edges), bin_matrix_kernel vs bin_matrix, and the chained tables vs the host path."""
    import torch
    mat = generate._Mat("cuda")
    N = 1000
    P = mat.wright_fisher(N, 0.0, 1e-3, 1e-3)
    ref = generate.wright_fisher_matrix(N, 0.0, 1e-3, 1e-3)
    np.testing.assert_allclose(mat.get(P), ref, rtol=1e-10, atol=1e-300)
    V = mat.var_n(20, 40, N, 1e-5)
    np.testing.assert_allclose(mat.get(V), generate.var_n_matrix(20, 40, N, 1e-5), rtol=1e-10, atol=1e-300)
    rng = np.random.RandomState(3)
    for n in (1001, 65, 208, 333):
        a, b = rng.rand(n, n), rng.rand(n, n) - 0.5
        got = mat.get(mat.mm(mat.put(a), mat.put(b)))
        np.testing.assert_allclose(got, a @ b, rtol=0, atol=1e-12 * n)
    br = generate.get_breaks_symmetric(N, 0.5, 0.02)
    np.testing.assert_allclose(mat.binner(N + 1, br)(P), generate.bin_matrix(mat.get(P), br), rtol=1e-14, atol=1e-300)
    host = generate.generate_tables(N=N, breaks=(0.5, 0.02), gens=[1, 3, 10], us=[1e-7, 1e-3], nbs=[2, 20, 500],
                                    bot_us=[1e-7, 1e-3], device="cpu")
    dev = generate.generate_tables(N=N, breaks=(0.5, 0.02), gens=[1, 3, 10], us=[1e-7, 1e-3], nbs=[2, 20, 500],
                                   bot_us=[1e-7, 1e-3], device="cuda")
    np.testing.assert_allclose(dev.drift, host.drift, rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(dev.bot, host.bot, rtol=1e-9, atol=1e-15)
    assert np.abs(dev.drift.sum(axis=-1) - 1).max() < 1e-9 and np.abs(dev.bot.sum(axis=-1) - 1).max() < 1e-9
    torch.cuda.synchronize()

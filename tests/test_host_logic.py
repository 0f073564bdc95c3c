"""CPU tests of the host side of mope_b200 (no CUDA calls): parser, model compiler, prior box, root prior,
and that the C-ABI library loads and exports every symbol include/mope_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from mope_b200 import _lib, model, newick, stationary, tables
from oracle import mope_oracle as orc
from tests import helpers as H

TB = H.load_tables()
CASES = H.load_cases()
OUT = H.load_outputs()


def test_library_exports_every_declared_symbol():
    _lib.build()
    hdr = open(os.path.join(H.GOLDEN, "..", "..", "include", "mope_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mope_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in _lib.lib().mope_b200_version()


def test_newick_matches_oracle_parser():
    for case in CASES.values():
        for ped in case["peds"]:
            a = newick.loads(ped["tree"])[0]
            b = orc.parse_newick(ped["tree"])
            la = [(n.name, n.varname, n.multipliername, n.is_bottleneck, len(n.children)) for n in a.postorder()]
            lb = [(n.name, n.varname, n.multipliername, n.is_bottleneck, len(n.children)) for n in b.postorder()]
            assert la == lb


def test_newick_errors():
    with pytest.raises(ValueError):
        newick.loads("((a:x,b:y)c:z;")
    with pytest.raises(ValueError):
        newick.loads("(a:x*y*z,b:y)r;")
    t = newick.loads("(a:x^,b:y*age)r;")[0]
    assert [(n.name, n.varname, n.multipliername, n.is_bottleneck) for n in t.postorder()] == \
        [("a", "x", None, True), ("b", "y", "age", False), ("r", None, None, False)]


@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_model_compile_matches_reference(name):
    case = CASES[name]
    for k, ped in enumerate(case["peds"]):
        pm = model.compile_pedigree(model.read_table(ped["data"], True), ped["tree"], model.read_table(ped["ages"], True),
                                    sort_rows=False)
        assert pm.branch_names == case["branch_names"][k]
        assert pm.leaf_names == case["leaf_names"][k]
        assert pm.num_loci == int(OUT[name + "/num_loci_%d" % k])
        np.testing.assert_array_equal(pm.fam_count.astype(np.int64), OUT[name + "/counts_%d" % k])
        # parents come after children, the root is node n_branches
        assert np.all(pm.br_parent > np.arange(len(pm.branch_names)))
        assert pm.br_parent.max() == len(pm.branch_names)
        # ascertainment pairs are the distinct multiplier tuples
        assert pm.asc_mult.shape[1] == len({tuple(pm.asc_mult[:, pm.fam_asc[f]]) for f in range(pm.n_fam)})
        o = H.oracle_inference(case, TB).peds[k]
        for j, m in enumerate(pm.multipliernames):
            np.testing.assert_array_equal(pm.row_mult[j], o.row_mult[m])
            np.testing.assert_array_equal(pm.asc_mult[j][pm.fam_asc], np.asarray(o.ages[m], dtype=np.float64))


def test_row_sorting_is_a_consistent_permutation():
    """The k-d ordering of families (narrow age ranges per 64-row tile) only permutes rows; every row keeps its
    counts, coverage, family and multiplier values, and families map to the same ascertainment multiplier tuples."""
    ped = CASES["example_both"]["peds"][0]
    args = (model.read_table(ped["data"], True), ped["tree"], model.read_table(ped["ages"], True))
    a = model.compile_pedigree(*args, sort_rows=False)
    b = model.compile_pedigree(*args, sort_rows=True)

    def rows(pm):
        return sorted((tuple(np.nan_to_num(pm.counts[i], nan=-1.0)), tuple(pm.coverage[i]), pm.family[i], tuple(pm.row_mult[:, i]))
                      for i in range(pm.num_loci))
    assert rows(a) == rows(b)
    assert not np.array_equal(a.family, b.family)          # it did reorder something
    # row_order maps every row back to its row of the (filtered) data file
    np.testing.assert_array_equal(a.row_order, np.arange(a.num_loci))
    np.testing.assert_array_equal(np.sort(b.row_order), np.arange(a.num_loci))
    np.testing.assert_array_equal(a.coverage[b.row_order], b.coverage)
    np.testing.assert_array_equal(a.fam_count, b.fam_count)
    for f in range(a.n_fam):
        np.testing.assert_array_equal(a.asc_mult[:, a.fam_asc[f]], b.asc_mult[:, b.fam_asc[f]])
    # families of neighbouring rows are close in every multiplier column: tile-wise spread shrinks
    def tile_spread(pm, tile=16):
        lm = np.log(pm.row_mult)
        return np.mean([np.ptp(lm[:, i:i + tile], axis=1).max() for i in range(0, pm.num_loci, tile)])
    assert tile_spread(b) < 0.6 * tile_spread(a)


def test_stationary_distribution_bit_identical_to_reference():
    abpp = OUT["unit/stat_abpp"]
    got = stationary.stationary_distribution_batch(TB["breaks"], int(TB["N"]), abpp[:, 0], abpp[:, 1])
    np.testing.assert_array_equal(got, OUT["unit/stat_dist"])
    # and against the oracle restatement on random draws over the whole prior box
    rng = np.random.RandomState(5)
    ab, pp = 10 ** rng.uniform(-9, 0, 40), 10 ** rng.uniform(-9, 0, 40)
    got = stationary.stationary_distribution_batch(TB["breaks"], int(TB["N"]), ab, pp)
    for i in range(40):
        np.testing.assert_array_equal(got[i], orc.get_stationary_distribution_double_beta(TB["breaks"], int(TB["N"]), ab[i], pp[i]))


def test_beta_cdf_table_is_scipy_beta_cdf_bit_for_bit():
    """mope_b200_beta_cdf_table (native threads through SciPy's betainc C entry point) == scipy.stats.beta.cdf, the call
    likelihoods.discretize_beta_distn makes (mope/likelihoods.py:40), over the whole prior range of ab."""
    import scipy.stats as st
    ab = 10 ** np.random.RandomState(3).uniform(-9, 0, 37)
    edges = np.arange(1, 2 * 998, 2) / (2 * 998)
    for nt in (1, 3, 16):
        got = stationary.beta_cdf_table(ab, edges, n_threads=nt)
        for i in (0, 5, 36):
            np.testing.assert_array_equal(got[i], st.beta.cdf(edges, ab[i], ab[i]))


def test_tables_validation(tmp_path):
    tb = tables.TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=float(TB["N"]), freqs=TB["freqs"],
                                 breaks=TB["breaks"], bot=TB["bot"], nbs=TB["nbs"], bot_us=TB["bot_us"])
    p = str(tmp_path / "t.npz")
    tb.save(p)
    tb2 = tables.load(p, p)
    np.testing.assert_array_equal(tb2.drift, tb.drift)
    assert tb2.get_max_coal_time() == TB["gens"].max() / float(TB["N"])
    assert tables.load(p, None).bot is None
    with pytest.raises(ValueError):
        tables.TransitionTables(drift=TB["drift"], gens=TB["gens"][::-1], us=TB["us"], N=1000, freqs=TB["freqs"], breaks=TB["breaks"])
    with pytest.raises(ValueError):
        tables.TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=1000, freqs=TB["freqs"], breaks=np.int64(0))


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from mope_b200 import Inference
    case = CASES["test_simple"]
    tb = tables.TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=float(TB["N"]), freqs=TB["freqs"],
                                 breaks=TB["breaks"], bot=TB["bot"], nbs=TB["nbs"], bot_us=TB["bot_us"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Inference([case["peds"][0]["data"]], [case["peds"][0]["tree"]], [case["peds"][0]["ages"]], tb,
                  log_unif_drift=True, _inputs_are_text=True)


class _FakeInf:
    """Counts the batched entry points GpuPool.map must use (no device involved)."""

    def __init__(self):
        self.calls = []

    def log_posterior_batch(self, X):
        self.calls.append(("post", X.shape))
        return np.arange(X.shape[0], dtype=np.float64)

    def loglike_batch(self, X):
        self.calls.append(("ll", X.shape))
        return np.arange(X.shape[0], dtype=np.float64) + 100

    def logprior(self, x):
        self.calls.append(("prior", x.shape))
        return -1.0


class _EmceeStyleWrapper:
    """What emcee hands to pool.map: emcee.ensemble._function_wrapper (2.x) / _FunctionWrapper (3.x) -- an object without
    __name__ that calls f(x, *args, **kwargs)."""

    def __init__(self, f, args=(), kwargs=None):
        self.f, self.args, self.kwargs = f, list(args), dict(kwargs or {})

    def __call__(self, x):
        return self.f(x, *self.args, **self.kwargs)


def test_gpu_pool_batches_through_emcee_style_wrappers():
    from mope_b200 import inference as inf_mod
    fake = _FakeInf()
    pool = inf_mod.GpuPool(fake)
    xs = [np.full(4, float(i)) for i in range(6)]
    # bare trampoline and emcee-wrapped trampoline: ONE batched call each, never the serial fallback
    assert pool.map(inf_mod.post_clone, xs) == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    assert pool.map(_EmceeStyleWrapper(inf_mod.post_clone), xs) == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    assert pool.map(_EmceeStyleWrapper(inf_mod.logl), xs)[0] == 100.0
    assert fake.calls == [("post", (6, 4)), ("post", (6, 4)), ("ll", (6, 4))]
    # the reference's own module-level trampolines (module name mope.inference) are recognised too
    ref_like = type(inf_mod.post_clone)(inf_mod.post_clone.__code__, {"__name__": "mope.inference"}, "post_clone")
    ref_like.__module__ = "mope.inference"
    assert pool.map(_EmceeStyleWrapper(ref_like), xs[:2]) == [0.0, 1.0]
    assert fake.calls[-1] == ("post", (2, 4))
    # a wrapper that carries extra arguments, or any other function, is mapped serially
    fake.calls.clear()
    assert pool.map(_EmceeStyleWrapper(lambda x, k: k * x[0], args=(2.0,)), xs[:3]) == [0.0, 2.0, 4.0]
    assert pool.map(lambda x: x.sum(), xs[:2]) == [0.0, 4.0]
    assert fake.calls == []
    assert pool.map(inf_mod.post_clone, []) == []


def test_beta_cdf_table_falls_back_to_the_ufunc(monkeypatch):
    """Without SciPy's private capsule the table comes from scipy.special.betainc itself: same values bit for bit."""
    ab = 10 ** np.random.RandomState(4).uniform(-9, 0, 9)
    edges = np.arange(1, 2 * 998, 2) / (2 * 998)
    want = stationary.beta_cdf_table(ab, edges)
    monkeypatch.setattr(stationary, "_ENTRY", None)
    np.testing.assert_array_equal(stationary.beta_cdf_table(ab, edges), want)


def _descs(drift, ped_kw=None):
    """Minimal C-ABI descriptions (one 2-leaf pedigree) for the host-only validation entry point."""
    import ctypes as C
    F = drift.shape[-1]
    keep = dict(drift=np.ascontiguousarray(drift), gens=np.array([1.0, 10.0]), us=np.array([1e-8, 1e-6]),
                freqs=np.linspace(0, 1, F), counts=np.array([[1.0, 2.0]]), coverage=np.array([[10.0, 10.0]]),
                row_mult=np.array([[30.0]]), asc_mult=np.array([[30.0]]), fam_count=np.array([1.0]), fam_lgamma=np.array([0.0]),
                fam_asc=np.array([0], dtype=np.int32), br_parent=np.array([2, 2], dtype=np.int32),
                br_leaf=np.array([0, 1], dtype=np.int32), br_var=np.array([0, 0], dtype=np.int32),
                br_mult=np.array([0, -1], dtype=np.int32), br_bottleneck=np.array([0, 0], dtype=np.int32))
    keep.update(ped_kw or {})
    td = _lib.TablesDesc()
    td.F, td.n_gens, td.n_us = F, 2, 2
    td.drift, td.gens, td.us, td.freqs = (_lib.dptr(keep[k]) for k in ("drift", "gens", "us", "freqs"))
    td.n_nbs = td.n_bot_us = 0
    td.N = 100.0
    pd_ = (_lib.PedigreeDesc * 1)()
    d = pd_[0]
    d.n_leaves, d.n_branches, d.n_loci, d.n_fam, d.n_mult, d.n_asc = 2, 2, 1, 1, 1, 1
    for k in ("counts", "coverage", "row_mult", "asc_mult", "fam_count", "fam_lgamma"):
        setattr(d, k, _lib.dptr(keep[k]))
    for k in ("fam_asc", "br_parent", "br_leaf", "br_var", "br_mult", "br_bottleneck"):
        setattr(d, k, _lib.iptr(keep[k]))
    md = _lib.ModelDesc()
    md.n_var, md.n_ped, md.peds = 1, 1, pd_
    md.min_phred_score, md.min_freq, md.genome_size, md.poisson_like_penalty = -1.0, 0.001, 16569.0, 1.0
    return td, md, (keep, pd_)


def test_c_abi_validates_tables_and_multipliers():
    """The [0,1] range of the drift tables and finite multipliers are enforced at the C ABI (mope_b200_arena_bytes runs
    the same validation as mope_b200_create and needs no device)."""
    import ctypes as C
    L = _lib.lib()
    F = 4
    good = np.full((2, 2, F, F), 0.25)
    n = C.c_size_t(0)
    td, md, keep = _descs(good)
    assert L.mope_b200_arena_bytes(C.byref(td), C.byref(md), C.byref(n)) == 0 and n.value > 0
    for bad_value in (1.0 + 1e-12, -1e-300, np.nan):
        bad = good.copy()
        bad[1, 0, 2, 3] = bad_value
        td, md, keep = _descs(bad)
        assert L.mope_b200_arena_bytes(C.byref(td), C.byref(md), C.byref(n)) == -1
        assert b"outside [0, 1]" in L.mope_b200_last_error()
    td, md, keep = _descs(good, dict(row_mult=np.array([[np.nan]])))
    assert L.mope_b200_arena_bytes(C.byref(td), C.byref(md), C.byref(n)) == -1
    assert b"not finite" in L.mope_b200_last_error()


def _schedule_stats(tree_str, F):
    """Park homes of the pruning kernel for a tree (host-only query of the C ABI)."""
    import ctypes as C
    t = newick.loads(tree_str)[0]
    nodes = [nd for nd in t.postorder() if nd is not t]
    idx = {id(nd): i for i, nd in enumerate(nodes)}
    idx[id(t)] = len(nodes)
    leaves = [nd for nd in nodes if nd.is_leaf]
    keep = dict(br_parent=np.array([idx[id(nd.parent)] for nd in nodes], dtype=np.int32),
                br_leaf=np.array([leaves.index(nd) if nd.is_leaf else -1 for nd in nodes], dtype=np.int32),
                br_var=np.zeros(len(nodes), dtype=np.int32),
                br_mult=np.array([0 if nd.multipliername else -1 for nd in nodes], dtype=np.int32),
                br_bottleneck=np.zeros(len(nodes), dtype=np.int32))
    pd_ = (_lib.PedigreeDesc * 1)()
    d = pd_[0]
    d.n_leaves, d.n_branches, d.n_loci, d.n_fam, d.n_mult, d.n_asc = len(leaves), len(nodes), 0, 1, 1, 1
    for k, a in keep.items():
        setattr(d, k, _lib.iptr(a))
    md = _lib.ModelDesc()
    md.n_var, md.n_ped, md.peds = 1, 1, pd_
    out = np.zeros(4, dtype=np.int32)
    assert _lib.lib().mope_b200_schedule_stats(C.byref(md), F, 0, _lib.iptr(out)) == 0
    return tuple(int(v) for v in out)


def test_park_homes_of_the_pruning_schedule():
    """Parked messages: a cherry's first leaf waits in the (idle) chain buffer, interval scheduling fills the warp's
    tensor-memory slots, only the rest goes through the L2 scratch."""
    from mope_b200 import synth
    bal16 = synth.balanced_tree(16)
    # 15 parks per tile: 8 cherries -> chain buffer; F = 201 has one tensor-memory slot: the 4 depth-2 parks (disjoint
    # lifetimes) share it; 2 depth-1 parks + the root's stay in L2 (2 slots live at once)
    assert _schedule_stats(bal16, 201) == (8, 4, 3, 2)
    # F = 115: three tensor-memory slots hold everything else
    assert _schedule_stats(bal16, 115) == (8, 7, 0, 1)
    # the reference's example tree (chains of unary nodes below three forks): one park per fork, all in tensor memory at F = 115
    both = CASES["example_both"]["peds"][0]["tree"]
    forks = sum(1 for nd in newick.loads(both)[0].postorder() if len(nd.children) >= 2)
    assert sum(_schedule_stats(both, 201)[:3]) == forks
    assert _schedule_stats(both, 115)[2] == 0
    # a star tree folds every later sibling into one park; its lifetime holds only leaf branches
    assert _schedule_stats("(a:x,b:x,c:x,d:x)r;", 201) == (1, 0, 0, 1)
    # wide grids use the other kernel
    assert _schedule_stats(bal16, 1001) == (0, 0, 0, 0)

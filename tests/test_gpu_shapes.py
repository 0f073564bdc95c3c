"""GPU tests on the BASELINE.json workload shapes: oracle comparisons at sizes the oracle finishes in seconds, and
size-independent properties (row-permutation invariance, additivity over a family partition, determinism) at the full
cfg2 row count.  All through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

TB = H.load_tables()
RTOL = 1e-10


def _tables_small():
    from mope_b200 import TransitionTables
    return TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=float(TB["N"]), freqs=TB["freqs"],
                            breaks=TB["breaks"], bot=TB["bot"], nbs=TB["nbs"], bot_us=TB["bot_us"])


def _oracle_for(ds, tables, **kw):
    from oracle import mope_oracle as orc
    drift = orc.TransitionData(tables.drift, tables.gens, tables.us, int(tables.N), tables.freqs, tables.breaks)
    bot = orc.TransitionDataBottleneck(tables.bot, tables.nbs, tables.bot_us, int(tables.N)) if tables.bot is not None else None
    return orc.Inference([(H.read_tsv(ds.data_tsv), ds.tree, H.read_tsv(ds.ages_tsv))], drift, bot, log_unif_drift=True, **kw)


def _gpu_for(ds, tables, **kw):
    from mope_b200 import Inference
    return Inference([ds.data_tsv], [ds.tree], [ds.ages_tsv], tables, log_unif_drift=True, _inputs_are_text=True, **kw)


def _check_vs_oracle(ds, tables, n_walkers, seed=3, shrink=0.1, **kw):
    from mope_b200 import synth
    I = _gpu_for(ds, tables, **kw)
    try:
        O = _oracle_for(ds, tables, **kw)
        np.testing.assert_allclose(I.lower, O.lower, rtol=1e-15)
        np.testing.assert_allclose(I.upper, O.upper, rtol=1e-15)
        X = synth.walkers_in_prior_box(I.lower, I.upper, n_walkers, seed=seed, shrink=shrink)
        ll = I.loglike_batch(X)
        for i, x in enumerate(X):
            O.clear_cache()
            ref = O.loglike(x)
            if np.isfinite(ref):
                assert abs(ll[i] - ref) <= RTOL * abs(ref), (i, ll[i], ref)
            else:
                assert not np.isfinite(ll[i]) or ll[i] == ref
        return I, X, ll
    except Exception:
        I.close()
        raise


def test_cfg2_shape_fixed_branches_vs_oracle():
    from mope_b200 import synth
    ds = synth.make_dataset(640, 16, tree=synth.balanced_tree(16), seed=5)
    I, X, ll = _check_vs_oracle(ds, _tables_small(), 3)
    assert I.peds[0].n_asc == 1 and I.num_loci[0] == 640      # one distinct ascertainment pair without multipliers
    I.close()


def test_cfg3_cfg4_shape_rate_and_bottleneck_branches_vs_oracle():
    from mope_b200 import synth
    tree = synth.balanced_tree(8, rate_leaves=True, rate_internal=True, bottleneck=True)
    ds = synth.make_dataset(200, 8, tree=tree, seed=6)
    I, X, ll = _check_vs_oracle(ds, _tables_small(), 3, shrink=0.25)
    assert I.peds[0].n_asc == I.peds[0].n_fam                   # ages differ per family: no sharing
    # the kernel's own count of (warp, generation segment) pairs of the age-scaled branches: every warp with real rows sits
    # through at least one segment per age-scaled branch that is not identity-forwarded, and multiplies in at most those
    segs, active = I.engine.last_eval_segment_stats()
    assert 0 < active <= segs
    I.close()
    # fixed-length branches only: no segments
    J, _, _ = _check_vs_oracle(synth.make_dataset(64, 4, tree=synth.balanced_tree(4), seed=8), _tables_small(), 1)
    assert J.engine.last_eval_segment_stats() == (0, 0)
    J.close()
    # several frequency bins below min_freq / above 1 - min_freq: the ascertainment tiles take the leaf-branch messages
    # from multi-column low / high sums of the matrices (no GEMM), the mixed tile still multiplies
    I, X, ll = _check_vs_oracle(ds, _tables_small(), 2, shrink=0.25, min_freq=0.004)
    I.close()


def test_family_independent_ascertainment_subtrees_vs_oracle():
    """Fixed leaves under one age-scaled internal branch and one bottleneck branch (the cfg4 shape): on pure pseudo-row
    tiles the subtrees without an age-scaled branch are taken from the once-per-walker canonical launch.  Enough families
    for >= 8 such tiles; with MOPE_NO_ASC_CSE the same evaluation goes through the GEMMs and must agree to rounding."""
    import os
    from mope_b200 import synth
    tree = synth.balanced_tree(8, rate_internal=True, bottleneck=True)
    ds = synth.make_dataset(960, 8, tree=tree, seed=11)
    I, X, ll = _check_vs_oracle(ds, _tables_small(), 2, shrink=0.25)
    assert 2 * I.peds[0].n_asc >= 8 * 64
    I.close()
    os.environ["MOPE_NO_ASC_CSE"] = "1"
    try:
        I2 = _gpu_for(ds, _tables_small())
        ll2 = I2.loglike_batch(X)
        I2.close()
    finally:
        del os.environ["MOPE_NO_ASC_CSE"]
    assert (np.abs(ll2 - ll) / np.abs(ll)).max() < 1e-13


def test_empty_batch_and_single_locus_vs_oracle():
    """W = 0 walkers is an empty answer; one locus in one family (R = 3 rows, one ragged tile) matches the oracle."""
    from mope_b200 import synth
    ds = synth.make_dataset(1, 2, tree="(t0:a,t1:b*child_age)root;", seed=12)
    I, X, ll = _check_vs_oracle(ds, _tables_small(), 3, shrink=0.25)
    assert I.num_loci[0] <= 1
    empty = I.loglike_batch(np.zeros((0, I.ndim)))
    assert empty.shape == (0,)
    I.close()


def test_leaf_minimum_and_unary_chain_trees_vs_oracle():
    from mope_b200 import synth
    for tree in ("(t0:a,t1:b)root;", "((((t0:a)n1:b)n2:c*child_age)n3:d^,t1:a)root;", "(t0:a,t1:a,t2:b,t3:c*mother_age)root;",
                 # multifurcations below the root and several live parked messages: exercises the tensor-memory slot, the
                 # L2 slots and the fold-into-a-parked-message path of tree_kernel4 together
                 "((t0:a,t1:b,t2:c)n1:d,((t3:a,t4:b)n2:c,(t5:d,t6:a*mother_age,t7:b)n3:e)n4:a)root;"):
        nl = tree.count("t")
        ds = synth.make_dataset(70, nl, tree=tree, seed=7)
        # make_dataset names leaves t0..t{n-1}
        I, X, ll = _check_vs_oracle(ds, _tables_small(), 2, shrink=0.25)
        I.close()


def test_f201_and_f1001_multipass_vs_oracle():
    """F = 201 (bench grid size), F = 115 (the published default binning) and F = 1001 (cfg5: five N-passes per branch)
    on tables from the GPU generator; F = 201 also with the generic kernel instance (no scalar last column)."""
    from mope_b200 import generate, synth
    for brk, F, nloci, nleaves in ((0.005, 201, 60, 4), (0.01, 115, 60, 4), (1e-9, 1001, 12, 2)):
        tb = generate.generate_tables(N=1000, breaks=(0.5, brk), gens=[1, 10, 100, 1000], us=[1e-8, 1e-5, 1e-3], device="cuda")
        assert tb.F == F
        ds = synth.make_dataset(nloci, nleaves, tree=synth.balanced_tree(nleaves), seed=8)
        I, X, ll = _check_vs_oracle(ds, tb, 2, shrink=0.2)
        I.close()


def test_full_size_properties_cfg2_rows():
    """L = 10 000 loci x 16 leaves (cfg2 rows): permutation invariance, family-partition additivity, determinism."""
    from mope_b200 import synth
    tables = _tables_small()
    ds = synth.make_dataset(10000, 16, tree=synth.balanced_tree(16), seed=9)
    I = _gpu_for(ds, tables)
    X = synth.walkers_in_prior_box(I.lower, I.upper, 6, seed=4, shrink=0.15)
    ll = I.loglike_batch(X)
    assert np.isfinite(ll).all()
    # determinism: same call twice, and duplicated walkers inside one batch, are bit-identical
    np.testing.assert_array_equal(ll, I.loglike_batch(X))
    dup = I.loglike_batch(np.concatenate([X, X]))
    np.testing.assert_array_equal(dup[:6], dup[6:])
    np.testing.assert_array_equal(dup[:6], ll)
    I.close()
    # row permutation
    lines = ds.data_tsv.splitlines()
    perm = np.random.RandomState(1).permutation(len(lines) - 1)
    ds_p = synth.SyntheticPedigree("\n".join([lines[0]] + [lines[1 + i] for i in perm]) + "\n", ds.tree, ds.ages_tsv, ds.leaf_names,
                                   ds.n_loci, ds.n_fam)
    Ip = _gpu_for(ds_p, tables)
    llp = Ip.loglike_batch(X)
    Ip.close()
    assert (np.abs(llp - ll) / np.abs(ll)).max() < 1e-12
    # additivity over a family partition (what locus sharding relies on), incl. the global prior box
    from mope_b200 import dist as mdist
    parts = []
    for r in range(3):
        Il = mdist.build_locus_sharded([ds.data_tsv], [ds.tree], [ds.ages_tsv], tables, r, 3, log_unif_drift=True)
        np.testing.assert_array_equal(Il.lower, I.lower)
        np.testing.assert_array_equal(Il.upper, I.upper)
        parts.append(Il.loglike_batch(X))
        Il.close()
    tot = parts[0] + parts[1] + parts[2]
    assert (np.abs(tot - ll) / np.abs(ll)).max() < 1e-12


def test_locus_shards_with_rate_branches_sum_to_full():
    from mope_b200 import synth
    from mope_b200 import dist as mdist
    tables = _tables_small()
    tree = synth.balanced_tree(4, rate_leaves=True, rate_internal=True, bottleneck=True)
    ds = synth.make_dataset(900, 4, tree=tree, seed=10)
    I = _gpu_for(ds, tables)
    X = synth.walkers_in_prior_box(I.lower, I.upper, 5, seed=2, shrink=0.25)
    ll = I.loglike_batch(X)
    I.close()
    tot = np.zeros_like(ll)
    for r in range(2):
        Il = mdist.build_locus_sharded([ds.data_tsv], [ds.tree], [ds.ages_tsv], tables, r, 2, log_unif_drift=True)
        tot += Il.loglike_batch(X)
        Il.close()
    assert (np.abs(tot - ll) / np.abs(ll)).max() < 1e-12


def test_resident_evaluation_matches_host_planned():
    """mope_b200_eval_resident (plans and root prior derived by kernels from device-resident parameters) against the
    host-planned path on every kind of branch: fixed, age-scaled, bottleneck, identity short-cuts, walkers outside the
    tables, and walkers skipped through lnprior = -inf.  Same arithmetic up to the device's pow(): rtol 1e-12."""
    import torch
    from tests import helpers as H
    CASES = H.load_cases()
    OUT = H.load_outputs()
    TBL = H.load_tables()
    from mope_b200 import Inference, TransitionTables
    tb = TransitionTables(drift=TBL["drift"], gens=TBL["gens"], us=TBL["us"], N=float(TBL["N"]), freqs=TBL["freqs"],
                          breaks=TBL["breaks"], bot=TBL["bot"], nbs=TBL["nbs"], bot_us=TBL["bot_us"])
    for name in ("example_both", "two_pedigrees", "example_fixed"):
        case = CASES[name]
        peds = case["peds"]
        opts = dict(log_unif_drift=True)
        opts.update(case.get("opts", {}))
        I = Inference([p["data"] for p in peds], [p["tree"] for p in peds], [p["ages"] for p in peds], tb, _inputs_are_text=True, **opts)
        try:
            X = OUT[name + "/X"].copy()
            nv = I.num_varnames
            bad_mut = X[1].copy()
            bad_mut[nv] = 3.0                                   # mutation rate beyond the table
            bad_time = X[2].copy()
            bad_time[:nv] = 1.5                                 # drift beyond the table
            X = np.vstack([X, bad_mut, bad_time])
            W = X.shape[0]
            lnprior = np.zeros(W)
            lnprior[3] = -np.inf
            # reference: the host-planned path with the same (device) root prior, so that the comparison isolates the planner
            stat = I.engine.root_prior_device(X).clone()
            out = torch.empty(W, dtype=torch.float64, device=stat.device)
            st_host = I.engine.loglike_device(X, stat, out)
            ll_host = out.cpu().numpy()
            Xd = torch.as_tensor(X, device=stat.device)
            ll_res, st_res = I.engine.loglike_resident(Xd, lnprior_dev=torch.as_tensor(lnprior, device=stat.device))
            ll_res, st_res = ll_res.cpu().numpy(), st_res.cpu().numpy()
            assert st_res[3] == 128 and np.isnan(ll_res[3])
            st_res[3] = st_host[3]
            np.testing.assert_array_equal(st_res, st_host)
            assert st_host[-1] != 0 and st_host[-2] != 0
            ok = st_host == 0
            ok[3] = False
            assert np.all(np.isnan(ll_res[~ok]))
            np.testing.assert_allclose(ll_res[ok], ll_host[ok], rtol=1e-12)
            # a workspace that holds two walkers at a time gives the same numbers (chunked, static pool layout)
            mn, full = C.c_size_t(0), C.c_size_t(0)
            I.engine.lib.mope_b200_workspace_bytes(I.engine.ctx, 2, C.byref(mn), C.byref(full))
            small = torch.empty(int(full.value), dtype=torch.uint8, device=stat.device)
            out2 = torch.empty(W, dtype=torch.float64, device=stat.device)
            st2 = torch.empty(W, dtype=torch.int32, device=stat.device)
            rc = I.engine.lib.mope_b200_eval_resident(I.engine.ctx, Xd.data_ptr(), None, None, W, small.data_ptr(), small.numel(),
                                                      torch.cuda.current_stream().cuda_stream, out2.data_ptr(), st2.data_ptr())
            assert rc == 0
            got2 = out2.cpu().numpy()
            np.testing.assert_array_equal(got2[ok], ll_res[ok])
            assert np.isfinite(got2[3]) and abs(got2[3] - ll_host[3]) <= 1e-12 * abs(ll_host[3])   # not skipped without lnprior
            # and against the reference's golden values (auto root prior rules apply: the device prior is within 1e-9 everywhere)
            ref = OUT[name + "/loglike"]
            assert H.relerr(ll_res[:ref.shape[0]][ok[:ref.shape[0]]], ref[ok[:ref.shape[0]]]).max() <= 1e-9
        finally:
            I.close()

"""GPU parity tests: the CUDA path, called through the C ABI, against the reference-generated golden vectors
and against the CPU oracle on seeded inputs.  Bar: |ll_gpu - ll_ref| / |ll_ref| <= 1e-10 per walker (fp64)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

TB = H.load_tables()
CASES = H.load_cases()
OUT = H.load_outputs()
RTOL_LL = 1e-10


def _tables():
    from mope_b200 import TransitionTables
    return TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=float(TB["N"]), freqs=TB["freqs"],
                            breaks=TB["breaks"], bot=TB["bot"], nbs=TB["nbs"], bot_us=TB["bot_us"])


def _inference(case, **kw):
    from mope_b200 import Inference
    opts = dict(log_unif_drift=True)
    opts.update(case.get("opts", {}))
    opts.update(kw)
    peds = case["peds"]
    return Inference([p["data"] for p in peds], [p["tree"] for p in peds], [p["ages"] for p in peds], _tables(),
                     _inputs_are_text=True, **opts)


@pytest.fixture(scope="module")
def inf_both():
    I = _inference(CASES["example_both"])
    yield I
    I.close()


def test_transition_matrices_match_reference(inf_both):
    from mope_b200 import OutOfTableError
    E = inf_both.engine
    for t, m, P in zip(OUT["unit/drift_t"], OUT["unit/drift_m"], OUT["unit/drift_P"]):
        np.testing.assert_allclose(E.transition_matrix(t, m), P, rtol=1e-15, atol=0)
    for nb, m, P in zip(OUT["unit/bot_nb"], OUT["unit/bot_m"], OUT["unit/bot_P"]):
        np.testing.assert_allclose(E.transition_matrix(nb, m, bottleneck=True), P, rtol=1e-15, atol=0)
    with pytest.raises(OutOfTableError):
        E.transition_matrix(TB["gens"].max() / float(TB["N"]) * 1.01, 0.01)
    with pytest.raises(OutOfTableError):
        E.transition_matrix(0.1, 1e-30)
    with pytest.raises(OutOfTableError):
        E.transition_matrix(1.5, 0.01, bottleneck=True)
    with pytest.raises(OutOfTableError):
        E.transition_matrix(600, 0.01, bottleneck=True)


def test_leaf_likelihoods_match_reference(inf_both):
    E = inf_both.engine
    X, n = OUT["unit/binom_X"], OUT["unit/binom_n"]
    for key, phred in (("unit/binom_like", -1.0), ("unit/binom_like_phred20", 20.0)):
        got = E.leaf_likelihoods(X, n, TB["freqs"], phred)
        ref = OUT[key]
        # lnchoose is a difference of lgamma(n+1) ~ n log n values: element parity is bounded by a few ulp of those
        tol = 8 * np.finfo(np.float64).eps * np.maximum(n, 10)[:, :, None] * np.log(np.maximum(n, 10))[:, :, None]
        assert np.all(np.abs(got - ref) <= tol * ref + 1e-300)
        assert np.array_equal(got == 0, ref == 0)
        assert np.array_equal(got == 1, ref == 1)


def test_branch_updates_match_reference(inf_both):
    E = inf_both.engine
    child = OUT["unit/branch_child"]
    p = OUT["unit/branch_parent_in"].copy()
    E.branch_update(0, child, p, 0.37, 0.004)
    np.testing.assert_allclose(p, OUT["unit/branch_length_out"], rtol=1e-13)
    p = OUT["unit/branch_parent_in"].copy()
    E.branch_update(1, child, p, OUT["unit/branch_lengths"], 0.004)
    np.testing.assert_allclose(p, OUT["unit/branch_rate_out"], rtol=1e-13)
    p = OUT["unit/branch_parent_in"].copy()
    E.branch_update(2, child, p, 33.3, 0.004)
    np.testing.assert_allclose(p, OUT["unit/branch_bottleneck_out"], rtol=1e-13)


def test_branch_update_ragged_and_identity(inf_both):
    """Row counts around the 64-row tile, times below the first generation (identity short-cut) and exact grid hits,
    against the oracle."""
    from oracle import mope_oracle as orc
    E = inf_both.engine
    drift, bot = H.oracle_tables(TB)
    rng = np.random.RandomState(2)
    F = TB["freqs"].shape[0]
    N = float(TB["N"])
    for nrows in (1, 63, 64, 65, 130):
        child = rng.uniform(0, 1, (nrows, F))
        par = rng.uniform(0, 1, (nrows, F))
        lens = 10 ** rng.uniform(-4.5, np.log10(TB["gens"].max() / N), nrows)
        lens[0] = 1e-6                      # identity
        if nrows > 2:
            lens[1] = TB["gens"][0] / N     # exact hit of the first grid generation
            lens[2] = TB["gens"][4] / N     # exact interior grid hit
        ref = par.copy()
        drift.clear_cache()
        orc.compute_rate_transition_likelihood(child, ref, lens, 0.02, drift)
        got = par.copy()
        E.branch_update(1, child, got, lens, 0.02)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-300)
        ref = par.copy()
        orc.compute_length_transition_likelihood(child, ref, 1e-6, 0.02, drift)   # identity matrix
        got = par.copy()
        E.branch_update(0, child, got, 1e-6, 0.02)
        np.testing.assert_allclose(got, ref, rtol=1e-14)


def test_poisson_binomial_matches_reference(inf_both):
    E = inf_both.engine
    sets = [OUT["unit/pb_q_%d" % k] for k in range(4)]
    for k in range(4):
        got = E.non_asc_probs([sets[k]], TB["freqs"], float(OUT["unit/pb_minfreq_%d" % k]))[0]
        np.testing.assert_allclose(got, OUT["unit/pb_out_%d" % k], rtol=1e-12, atol=1e-300)
    # ragged batch in one call (same min_freq), incl. an empty read set
    got = E.non_asc_probs([sets[1], np.zeros(0), sets[0][:7]], TB["freqs"], 0.02)
    from oracle import mope_oracle as orc
    np.testing.assert_allclose(got[0], OUT["unit/pb_out_1"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(got[2], orc.get_non_asc_probs(sets[0][:7], TB["freqs"], 0.02), rtol=1e-12, atol=1e-300)
    assert np.all(got[1] == 0)


def test_poisson_binomial_any_depth(inf_both):
    """K = ceil(min_freq * nreads) beyond one register block of the DP (the reference handles any depth,
    _poisson_binom.pyx:18-90): block boundaries, deep read sets and a mixed ragged batch, against the oracle."""
    from oracle import mope_oracle as orc
    E = inf_both.engine
    rng = np.random.RandomState(17)
    freqs = TB["freqs"]
    for nr, mf in ((640, 0.05), (660, 0.05), (1280, 0.05), (1300, 0.05), (2000, 0.1), (3300, 0.01)):   # K = 32, 33, 64, 65, 200, 33
        q = rng.choice([20.0, 30.0, 40.0], size=nr, p=[0.1, 0.6, 0.3])
        got = E.non_asc_probs([q], freqs, mf)[0]
        np.testing.assert_allclose(got, orc.get_non_asc_probs(q, freqs, mf), rtol=1e-12, atol=1e-300)
    sets = [rng.choice([20.0, 30.0, 40.0], size=n) for n in (50, 1500, 0, 700, 3)]                    # K = 3, 75, 0, 35, 1
    got = E.non_asc_probs(sets, freqs, 0.05)
    for k, q in enumerate(sets):
        ref = orc.get_non_asc_probs(q, freqs, 0.05) if len(q) else np.zeros(freqs.shape[0])
        np.testing.assert_allclose(got[k], ref, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("name", sorted(CASES.keys()))
def test_full_likelihood_matches_reference(name):
    I = _inference(CASES[name])
    try:
        np.testing.assert_allclose(I.lower, OUT[name + "/lower"], rtol=1e-15)
        np.testing.assert_allclose(I.upper, OUT[name + "/upper"], rtol=1e-15)
        assert I.header == CASES[name]["header"]
        X = OUT[name + "/X"]
        ll, status, comps = I.loglike_components(X)
        assert not status.any()
        ref = OUT[name + "/loglike"]
        rel = H.relerr(ll, ref)
        assert rel.max() <= RTOL_LL, (name, rel.max())
        for k, c in enumerate(comps):
            assert H.relerr(c["root_ll"], OUT[name + "/root_ll"][:, k]).max() <= RTOL_LL
            np.testing.assert_allclose(c["log_asc"], OUT[name + "/log_asc_%d" % k], rtol=1e-9, atol=1e-14)
        # single-vector entry points and the posterior (sign folding + prior)
        assert H.relerr(I.loglike(X[0]), ref[0]) <= RTOL_LL
        nv = I.num_varnames
        Xs = X.copy()
        Xs[:, nv:2 * nv] *= np.where(np.arange(nv) % 2 == 0, -1.0, 1.0)
        Xs[:, -1] *= -1.0
        lp = I.log_posterior_batch(Xs)
        assert H.relerr(lp, OUT[name + "/log_posterior"]).max() <= RTOL_LL
        assert H.relerr(I.log_posterior(Xs[1]), OUT[name + "/log_posterior"][1]) <= RTOL_LL
        for i in (0, X.shape[0] - 1):
            assert H.relerr(I.logprior(X[i]), OUT[name + "/logprior"][i]) <= 1e-14
    finally:
        I.close()


def test_out_of_prior_and_out_of_table(inf_both):
    from mope_b200 import OutOfTableError
    I = inf_both
    x = OUT["example_both/X"][0].copy()
    bad = x.copy()
    bad[0] = I.upper[0] + 0.5
    assert I.log_posterior(bad) == -np.inf          # prior short-circuits, likelihood never evaluated
    far = x.copy()
    far[I.num_varnames] = 3.0                       # mutation rate beyond the table
    with pytest.raises(OutOfTableError):
        I.loglike(far)
    ll, status = I.loglike_batch(np.stack([x, far, x]), raise_on_error=False, return_status=True)
    assert status[0] == 0 and status[2] == 0 and status[1] != 0
    assert np.isnan(ll[1]) and ll[0] == ll[2]
    assert H.relerr(ll[0], OUT["example_both/loglike"][0]) <= RTOL_LL


def test_small_workspace_chunks_walkers():
    """A workspace that only fits a few walkers gives identical results (chunked evaluation)."""
    case = CASES["example_both"]
    X = OUT["example_both/X"][:24]
    I1 = _inference(case)
    ll_full = I1.loglike_batch(X)
    I1.close()
    I2 = _inference(case, workspace_bytes=1)        # clamps to the minimum: one walker per chunk
    ll_small = I2.loglike_batch(X)
    I2.close()
    np.testing.assert_array_equal(ll_full, ll_small)


def test_gpu_pool_map(inf_both):
    from mope_b200 import GpuPool
    from mope_b200 import inference as minf
    minf.inf_data = inf_both
    pool = GpuPool(inf_both)
    X = OUT["example_both/X"][:5]
    got = pool.map(minf.post_clone, list(X))
    assert H.relerr(got, [inf_both.log_posterior(x) for x in X]).max() == 0
    got = pool.map(minf.logl, list(X))
    assert H.relerr(got, OUT["example_both/loglike"][:5]).max() <= RTOL_LL
    assert pool.map(minf.logp, list(X)) == [inf_both.logprior(x) for x in X]
    assert pool.map(lambda v: float(v.sum()), list(X)) == [float(x.sum()) for x in X]


def test_run_mcmc_chain_file_and_resume(tmp_path):
    """Inference.run_mcmc: chain file format of the reference (print_csv_lines / translate_positions), the printed
    log-posteriors belong to the printed positions, and --prev-chain style resume."""
    import io
    I = _inference(CASES["test_rate_two_loci"])
    try:
        nw, nit = 2 * I.ndim + 2, 6
        buf = io.StringIO()
        sampler = I.run_mcmc(nit, nw, chain_alpha=1.4, seed=11, file=buf)
        lines = buf.getvalue().splitlines()
        assert lines[0].startswith("# ") and lines[1] == I.header and lines[-1].startswith("# mean acceptance across chains:")
        body = [l for l in lines[2:-1]]
        assert len(body) == nw * nit
        vals = np.array([[float(t) for t in l.split("\t")] for l in body])
        assert vals.shape == (nw * nit, I.ndim + 1)
        nv = I.num_varnames
        # zero drift <=> NaN mutation rate; otherwise natural-scale values inside the prior box
        zero = vals[:, 1:1 + nv] == 0
        assert np.array_equal(zero, np.isnan(vals[:, 1 + nv:1 + 2 * nv]))
        assert np.all(np.isfinite(vals[:, 0]))
        # last ensemble: recompute the log-posterior of the non-degenerate walkers from the printed values
        last = vals[-nw:]
        ok = ~zero[-nw:].any(axis=1)
        if ok.any():
            x = last[ok, 1:].copy()
            x[:, :2 * nv] = np.log10(x[:, :2 * nv])
            lp = I.log_posterior_batch(x)
            assert (np.abs(lp - last[ok, 0]) / np.abs(last[ok, 0])).max() < 1e-9
        assert 0.0 <= np.mean(sampler.acceptance_fraction) <= 1.0
        # resume from the written chain
        fn = tmp_path / "chain.txt"
        fn.write_text(buf.getvalue())
        buf2 = io.StringIO()
        I.run_mcmc(2, nw, prev_chain=str(fn), chain_alpha=1.4, seed=5, file=buf2)
        assert len([l for l in buf2.getvalue().splitlines() if not l.startswith("#")]) == 1 + 2 * nw
    finally:
        I.close()


def test_device_root_prior(inf_both):
    """mope_b200_root_prior against the reference's stationary distributions (golden, absolute 1e-13: the reference's own
    cell masses carry that much rounding noise at small ab) and against an exact evaluation (mpmath, relative 1e-13)."""
    I = inf_both
    abpp = OUT["unit/stat_abpp"]
    X = np.zeros((abpp.shape[0], I.ndim))
    X[:, -2:] = np.log10(abpp)
    got = I.engine.root_prior_device(X).cpu().numpy()
    ref = OUT["unit/stat_dist"]
    assert np.abs(got - ref).max() <= 1e-13
    assert np.abs(got.sum(1) - 1).max() <= 1e-14
    # large ab: the reference's differences are accurate there, so the relative agreement is tight as well
    big = abpp[:, 0] >= 1e-2
    np.testing.assert_allclose(got[big], ref[big], rtol=1e-11)
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    N = int(TB["N"])
    Ni = N - 2
    edges = np.arange(1, 2 * Ni, 2) / (2 * Ni)
    for row, (a, pp) in zip(got, abpp):
        cdf = [mp.betainc(float(a), float(a), 0, mp.mpf(float(x)), regularized=True) for x in edges]
        cells = np.array([float(h - l) for h, l in zip(cdf + [mp.mpf(1)], [mp.mpf(0)] + cdf)])
        d = np.zeros(N + 1)
        d[0] = d[-1] = (1 - pp) / 2
        d[1:-1] = cells * pp
        exact = np.add.reduceat(d, TB["breaks"])
        np.testing.assert_allclose(row, exact, rtol=1e-13, atol=0)


def test_root_prior_modes_agree_within_parity(inf_both):
    """auto (device kernel above ab = 1e-6, SciPy below) is the default; 'scipy' reproduces round 1; 'device' is the
    all-device mode of the resident sampler.  All three within the parity bar of the reference on the golden vectors,
    which span log10 ab from -8.5 to -0.5."""
    I = inf_both
    X = OUT["example_both/X"]
    ref = OUT["example_both/loglike"]
    assert I.root_prior == "auto"
    mask = I.root_prior_host_mask(X)
    assert mask.any() and (~mask).any()            # both branches exercised
    try:
        for mode, tol in (("auto", 1e-10), ("scipy", 1e-10), ("device", 1e-9)):
            I.root_prior = mode
            ll = I.loglike_batch(X)
            assert H.relerr(ll, ref).max() <= tol, (mode, H.relerr(ll, ref).max())
    finally:
        I.root_prior = "auto"


def test_device_sampler_matches_host_sampler():
    """The device-resident stretch move (mope_b200_mcmc_halfstep) against the host EnsembleSampler fed the same uniform
    stream: same proposals, same accept/reject decisions, hence identical positions; log-posteriors within the parity bar.
    Then the counter-based generator: reproducible under a seed, sensible acceptance, no host evaluation calls."""
    import time
    from mope_b200.ensemble import DeviceEnsembleSampler, EnsembleSampler
    I = _inference(CASES["example_both"])
    try:
        I.root_prior = "device"                    # the host sampler evaluates with the same root prior as the resident step
        nw, nit = 2 * I.ndim + 4, 6
        p0 = np.random.RandomState(9).uniform(0.7 * I.lower + 0.3 * I.upper, 0.3 * I.lower + 0.7 * I.upper, (nw, I.ndim))
        # flip signs on the folded coordinates of a few walkers: positions are stored unfolded
        p0[::3, I.num_varnames:2 * I.num_varnames] *= -1.0
        host = EnsembleSampler(nw, I.ndim, I.log_posterior_batch, a=1.4, seed=77, randomize_split=False)
        hchain = [(p, lp) for p, lp, _ in host.sample(p0, iterations=nit)]
        dev = DeviceEnsembleSampler(I, nw, a=1.4, streams=np.random.RandomState(77))
        calls = []
        real = I.log_posterior_batch
        I.log_posterior_batch = lambda X: (calls.append(np.shape(X)), real(X))[1]
        dchain = [(p, lp) for p, lp, _ in dev.sample(p0, iterations=nit)]
        I.log_posterior_batch = real
        assert calls == [(nw, I.ndim)]             # only the initial ensemble goes through the host API
        moved = 0
        for (ph, lph), (pd_, lpd) in zip(hchain, dchain):
            np.testing.assert_array_equal(ph, pd_)
            assert H.relerr(lpd, lph).max() <= 1e-10
            moved += int((ph != p0).any(axis=1).sum())
        assert moved > 0
        np.testing.assert_array_equal(dev.acceptance_fraction, host.acceptance_fraction)
        assert int(dev.status_or.max().item()) == 0                 # no in-prior proposal left the tables
        # Philox streams: deterministic under a seed, different under another
        a = [p for p, _, _ in DeviceEnsembleSampler(I, nw, a=1.4, seed=5).sample(p0, iterations=4)][-1]
        b = [p for p, _, _ in DeviceEnsembleSampler(I, nw, a=1.4, seed=5).sample(p0, iterations=4)][-1]
        c = [p for p, _, _ in DeviceEnsembleSampler(I, nw, a=1.4, seed=6).sample(p0, iterations=4)][-1]
        np.testing.assert_array_equal(a, b)
        assert (a != c).any()
        s = DeviceEnsembleSampler(I, nw, a=1.4, seed=11)
        last = None
        for p, lp, _ in s.sample(p0, iterations=40):
            last = (p, lp)
        acc = s.acceptance_fraction.mean()
        assert 0.05 < acc < 0.95, acc
        # the stored log-posteriors belong to the stored positions
        chk = I.log_posterior_batch(last[0])
        assert H.relerr(last[1], chk).max() <= 1e-10
        # host time per half-step of the resident sampler (enqueue only)
        s2 = DeviceEnsembleSampler(I, nw, a=1.4, seed=12)
        list(s2.sample(p0, iterations=3))
        ws = I.engine._workspace(nw // 2)
        st = s2._state(1000)
        import ctypes as C
        import torch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 50
        for k in range(n):
            st.step = 1000 + k
            assert I.engine.lib.mope_b200_mcmc_halfstep(I.engine.ctx, C.byref(st), k & 1, ws.data_ptr(), ws.numel(),
                                                        torch.cuda.current_stream().cuda_stream) == 0
        host_us = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
        print("host time per resident half-step: %.1f us" % host_us)
        # full chain through Inference.run_mcmc(device_sampler=True): reference chain-file format
        import io
        buf = io.StringIO()
        I.run_mcmc(3, nw, chain_alpha=1.4, seed=3, file=buf, device_sampler=True, init_pos=p0)
        body = [l for l in buf.getvalue().splitlines() if not l.startswith("#")]
        assert body[0] == I.header and len(body) == 1 + 3 * nw
    finally:
        I.close()

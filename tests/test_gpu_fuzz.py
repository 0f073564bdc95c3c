"""Seeded random pedigrees against the oracle: random topologies (binary and multifurcating nodes, unary chains), parameter
names shared at random between branches, age-scaled branches on any of the three multiplier columns, bottleneck branches,
ragged locus counts, with and without the ascertainment short-cuts.  Bar: 1e-10 relative per walker, through the C ABI."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_shapes import _check_vs_oracle, _tables_small

pytestmark = pytest.mark.gpu

MULTS = ["mother_age", "child_age", "mother_birth_age"]


def random_tree(rng, n_leaves):
    """Newick text in the reference's dialect (name:var, name:var*multiplier, name:var^) over leaves t0..t<n-1>."""
    n_plain, n_bott = rng.randint(2, 5), rng.randint(1, 3)
    counter = [0]

    def label(name):
        kind = rng.choice(["fixed", "rate", "bott"], p=[0.5, 0.35, 0.15])
        if kind == "bott":
            return "%s:b%d^" % (name, rng.randint(n_bott))
        var = "v%d" % rng.randint(n_plain)
        if kind == "rate":
            return "%s:%s*%s" % (name, var, MULTS[rng.randint(3)])
        return "%s:%s" % (name, var)

    nodes = ["t%d" % i for i in range(n_leaves)]     # subtree texts without their own branch label
    rng.shuffle(nodes)
    nodes = [label(n) for n in nodes]
    while len(nodes) > 1:
        k = min(len(nodes), rng.choice([1, 2, 2, 2, 3]))
        if k == 1 and len(nodes) > 1 and rng.rand() < 0.5:
            k = 2
        picks = [nodes.pop(rng.randint(len(nodes))) for _ in range(k)]
        counter[0] += 1
        name = "n%d" % counter[0]
        if len(nodes) == 0 and k > 1:
            return "(%s)root;" % ",".join(picks)
        nodes.append(label("(%s)%s" % (",".join(picks), name)))
    return "(%s)root;" % nodes[0]


def _tables_f115():
    from mope_b200 import TransitionTables
    tb = H.load_tables("tables_f115.npz")
    return TransitionTables(drift=tb["drift"], gens=tb["gens"], us=tb["us"], N=float(tb["N"]), freqs=tb["freqs"], breaks=tb["breaks"],
                            bot=tb["bot"], nbs=tb["nbs"], bot_us=tb["bot_us"])


@pytest.mark.parametrize("seed", range(24))
def test_random_pedigree_vs_oracle(seed):
    from mope_b200 import synth
    rng = np.random.RandomState(1000 + seed)
    n_leaves = int(rng.randint(2, 9))
    tree = random_tree(rng, n_leaves)
    n_loci = int(rng.choice([1, 7, 63, 64, 65, 130, 200]))
    ds = synth.make_dataset(n_loci, n_leaves, tree=tree, seed=50 + seed, loci_per_family=int(rng.randint(1, 5)),
                            missing_frac=float(rng.choice([0.0, 0.1, 0.3])))
    kw = {}
    if seed % 3 == 1:
        kw = dict(min_freq=0.004, poisson_like_penalty=0.5)
    if seed % 3 == 2:
        kw = dict(min_phred_score=30.0, genome_size=5000)
    tables = _tables_f115() if seed % 4 == 3 else _tables_small()     # F = 115 (another kernel instance) for a quarter of the cases
    I, X, ll = _check_vs_oracle(ds, tables, 2, seed=seed, shrink=0.3, **kw)
    try:
        assert np.isfinite(ll).any(), tree
        # the device-planned, device-prior evaluation of the same vectors (what the resident sampler runs)
        import torch
        Xd = torch.as_tensor(X, device="cuda:%d" % I.engine.device)
        ll_res, st_res = I.engine.loglike_resident(Xd)
        ll_res, st_res = ll_res.cpu().numpy(), st_res.cpu().numpy()
        ok = np.isfinite(ll)
        assert np.array_equal(st_res == 0, ok)
        assert H.relerr(ll_res[ok], ll[ok]).max() <= 1e-9, (tree, ll_res, ll)
    finally:
        I.close()


def _random_tree_no_bottleneck(rng, n_leaves):
    while True:
        t = random_tree(rng, n_leaves)
        if "^" not in t:
            return t


@pytest.mark.parametrize("seed", range(6))
def test_random_pedigree_selection_model_vs_oracle(seed):
    """The selection model on random pedigrees (no bottleneck branches: the reference's selection path reads every branch
    from the drift tables): random positions shared between families, coefficients of both signs."""
    import io
    import warnings
    from mope_b200 import Inference, TransitionTables, synth
    from oracle import mope_oracle as orc
    from oracle import mope_oracle_selection as osel
    rng = np.random.RandomState(2000 + seed)
    n_leaves = int(rng.randint(2, 7))
    tree = _random_tree_no_bottleneck(rng, n_leaves)
    n_loci = int(rng.choice([5, 40, 90]))
    ds = synth.make_dataset(n_loci, n_leaves, tree=tree, seed=70 + seed, loci_per_family=int(rng.randint(1, 4)), missing_frac=0.1)
    df = ds.data.copy()
    df["position"] = 100 + 7 * rng.randint(max(2, n_loci // 3), size=n_loci)
    buf = io.StringIO()
    df.to_csv(buf, sep="\t", na_rep="NA", index=False)
    data_tsv = buf.getvalue()
    tb = H.load_tables("tables_sel.npz")
    tables = TransitionTables(drift=tb["drift"], gens=tb["gens"], us=tb["us"], N=float(tb["N"]), freqs=tb["freqs"], breaks=tb["breaks"])
    I = Inference([data_tsv], [tree], [ds.ages_tsv], tables, selection_model=True, log_unif_drift=True, _inputs_are_text=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        drift = orc.TransitionData(tb["drift"], tb["gens"], tb["us"], int(tb["N"]), tb["freqs"], tb["breaks"])
        O = osel.SelectionInference([(H.read_tsv(data_tsv), tree, H.read_tsv(ds.ages_tsv))], drift, log_unif_drift=True)
    try:
        assert I.ndim == O.ndim
        np.testing.assert_allclose(I.lower, O.lower, rtol=1e-14)
        npos, nv = I.num_unique_positions, I.num_varnames
        for _ in range(2):
            x = rng.uniform(0.7 * I.lower + 0.3 * I.upper, 0.3 * I.lower + 0.7 * I.upper)
            x[nv:2 * nv] = rng.uniform(-0.2, 0.2, nv)
            x[-npos:] = rng.choice([-1.0, 1.0], npos) * (10.0 + rng.uniform(0, 4, npos)) * (rng.rand(npos) < 0.7)
            O.clear_cache()
            ref = O.loglike(x.copy())
            got = I.loglike(x.copy())
            assert np.isfinite(ref) and abs(got - ref) <= 1e-10 * abs(ref), (tree, got, ref)
    finally:
        I.close()

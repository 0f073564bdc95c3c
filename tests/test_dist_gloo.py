"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in mope_b200.dist: family partition, the all-reduce of
per-walker partial sums (locus sharding) and the all-gather of walker slices.  The oracle stands in for the device
evaluation (test infrastructure only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from mope_b200 import dist as mdist
from tests import helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _OracleAsInference:
    """Duck-typed stand-in for mope_b200.Inference.loglike_batch on CPU."""

    def __init__(self, orc_inf):
        self.o = orc_inf

    def loglike_batch(self, X, raise_on_error=True, return_status=False):
        ll = np.array([self.o.loglike(np.array(x)) for x in X])
        st = np.zeros(len(X), dtype=np.int32)
        return (ll, st) if return_status else ll


def _worker(rank, world, port, case_name, nvec, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        tdist.init_process_group("gloo", rank=rank, world_size=world)
        tb = H.load_tables()
        case = H.load_cases()[case_name]
        out = H.load_outputs()
        X = out[case_name + "/X"][:nvec]
        # --- locus sharding: this rank's families, all walkers, one all-reduce
        local_case = dict(case)
        local_case["peds"] = []
        for ped in case["peds"]:
            d, a = mdist.shard_pedigree_text(ped["data"], ped["ages"], world, rank)
            local_case["peds"].append(dict(data=d, tree=ped["tree"], ages=a))
        inf_local = _OracleAsInference(H.oracle_inference(local_case, tb))
        calls = []
        real_all_reduce = tdist.all_reduce
        tdist.all_reduce = lambda *a, **k: (calls.append(1), real_all_reduce(*a, **k))[1]
        try:
            total, st = mdist.loglike_locus_sharded(inf_local, X)
        finally:
            tdist.all_reduce = real_all_reduce
        assert len(calls) == 1, "the locus-sharded evaluation must issue exactly one collective, saw %d" % len(calls)
        ref = out[case_name + "/loglike"][:nvec]
        rel = np.abs(total - ref) / np.abs(ref)
        assert rel.max() < 1e-11, rel
        assert not st.any()
        # --- walker sharding: full data, this rank's slice, all-gather
        inf_full = _OracleAsInference(H.oracle_inference(case, tb))
        full = mdist.loglike_walker_sharded(inf_full, X)
        assert (np.abs(full - ref) / np.abs(ref)).max() < 1e-11
        # status bits are OR-ed and turn the walker into NaN on every rank
        ll = np.arange(4, dtype=np.float64) + rank
        st = np.array([0, 1 if rank == 0 else 0, 4 if rank == 1 else 0, 0], dtype=np.int32)
        tot, sst = mdist.allreduce_partials(ll, st)
        assert list(sst) == [0, 1, 4, 0] and np.isnan(tot[1]) and np.isnan(tot[2]) and tot[0] == 1.0 and tot[3] == 7.0
        # a NaN partial of a flagged walker must not poison the sum of the others; every status bit survives the packing
        ll2 = np.array([np.nan if rank == 0 else 5.0, 2.0])
        st2 = np.array([(64 | 32 | 2) if rank == 0 else 16, 0], dtype=np.int32)
        tot2, sst2 = mdist.allreduce_partials(ll2, st2)
        assert list(sst2) == [64 | 32 | 16 | 2, 0] and np.isnan(tot2[0]) and tot2[1] == 4.0
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        if tdist.is_initialized():
            tdist.destroy_process_group()


@pytest.mark.parametrize("case_name,nvec", [("example_both", 2), ("two_pedigrees", 2)])
def test_locus_and_walker_sharding_world2(case_name, nvec):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case_name, nvec, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_partition_covers_every_family_once():
    case = H.load_cases()["example_both"]["peds"][0]
    from mope_b200 import model
    data = model.read_table(case["data"], True)
    ages = model.read_table(case["ages"], True)
    for world in (1, 2, 3, 8):
        groups = mdist.partition_families(data["family"].values, ages["family"].values, world)
        allf = np.concatenate(groups)
        assert sorted(allf.tolist()) == sorted(ages["family"].values.tolist())
        assert len(allf) == len(set(allf.tolist()))
        rows = [sum((data["family"].values == f).sum() + 2 for f in g) for g in groups]
        assert max(rows) - min(rows) <= 0.5 * max(rows) + 8
    for n, world in ((10, 3), (256, 8), (5, 8)):
        sl = [mdist.walker_slice(n, world, r) for r in range(world)]
        assert sl[0].start == 0 and sl[-1].stop == n and all(a.stop == b.start for a, b in zip(sl, sl[1:]))


def test_global_mult_bounds_match_full_inference():
    case = H.load_cases()["example_both"]
    tb = H.load_tables()
    names, lo, hi = mdist.global_mult_bounds([p["tree"] for p in case["peds"]], [p["ages"] for p in case["peds"]])
    o = H.oracle_inference(case, tb)
    assert names == o.varnames
    mn, mx = o._min_max_mults()
    np.testing.assert_array_equal(lo, mn)
    np.testing.assert_array_equal(hi, mx)

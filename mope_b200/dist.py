"""Multi-GPU sharding of the log-likelihood: one process per GPU, torch.distributed for the plumbing.

The path shards two ways, both without any exchange inside the pruning pass:

* by WALKER -- what the reference's multiprocessing/MPI pool does (mope/inference.py:1194-1244): every rank holds
  the tables and the whole dataset and evaluates its slice of the ensemble.  No data-path collective; an all-gather
  of the per-walker results only if every rank needs the whole vector.
* by LOCUS (large datasets) -- rows are independent given a walker's parameters.  Families are partitioned over the
  ranks (a family's loci, its ascertainment pseudo-rows and its count/Poisson terms stay together), every rank
  evaluates ALL walkers on its rows, and the per-walker partial sums (with the walkers' status bits packed behind them)
  are combined with ONE all-reduce(sum) (NCCL over NVLink on GPUs, gloo in the CPU tests).

The prior box depends on the multiplier (age) range of ALL families (inference.py:154-258), so locus shards are
built with the global range (`mult_bounds`).
"""
from __future__ import annotations

import io
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import model as _model


def walker_slice(n_walkers: int, world: int, rank: int) -> slice:
    """Contiguous, balanced slice of the ensemble owned by `rank`."""
    bounds = np.linspace(0, n_walkers, world + 1).astype(int)
    return slice(int(bounds[rank]), int(bounds[rank + 1]))


def partition_families(family_of_locus: np.ndarray, families: np.ndarray, world: int) -> List[np.ndarray]:
    """Split the family list into `world` contiguous groups with balanced row counts (loci + 2 ascertainment rows).
    Every family of the ages table lands in exactly one group -- also families without any locus, whose Poisson
    term still counts (inference.py:1113-1129)."""
    import pandas as pd
    rows = pd.Series(family_of_locus).value_counts().reindex(families, fill_value=0).values.astype(np.float64) + 2.0
    cum = np.cumsum(rows)
    total = cum[-1] if len(cum) else 0.0
    groups, start = [], 0
    for r in range(world):
        target = total * (r + 1) / world
        end = int(np.searchsorted(cum, target - 1e-9, side="left")) + 1 if r < world - 1 else len(families)
        end = max(end, start)
        end = min(end, len(families))
        groups.append(np.asarray(families[start:end]))
        start = end
    return groups


def shard_pedigree(data, ages, world: int, rank: int):
    """The rows of one pedigree's data / ages tables (DataFrames, or TSV text) that belong to `rank` under the family
    partition, as DataFrames."""
    data = _model.read_table(data, True)
    ages = _model.read_table(ages, True)
    groups = partition_families(data["family"].values, ages["family"].values, world)
    if any(len(g) == 0 for g in groups):
        # every rank computes the same partition, so every rank raises here -- before any collective could hang
        raise ValueError("cannot shard {} families over {} ranks: some rank would own no family".format(
            len(ages["family"].values), world))
    mine = groups[rank]
    d = data[data["family"].isin(mine)].reset_index(drop=True)
    a = ages[ages["family"].isin(mine)].reset_index(drop=True)
    return d, a


def shard_pedigree_text(data_tsv: str, ages_tsv: str, world: int, rank: int) -> Tuple[str, str]:
    """shard_pedigree with TSV text out (the reference's file format)."""
    d, a = shard_pedigree(data_tsv, ages_tsv, world, rank)
    dbuf, abuf = io.StringIO(), io.StringIO()
    d.to_csv(dbuf, sep="\t", index=False, na_rep="NA")
    a.to_csv(abuf, sep="\t", index=False, na_rep="NA")
    return dbuf.getvalue(), abuf.getvalue()


def global_mult_bounds(tree_strs: Sequence[str], ages_tsvs: Sequence[str]):
    """min/max multiplier per variable name over ALL families of all pedigrees (Inference.get_min_max_mults)."""
    from . import newick
    mins, maxes = {}, {}
    for tree_str, ages_tsv in zip(tree_strs, ages_tsvs):
        ages = _model.read_table(ages_tsv, True)
        tree = newick.loads(tree_str)[0]
        for nd in tree.postorder():
            if nd is tree:
                continue
            v, m = nd.varname, nd.multipliername
            lo, hi = (np.nanmin(ages[m].values), np.nanmax(ages[m].values)) if m else (1, 1)
            mins.setdefault(v, []).append(lo)
            maxes.setdefault(v, []).append(hi)
    names = sorted(mins)
    return names, np.array([min(mins[v]) for v in names]), np.array([max(maxes[v]) for v in names])


def build_locus_sharded(data_tsvs: Sequence[str], tree_strs: Sequence[str], ages_tsvs: Sequence[str], tables, rank: int, world: int,
                        **kw):
    """An Inference over this rank's families (TSV texts or DataFrames in), with the GLOBAL prior box."""
    from .inference import Inference
    _names, lo, hi = global_mult_bounds(tree_strs, ages_tsvs)
    local = [shard_pedigree(d, a, world, rank) for d, a in zip(data_tsvs, ages_tsvs)]
    return Inference([l[0] for l in local], list(tree_strs), [l[1] for l in local], tables, _inputs_are_text=True,
                     mult_bounds=(lo, hi), **kw)


def _dist_device(group=None):
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


N_STATUS_BITS = 7   # MOPE_ST_* bits of include/mope_b200.h
_PINNED, _PINNED_BUSY = {}, {}   # pinned staging of the status planes, by size


def allreduce_partials_packed(ll_t, status: Optional[np.ndarray], group=None, packed=None):
    """The path's only collective: ONE all-reduce(sum) over the ranks of a packed float64 buffer
    [ W partial log-likelihoods | N_STATUS_BITS x W status bit planes ].  A walker's partial sum enters as 0 where this
    rank flagged it; a bit plane sums to the number of ranks that set the bit (exact in float64), so the OR of the status
    words falls out of the same reduction.  `ll_t` is a float64 tensor on the collective's device (CUDA for NCCL: the
    reduction is enqueued behind the evaluation on the current stream, no host round trip before it).  Returns host arrays
    (ll with NaN where any rank flagged the walker, OR-ed status)."""
    import torch
    import torch.distributed as dist
    W = ll_t.shape[0]
    n = W * (1 + N_STATUS_BITS)
    if packed is None or packed.shape[0] < n or packed.device != ll_t.device:
        packed = torch.empty(n, dtype=torch.float64, device=ll_t.device)
    buf = packed[:n]
    st = np.zeros(W, dtype=np.int32) if status is None else np.ascontiguousarray(status, dtype=np.int32)
    planes = ((st[None, :] >> np.arange(N_STATUS_BITS, dtype=np.int32)[:, None]) & 1).astype(np.float64)   # [bits, W]
    if ll_t.is_cuda:
        pin = _PINNED.get(n - W)
        if pin is None:
            pin = _PINNED[n - W] = torch.empty(n - W, dtype=torch.float64, pin_memory=True)
        else:
            torch.cuda.current_stream().synchronize() if _PINNED_BUSY.get(n - W) else None
        pin.numpy()[:] = planes.reshape(-1)
        _PINNED_BUSY[n - W] = True
        buf[W:].copy_(pin, non_blocking=True)
    else:
        buf[W:].copy_(torch.from_numpy(planes.reshape(-1)))
    flagged = buf[W:].view(N_STATUS_BITS, W).sum(0) != 0
    buf[:W].copy_(torch.where(flagged, torch.zeros_like(ll_t), ll_t))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    host = buf.cpu().numpy()      # synchronises: the pinned staging buffer is free again
    _PINNED_BUSY[n - W] = False
    counts = host[W:].reshape(N_STATUS_BITS, W)
    st_or = ((counts != 0).astype(np.int32) << np.arange(N_STATUS_BITS, dtype=np.int32)[:, None]).sum(0).astype(np.int32)
    out = np.where(st_or != 0, np.nan, host[:W])
    return out, st_or


def allreduce_partials(ll_partial: np.ndarray, status: Optional[np.ndarray] = None, group=None):
    """Host-array front end of allreduce_partials_packed (one collective)."""
    import torch
    dev = _dist_device(group)
    t = torch.as_tensor(np.ascontiguousarray(ll_partial, dtype=np.float64)).to(dev)
    out, st = allreduce_partials_packed(t, status, group)
    return out if status is None else (out, st)


def allgather_walkers(ll_local: np.ndarray, n_walkers: int, group=None) -> np.ndarray:
    """Walker sharding: concatenate the ranks' slices (padded all-gather; slices follow walker_slice)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = _dist_device(group)
    width = max(walker_slice(n_walkers, world, r).stop - walker_slice(n_walkers, world, r).start for r in range(world))
    buf = torch.zeros(width, dtype=torch.float64, device=dev)
    buf[: ll_local.shape[0]] = torch.as_tensor(np.ascontiguousarray(ll_local, dtype=np.float64)).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.empty(n_walkers)
    for r in range(world):
        sl = walker_slice(n_walkers, world, r)
        out[sl] = parts[r][: sl.stop - sl.start].cpu().numpy()
    return out


def loglike_walker_sharded(inf, X: np.ndarray, group=None) -> np.ndarray:
    """Every rank evaluates its slice of X; all ranks return the full vector."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sl = walker_slice(X.shape[0], world, rank)
    local = inf.loglike_batch(X[sl], raise_on_error=False) if sl.stop > sl.start else np.zeros(0)
    return allgather_walkers(local, X.shape[0], group)


def loglike_locus_sharded(inf_local, X: np.ndarray, group=None):
    """`inf_local` holds this rank's families (build_locus_sharded); returns (log-likelihood of every walker on the full
    data, OR-ed status) on all ranks.  Device path: the partial sums stay in HBM between the evaluation and the all-reduce;
    the only host<->device traffic is the parameter vectors / root prior in and one read of the reduced buffer."""
    eng = getattr(inf_local, "engine", None)
    if eng is not None and hasattr(eng, "loglike_enqueue"):
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        if hasattr(inf_local, "root_prior_host_mask"):
            ll_t, status = eng.loglike_hybrid(X, inf_local.root_prior_host_mask(X), inf_local.stationary_distributions, keep_on_device=True)
        else:
            ll_t, status = eng.loglike_enqueue(X, inf_local.stationary_distributions)
        packed = getattr(eng, "_packed_reduce_buf", None)
        n = X.shape[0] * (1 + N_STATUS_BITS)
        if packed is None or packed.shape[0] < n:
            packed = eng._packed_reduce_buf = eng.torch.empty(n, dtype=eng.torch.float64, device=ll_t.device)
        return allreduce_partials_packed(ll_t, status, group, packed)
    ll, status = inf_local.loglike_batch(X, raise_on_error=False, return_status=True)
    ll = np.where(status != 0, 0.0, ll)
    return allreduce_partials(ll, status, group)

#!/usr/bin/env python
"""Benchmark of the hot path: walker x locus log-likelihoods per second (BASELINE.json metric).

Workload (configs[1] of BASELINE.json, "cfg2" in SURVEY.md section 8d): synthetic heteroplasmy data, L = 10 000
loci x S = 16 tissues, balanced 16-leaf tree (30 fixed-length branches, 8 parameter names), F = 201 frequency
bins (WF N = 1000, --breaks 0.5 0.005), the reference's default 94 x 44 drift grid generated locally, W = 256
walkers drawn uniformly in the prior box.  A step = one evaluation of Inference.loglike for all W walkers.
Multi-GPU: walkers are sharded, W = 256 per GPU (weak scaling), no data-path collective.

    python bench.py [--gpus N --steps K --warmup W]          our arm (one JSON line on rank 0)
    python bench.py --impl reference [...]                   the reference's own CPU implementation (baseline/_ref
                                                             if installed, else the oracle port), same metric
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "walker_x_locus_loglikelihoods_per_sec"
UNIT = "walker*locus/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # workload overrides (defaults are the BASELINE config)
    ap.add_argument("--loci", type=int, default=10000)
    ap.add_argument("--leaves", type=int, default=16)
    ap.add_argument("--walkers", type=int, default=256, help="walkers per GPU")
    ap.add_argument("--breaks", type=float, default=0.005, help="min_bin_size of --breaks 0.5 m (0.005 -> F=201)")
    ap.add_argument("--grid", default="default", choices=["default", "small"])
    ap.add_argument("--tree", default="fixed", choices=["fixed", "rates", "bottleneck"],
                    help="fixed: cfg2 (default); rates: cfg3 (age-scaled leaf branches + one age-scaled internal); "
                         "bottleneck: cfg4 (one ^ branch + age-scaled internal)")
    ap.add_argument("--shard", default="walker", choices=["walker", "locus"],
                    help="N > 1 only.  walker (default, the bench contract): every GPU owns its own --walkers walkers, weak scaling; "
                         "locus: the families are partitioned over the GPUs, every GPU evaluates all --walkers walkers on its rows "
                         "and the per-walker partial sums are combined by ONE NCCL all-reduce (strong scaling; BASELINE config 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-walkers", type=int, default=0, help="walkers per step of the CPU sample (default: host cores)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_tables(args, device=None):
    from mope_b200 import generate
    if args.grid == "default":
        gens, us = generate.DEFAULT_GENS, generate.DEFAULT_MUTS
    else:
        gens, us = [1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 3000], [1e-9, 1e-7, 1e-5, 1e-3]
    nbs = generate.DEFAULT_BOTS if args.tree == "bottleneck" else None
    return generate.generate_tables(N=1000, breaks=(0.5, args.breaks), gens=gens, us=us, nbs=nbs, device=device)


def make_workload(args):
    from mope_b200 import synth
    tree = synth.balanced_tree(args.leaves, rate_leaves=(args.tree == "rates"), rate_internal=(args.tree != "fixed"),
                               bottleneck=(args.tree == "bottleneck"))
    ds = synth.make_dataset(args.loci, args.leaves, tree=tree)
    return ds


def workload_name(args, F):
    kind = {"fixed": "cfg2", "rates": "cfg3-like", "bottleneck": "cfg4-like"}[args.tree]
    return "%s: synthetic L=%d loci x S=%d tissues, balanced %d-leaf tree (B=%d branches, %s), F=%d, W=%d walkers/GPU" % (
        kind, args.loci, args.leaves, args.leaves, 2 * args.leaves - 2, args.tree, F, args.walkers)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (NVML; nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []   # (sm_mhz, max_mhz, power_w, [reason names])
        self.stop = False
        self.thread = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            parts = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(parts) and parts[index].strip().isdigit():
                return int(parts[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        names = []
        for nm, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            if mask & bit:
                names.append(nm)
        self.samples.append((float(sm), float(mx), float(pw), names))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            t = [x.strip() for x in out.split(",")]
            names = [nm for j, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"])
                     if len(t) > 3 + j and t[3 + j].lower().startswith("active")]
            self.samples.append((float(t[0]), float(t[1]), float(t[2]), names))

    def _run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted({r for s in self.samples for r in s[3]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.samples[0][1],
                "power_w_max": max(s[2] for s in self.samples), "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------
_REF_INF = None


def _ref_logl(x):
    return _REF_INF.loglike(np.array(x, dtype=np.float64))


def _write_shim_h5(path, tb):
    """Tables in the pickle format of oracle/ref_shims/h5py.py (what the installed reference reads as HDF5)."""
    import pickle
    datasets = {}
    k = 0
    for gi, g in enumerate(tb.gens):
        for ui, u in enumerate(tb.us):
            datasets["D%d" % k] = (tb.drift[gi, ui], {"N": int(tb.N), "s": 0.0, "u": float(u), "v": float(u), "gen": int(g)})
            k += 1
    blob = {"attrs": {"frequencies": tb.freqs, "breaks": tb.breaks, "N": int(tb.N)}, "datasets": datasets}
    with open(path, "wb") as f:
        pickle.dump(blob, f, protocol=4)


def build_cpu_inference(ds, tb, workdir):
    """The reference's own Inference when baseline/_ref is installed (kind 'reference'), else the oracle port."""
    global _REF_INF
    from oracle import install_ref
    kind = "port"
    if install_ref.activate():
        try:
            import warnings
            warnings.filterwarnings("ignore")
            from mope import inference as refinf  # noqa
            paths = {}
            for key, text in (("data", ds.data_tsv), ("tree", ds.tree), ("ages", ds.ages_tsv)):
                paths[key] = os.path.join(workdir, key + ".txt")
                with open(paths[key], "w") as f:
                    f.write(text)
            h5 = os.path.join(workdir, "drift.h5")
            _write_shim_h5(h5, tb)
            _REF_INF = refinf.Inference(data_files=[paths["data"]], tree_files=[paths["tree"]], age_files=[paths["ages"]],
                                        transitions_file=h5, true_parameters=None, start_from="prior", data_are_freqs=False,
                                        genome_size=16569, bottleneck_file=None, min_freq=0.001, poisson_like_penalty=1.0,
                                        log_unif_drift=True, lower_drift_limit=1e-3, upper_drift_limit=3, min_phred_score=None)
            kind = "reference"
        except Exception as e:  # pragma: no cover
            print("# reference install unusable (%s); using the oracle port" % (e,), file=sys.stderr)
            _REF_INF = None
    if _REF_INF is None:
        from oracle import mope_oracle as orc
        from mope_b200 import model
        drift = orc.TransitionData(tb.drift, tb.gens, tb.us, int(tb.N), tb.freqs, tb.breaks)
        d = model.read_table(ds.data_tsv, True)
        a = model.read_table(ds.ages_tsv, True)
        _REF_INF = orc.Inference([({c: d[c].values for c in d.columns}, ds.tree, {c: a[c].values for c in a.columns})], drift, None,
                                 log_unif_drift=True)
    return kind


def cpu_throughput(X, n_loci, cores, steps, warmup):
    """pool.map(logl, X) with a fork pool, as Inference._get_pool does (mope/inference.py:1197-1239); the workers
    inherit the parent's Inference.  Returns (walker*locus/s, seconds per step)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_ref_logl, list(X), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            vals = pool.map(_ref_logl, list(X), chunksize=1)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    return X.shape[0] * n_loci / dt, dt, np.array(vals)


def subset_dataset(ds, n_loci):
    """First n_loci loci (whole families) of a synthetic dataset, same tree."""
    from mope_b200 import synth
    lines = ds.data_tsv.splitlines()
    n_loci = min(n_loci, len(lines) - 1)
    data = "\n".join(lines[: n_loci + 1]) + "\n"
    fams = {l.split("\t")[-1] for l in lines[1: n_loci + 1]}
    alines = ds.ages_tsv.splitlines()
    ages = "\n".join([alines[0]] + [l for l in alines[1:] if l.split("\t")[0] in fams]) + "\n"
    return synth.SyntheticPedigree(data, ds.tree, ages, ds.leaf_names, n_loci, len(fams))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    try:
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        dev = "cpu"
    tb = make_tables(args, dev)
    ds = make_workload(args)
    from mope_b200 import synth
    nw = args.cpu_walkers or cores
    with tempfile.TemporaryDirectory() as wd:
        # probe: per-core rate on a small subset, to bound the sample so the whole run ends within minutes
        kind = build_cpu_inference(subset_dataset(ds, 150), tb, wd)
        Xp = synth.walkers_in_prior_box(np.asarray(_REF_INF.lower), np.asarray(_REF_INF.upper), max(args.walkers, nw), seed=1)
        t0 = time.perf_counter()
        _ref_logl(Xp[0])
        rate_core = 150 / max(time.perf_counter() - t0, 1e-6)
        budget = 150.0 / max(args.steps + args.warmup, 1)
        n_sample = int(min(args.loci, max(300, rate_core * budget * 0.7)))
        n_sample -= n_sample % 3
        kind = build_cpu_inference(subset_dataset(ds, n_sample), tb, wd)
        X = Xp[:nw]
        n_loci = int(_REF_INF.num_loci[0]) if kind == "reference" else _REF_INF.peds[0].num_loci
        val, dt, _ = cpu_throughput(X, n_loci, min(cores, nw), args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, tb.F), "sample": "%d walkers x first %d of %d loci per step" % (nw, n_loci, args.loci)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": min(cores, nw), "kind": kind,
                         "sample": "%d walkers x first %d of %d loci per step, fork pool of %d processes, BLAS threads 1" % (
                             nw, n_loci, args.loci, min(cores, nw))},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from mope_b200 import Inference, synth
    tb = make_tables(args, "cuda:%d" % local_rank)
    ds = make_workload(args)
    inf = Inference([ds.data_tsv], [ds.tree], [ds.ages_tsv], tb, log_unif_drift=True, device=local_rank, _inputs_are_text=True)
    eng = inf.engine
    W = args.walkers
    # walker-sharded: every rank owns its own W walkers of one global ensemble of W*world
    Xall = synth.walkers_in_prior_box(inf.lower, inf.upper, W * world, seed=1)
    X = np.ascontiguousarray(Xall[rank * W:(rank + 1) * W])
    L = inf.num_loci[0]
    R = L + 2 * inf.peds[0].n_asc
    B = inf.num_branches[0]
    F = tb.F

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): root distributions already in HBM ----
    stat_dev = torch.as_tensor(inf.stationary_distributions(X), device="cuda:%d" % local_rank)
    out_dev = torch.empty(W, dtype=torch.float64, device="cuda:%d" % local_rank)
    for _ in range(args.warmup):
        eng.loglike_device(X, stat_dev, out_dev)
    barrier()
    eng.set_profiling(True)
    eng.kernel_times()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            eng.loglike_device(X, stat_dev, out_dev)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (eng.launch_count() - launches0) // max(args.steps, 1)
    tree_ms, tree_n, interp_ms, interp_n = eng.kernel_times()
    gemm_ops, identity_ops = eng.last_eval_stats()
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = W * world * L / (ms_max * 1e-3)
    ll_dev = out_dev.cpu().numpy()

    # ---- end to end through the public API: host parameter vectors in, host log-likelihoods out ----
    for _ in range(max(1, args.warmup // 2)):
        ll_host = inf.loglike_batch(X, raise_on_error=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll_host = inf.loglike_batch(X, raise_on_error=False)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = W * world * L / float(t.item())
    n_keys = len({(inf.peds[0].varname_of[b], bool(inf.peds[0].br_bottleneck[i])) for i, b in enumerate(inf.peds[0].branch_names)})
    h2d = W * inf.ndim * 8 + W * F * 8 + W * n_keys * (88 + 8) + W * 4
    d2h = W * 8
    if not os.environ.get("MOPE_DEBUG_MODE"):
        assert np.array_equal(np.isnan(ll_host), np.isnan(ll_dev)) and np.allclose(ll_host[np.isfinite(ll_host)], ll_dev[np.isfinite(ll_dev)], rtol=1e-12)

    if rank == 0:
        # ---- roofline of the dominant kernel (the fused pruning kernel), fp64 tensor pipe ----
        # algorithmic flops of the GEMMs the kernel actually executes: 2 F^2 per (row, branch), minus the branches
        # that take the reference's identity short-cut (those are forwarded without a product)
        flops_per_launch = 2.0 * F * F * R * gemm_ops
        tree_avg_ms = tree_ms / max(tree_n, 1)
        achieved = flops_per_launch / (tree_avg_ms * 1e-3) * 1e-12
        peak = eng.fp64_peak_tflops(300.0)
        hbm = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm = json.load(f).get("hbm_gbs")
        except Exception:
            pass
        # compulsory HBM bytes per launch (fused ideal): emissions once per walker pass + matrices + outputs
        n_leaves = len(inf.leaf_names[0])
        alg_bytes = (R * n_leaves * 208 * 8) * 1.0 + W * n_keys * (5 * F * F * 8)
        traffic = None
        try:
            # DRAM bytes of the dominant kernel from the committed ncu capture of this same workload (per launch)
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                tj = json.load(f)
            if args.loci == 10000 and args.leaves == 16 and W == 256 and F == 201:
                traffic = tj["traffic_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": "tree_kernel4<26,25,scalar> (fused pruning pass: TMA ring + mma.sync.m8n8k4.f64, warp-owned rows, messages in tensor memory)", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read+write, profiles/r01_traffic.json)",
                    "peak_source": "fp64 DMMA register-only loop measured live in this run (MEASURED_PEAKS.json has no fp64 entry; "
                                   "its dense-bf16 figure does not bound an fp64 kernel)",
                    "flops_per_launch": flops_per_launch, "branch_gemms_per_launch": int(gemm_ops),
                    "identity_branches_per_launch": int(identity_ops), "avg_launch_ms": tree_avg_ms, "launches_timed": int(tree_n),
                    "share_of_step": tree_avg_ms * (tree_n / max(args.steps, 1)) / ms,
                    "hbm_frac_if_memory_bound": (alg_bytes / (tree_avg_ms * 1e-3) * 1e-9 / hbm) if hbm else None,
                    "interp_avg_ms": interp_ms / max(interp_n, 1)}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            with tempfile.TemporaryDirectory() as wd:
                kind = build_cpu_inference(ds, tb, wd)
                cores = os.cpu_count() or 1
                nw = args.cpu_walkers or cores
                Xc = X[:nw]
                val, dt, ref_ll = cpu_throughput(Xc, L, min(cores, nw), 1, 0)
            fin = np.isfinite(ref_ll) & np.isfinite(ll_host[:nw])
            rel = np.abs(ll_host[:nw][fin] - ref_ll[fin]) / np.abs(ref_ll[fin])
            cpu = {"value": val, "unit": UNIT, "cores": min(cores, nw), "kind": kind,
                   "sample": "%d walkers x all %d loci, one pass (%.1f s), fork pool of %d processes" % (nw, L, dt, min(cores, nw)),
                   "parity_max_rel_err_vs_gpu": float(rel.max()) if rel.size else None,
                   "parity_walkers_compared": int(fin.sum())}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, F), "walkers_total": W * world, "rows_per_walker": R, "branches": B,
                       "drift_grid": "%dx%d" % (tb.drift.shape[0], tb.drift.shape[1]), "parallelism": "walker-sharded x%d" % world,
                       "l2_policy": "inputs larger than L2 (emissions %.0f MB + per-walker matrices %.0f MB per step)" % (
                           R * n_leaves * 208 * 8 / 1e6, W * n_keys * 208 * 209 * 8 / 1e6),
                       "finite_walkers": int(np.isfinite(ll_host).sum())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": float(t.item()) * 1e3},
            "gpu_launches": int(launches * args.steps),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    inf.close()


def run_ours_locus(args):
    """Locus-sharded evaluation over the ranks (mope_b200.dist): not the default bench line, a measured demonstration of
    the path's only collective.  value = walkers x ALL loci / step time (strong scaling in the rows)."""
    import torch
    import torch.distributed as dist
    from mope_b200 import Inference, synth
    from mope_b200 import dist as mdist

    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tb = make_tables(args, "cuda:%d" % local_rank)
    ds = make_workload(args)
    inf = mdist.build_locus_sharded([ds.data_tsv], [ds.tree], [ds.ages_tsv], tb, rank, world, log_unif_drift=True, device=local_rank)
    W = args.walkers
    X = synth.walkers_in_prior_box(inf.lower, inf.upper, W, seed=1)     # the same ensemble on every rank
    Lt = torch.tensor([inf.num_loci[0]], dtype=torch.int64, device="cuda")
    dist.all_reduce(Lt)
    L_total = int(Lt.item())
    dev = "cuda:%d" % local_rank
    stat_dev = torch.as_tensor(inf.stationary_distributions(X), device=dev)
    out_dev = torch.empty(W, dtype=torch.float64, device=dev)

    def step():
        inf.engine.loglike_device(X, stat_dev, out_dev)
        dist.all_reduce(out_dev, op=dist.ReduceOp.SUM)     # the path's only collective: W float64 partial sums

    for _ in range(args.warmup):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ll_dev = out_dev.cpu().numpy()
    # end to end through the public API (host vectors in, host log-likelihoods out, all-reduce inside)
    mdist.loglike_locus_sharded(inf, X)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll_host, st = mdist.loglike_locus_sharded(inf, X)
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    parity = None
    if rank == 0:
        # the sharded sum against the whole dataset on one GPU
        full = Inference([ds.data_tsv], [ds.tree], [ds.ages_tsv], tb, log_unif_drift=True, device=local_rank, _inputs_are_text=True)
        ref = full.loglike_batch(X[:32], raise_on_error=False)
        full.close()
        fin = np.isfinite(ref) & np.isfinite(ll_host[:32])
        parity = float((np.abs(ll_host[:32][fin] - ref[fin]) / np.abs(ref[fin])).max())
        assert parity < 1e-11, parity
        line = {"metric": METRIC, "value": W * L_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, tb.F), "walkers_total": W, "loci_total": L_total,
                           "parallelism": "locus-sharded x%d (families partitioned; one NCCL all-reduce of %d doubles per evaluation)" % (world, W),
                           "finite_walkers": int(np.isfinite(ll_dev).sum())},
                "e2e": {"value": W * L_total / float(te.item()), "unit": UNIT, "ms_per_step": float(te.item()) * 1e3},
                "parity_max_rel_err_vs_single_gpu": parity, "clocks": clocks.summary()}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    inf.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.shard == "locus" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_ours_locus(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

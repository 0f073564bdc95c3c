#!/usr/bin/env python
"""Benchmark of the hot path: walker x locus log-likelihoods per second (BASELINE.json metric).

Headline workload (configs[1] of BASELINE.json, "cfg2" in SURVEY.md section 8d): synthetic heteroplasmy data, L = 10 000
loci x S = 16 tissues, balanced 16-leaf tree (30 fixed-length branches, 8 parameter names), F = 201 frequency
bins (WF N = 1000, --breaks 0.5 0.005), the reference's default 94 x 44 drift grid generated locally, W = 256
walkers drawn uniformly in the prior box.  A step = one evaluation of Inference.loglike for all W walkers.
Multi-GPU: walkers are sharded, W = 256 per GPU (weak scaling), no data-path collective.

The same JSON line carries a `configs` array with the other BASELINE.json configurations at their named sizes, each
with its own ms/step, e2e, roofline (executed and useful flops) and -- on one GPU -- parity against the reference's own
code on a stated sample:
    cfg3  100 000 loci, 16 age-scaled leaf branches + 1 age-scaled internal branch, 32 walkers per GPU (walker-sharded)
    cfg4  1 000 000 loci, one bottleneck (^) branch + ascertainment, 64 walkers; with N > 1 the FAMILIES are partitioned
          over the ranks and the per-walker partial sums are combined by ONE NCCL all-reduce (strong scaling)
    cfg5  F = 1001 (five N passes per branch, wide-grid kernel), 1000 loci, 512 walkers per GPU

    python bench.py [--gpus N --steps K --warmup W]          our arm (one JSON line on rank 0)
    python bench.py --configs cfg2                           headline only (quick)
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

SMALL_GENS = [1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 3000]
SMALL_US = [1e-9, 1e-7, 1e-5, 1e-3]
# reduced grids of the extra configurations (grid extent does not enter the cost of an evaluation)
CFG5_GENS = [1, 2, 5, 10, 20, 50, 100, 200, 400, 700, 1000, 1500, 2000, 3000, 5000, 10000]
CFG5_US = [1e-12, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4]
CFG4_NBS = [2, 5, 10, 20, 50, 100, 200, 350, 500]
CFG4_BOT_US = [1e-12, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-3]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # workload overrides of the headline (defaults are the BASELINE config)
    ap.add_argument("--loci", type=int, default=10000)
    ap.add_argument("--leaves", type=int, default=16)
    ap.add_argument("--walkers", type=int, default=256, help="walkers per GPU")
    ap.add_argument("--breaks", type=float, default=0.005, help="min_bin_size of --breaks 0.5 m (0.005 -> F=201)")
    ap.add_argument("--grid", default="default", choices=["default", "small"])
    ap.add_argument("--tree", default="fixed", choices=["fixed", "rates", "bottleneck"],
                    help="fixed: cfg2 (default); rates: cfg3 (age-scaled leaf branches + one age-scaled internal); "
                         "bottleneck: cfg4 (one ^ branch + age-scaled internal)")
    ap.add_argument("--shard", default="walker", choices=["walker", "locus"],
                    help="N > 1 only, headline workload.  walker (default, the bench contract): every GPU owns its own --walkers "
                         "walkers, weak scaling; locus: the families are partitioned over the GPUs (strong scaling)")
    ap.add_argument("--configs", default="all",
                    help="all (default): the headline plus cfg3, cfg4, cfg5 in the `configs` array; or a comma list, e.g. cfg2 "
                         "(headline only) or cfg2,cfg4.  The extra configurations run only with the default headline workload.")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every CPU leg (cpu_baseline and the parity samples)")
    ap.add_argument("--cpu-walkers", type=int, default=0, help="walkers per step of the CPU sample (default: host cores)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_tables(args, device=None, with_bottleneck=None):
    from mope_b200 import generate
    if args.grid == "default":
        gens, us = generate.DEFAULT_GENS, generate.DEFAULT_MUTS
    else:
        gens, us = SMALL_GENS, SMALL_US
    want_bot = (args.tree == "bottleneck") if with_bottleneck is None else with_bottleneck
    nbs = generate.DEFAULT_BOTS if want_bot else None
    return generate.generate_tables(N=1000, breaks=(0.5, args.breaks), gens=gens, us=us, nbs=nbs, device=device)


def make_workload(args):
    from mope_b200 import synth
    tree = synth.balanced_tree(args.leaves, rate_leaves=(args.tree == "rates"), rate_internal=(args.tree != "fixed"),
                               bottleneck=(args.tree == "bottleneck"))
    return synth.make_dataset(args.loci, args.leaves, tree=tree)


def workload_name(args, F):
    kind = {"fixed": "cfg2", "rates": "cfg3-like", "bottleneck": "cfg4-like"}[args.tree]
    return "%s: synthetic L=%d loci x S=%d tissues, balanced %d-leaf tree (B=%d branches, %s), F=%d, W=%d walkers/GPU" % (
        kind, args.loci, args.leaves, args.leaves, 2 * args.leaves - 2, args.tree, F, args.walkers)


def is_default_headline(args):
    return (args.loci, args.leaves, args.walkers, args.breaks, args.grid, args.tree, args.shard) == \
        (10000, 16, 256, 0.005, "default", "fixed", "walker")


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
# CPU reference arm (test infrastructure: the only code here that touches oracle/ and baseline/_ref)
# ------------------------------------------------------------------------------------------------
_REF_INF = None


def _ref_logl(x):
    return _REF_INF.loglike(np.array(x, dtype=np.float64))


def _write_shim_h5(path, tb, bottleneck=False):
    """Tables in the pickle format of oracle/ref_shims/h5py.py (what the installed reference reads as HDF5): one dataset
    per grid point with the attributes the reference's loaders read (transition_data_2d.py:78-111,
    transition_data_bottleneck.py:74-96)."""
    import pickle
    datasets = {}
    k = 0
    if bottleneck:
        for bi, nb in enumerate(tb.nbs):
            for ui, u in enumerate(tb.bot_us):
                datasets["D%d" % k] = (tb.bot[bi, ui], {"N": int(tb.N), "Nb": int(nb), "u": float(u)})
                k += 1
    else:
        for gi, g in enumerate(tb.gens):
            for ui, u in enumerate(tb.us):
                datasets["D%d" % k] = (tb.drift[gi, ui], {"N": int(tb.N), "s": 0.0, "u": float(u), "v": float(u), "gen": int(g)})
                k += 1
    blob = {"attrs": {"frequencies": tb.freqs, "breaks": tb.breaks, "N": int(tb.N)}, "datasets": datasets}
    with open(path, "wb") as f:
        pickle.dump(blob, f, protocol=4)


class RefTables:
    """The shim-HDF5 files of one table set, written once per bench run and shared by the CPU legs that use it."""

    def __init__(self, tb, workdir, tag):
        self.tb = tb
        self.drift = os.path.join(workdir, "drift_%s.h5" % tag)
        self.bot = os.path.join(workdir, "bot_%s.h5" % tag) if tb.bot is not None else None
        self._written = False

    def ensure(self):
        if not self._written:
            _write_shim_h5(self.drift, self.tb)
            if self.bot:
                _write_shim_h5(self.bot, self.tb, bottleneck=True)
            self._written = True


def build_cpu_inference(ds, ref_tables, workdir):
    """The reference's own Inference when baseline/_ref is installed (kind 'reference'), else the oracle port."""
    global _REF_INF
    from oracle import install_ref
    tb = ref_tables.tb
    kind = "port"
    _REF_INF = None
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
            ref_tables.ensure()
            _REF_INF = refinf.Inference(data_files=[paths["data"]], tree_files=[paths["tree"]], age_files=[paths["ages"]],
                                        transitions_file=ref_tables.drift, true_parameters=None, start_from="prior", data_are_freqs=False,
                                        genome_size=16569, bottleneck_file=ref_tables.bot, min_freq=0.001, poisson_like_penalty=1.0,
                                        log_unif_drift=True, lower_drift_limit=1e-3, upper_drift_limit=3, min_phred_score=None)
            kind = "reference"
        except Exception as e:  # pragma: no cover
            print("# reference install unusable (%r); using the oracle port" % (e,), file=sys.stderr)
            _REF_INF = None
    if _REF_INF is None:
        from oracle import mope_oracle as orc
        drift = orc.TransitionData(tb.drift, tb.gens, tb.us, int(tb.N), tb.freqs, tb.breaks)
        bot = orc.TransitionDataBottleneck(tb.bot, tb.nbs, tb.bot_us, int(tb.N)) if tb.bot is not None else None
        d, a = ds.data, ds.ages
        _REF_INF = orc.Inference([({c: d[c].values for c in d.columns}, ds.tree, {c: a[c].values for c in a.columns})], drift, bot,
                                 log_unif_drift=True)
    return kind


def cpu_throughput(X, n_loci, cores, steps, warmup):
    """pool.map(logl, X) with a fork pool, as Inference._get_pool does (mope/inference.py:1197-1239); the workers
    inherit the parent's Inference.  Returns (walker*locus/s, seconds per step, values)."""
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


def reference_parity(ds, tb, ref_tables, X, n_loci_sample, workdir, device, label, max_walkers=16):
    """Parity of the GPU path against the reference's own code on a sample: the first `n_loci_sample` loci (whole
    families) of the configuration's dataset, the configuration's tables, the first walkers of its ensemble.  Both
    implementations evaluate the same sample; returns the dict that goes into the bench line."""
    from mope_b200 import Inference
    cores = os.cpu_count() or 1
    nw = int(min(X.shape[0], max(8, min(cores, max_walkers))))
    sub = ds.subset(n_loci_sample)
    t0 = time.perf_counter()
    kind = build_cpu_inference(sub, ref_tables, workdir)
    Xs = np.ascontiguousarray(X[:nw])
    n_loci = int(_REF_INF.num_loci[0]) if kind == "reference" else _REF_INF.peds[0].num_loci
    val, dt, ref_ll = cpu_throughput(Xs, n_loci, min(cores, nw), 1, 0)
    g = Inference([sub.data], [sub.tree], [sub.ages], tb, log_unif_drift=True, device=device, _inputs_are_text=True)
    try:
        gpu_ll = g.loglike_batch(Xs, raise_on_error=False)
    finally:
        g.close()
    fin = np.isfinite(ref_ll) & np.isfinite(gpu_ll)
    rel = np.abs(gpu_ll[fin] - ref_ll[fin]) / np.abs(ref_ll[fin])
    return {"max_rel_err": float(rel.max()) if rel.size else None, "walkers_compared": int(fin.sum()), "walkers": nw,
            "against": "baseline/_ref (the reference's own Cython/NumPy code)" if kind == "reference" else "oracle port",
            "sample": "%s: first %d loci (%d families) of the dataset, same tables, first %d walkers of the ensemble" % (
                label, n_loci, sub.n_fam, nw),
            "nonfinite_mismatch": bool(np.any(np.isfinite(ref_ll) != np.isfinite(gpu_ll))),
            "cpu_value": val, "cpu_seconds": dt, "cpu_cores": min(cores, nw), "wall_s": time.perf_counter() - t0}


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
        rt = RefTables(tb, wd, "main")
        # probe: per-core rate on a small subset, to bound the sample so the whole run ends within minutes
        kind = build_cpu_inference(ds.subset(150), rt, wd)
        Xp = synth.walkers_in_prior_box(np.asarray(_REF_INF.lower), np.asarray(_REF_INF.upper), max(args.walkers, nw), seed=1)
        t0 = time.perf_counter()
        _ref_logl(Xp[0])
        rate_core = 150 / max(time.perf_counter() - t0, 1e-6)
        budget = 150.0 / max(args.steps + args.warmup, 1)
        n_sample = int(min(args.loci, max(300, rate_core * budget * 0.7)))
        n_sample -= n_sample % 3
        kind = build_cpu_inference(ds.subset(n_sample), rt, wd)
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
class Dist:
    """torch.distributed plumbing of one bench process (one process per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.device = "cuda:%d" % self.local_rank
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, v):
        """[value of every rank] on every rank."""
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world == 1:
            return [float(v)]
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t)
        return [float(p.item()) for p in parts]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure(D, inf, X, steps, warmup, locus_sharded=False, sample_clocks=True):
    """Time `steps` evaluations of all walkers in X on this rank's Inference.
    value leg: everything resident (tables, emissions, the walkers' root prior already in HBM), CUDA events on the launching
    stream, bracketed by barrier + synchronize, max over ranks.  With locus_sharded the step ends with the all-reduce of
    the per-walker partial sums, enqueued on the same stream.
    e2e leg: the public API with host NumPy vectors in and host log-likelihoods out (root prior on the host, H2D of
    parameters / root prior / plans, D2H of the result inside the timed region)."""
    torch, dist = D.torch, D.dist
    from mope_b200 import dist as mdist
    eng = inf.engine
    W = X.shape[0]
    stat_dev = torch.as_tensor(inf.stationary_distributions(X), device=D.device)
    out_dev = torch.empty(W, dtype=torch.float64, device=D.device)

    def step():
        eng.loglike_device(X, stat_dev, out_dev)
        if locus_sharded and D.world > 1:
            dist.all_reduce(out_dev, op=dist.ReduceOp.SUM)     # the path's only collective: W float64 partial sums

    for _ in range(warmup):
        step()
    D.barrier()
    eng.set_profiling(True)
    eng.kernel_times()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(D.local_rank) if sample_clocks else None
    if clocks:
        clocks.__enter__()
    D.barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    D.barrier()
    if clocks:
        clocks.__exit__()
    ms_local = e0.elapsed_time(e1) / steps
    launches = (eng.launch_count() - launches0) // max(steps, 1)
    tree_ms, tree_n, interp_ms, interp_n = eng.kernel_times()
    gemm_ops, identity_ops = eng.last_eval_stats()
    unit_rows = eng.last_eval_unit_rows()
    warp_segs, warp_segs_active = eng.last_eval_segment_stats()
    eng.set_profiling(False)
    ms = D.max_over_ranks(ms_local)
    ll_dev = out_dev.cpu().numpy()

    def e2e_step():
        if locus_sharded and D.world > 1:
            return mdist.loglike_locus_sharded(inf, X)[0]
        return inf.loglike_batch(X, raise_on_error=False)

    for _ in range(max(1, warmup // 2)):
        ll_host = e2e_step()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ll_host = e2e_step()
    torch.cuda.synchronize()
    e2e_local = (time.perf_counter() - t0) / steps
    e2e_s = D.max_over_ranks(e2e_local)
    # the two legs evaluate the same walkers; they differ only in where the root prior comes from (value leg: SciPy for every
    # walker; e2e leg: the device kernel for ab >= 1e-6), i.e. by ~1e-13 relative.  Anything beyond 1e-10 is an error: it is
    # reported on stderr with the walker, both legs are evaluated once more, and a second disagreement aborts the run.
    def legs_diff(a_host, a_dev):
        both = np.isfinite(a_host) & np.isfinite(a_dev)
        rel = np.where(both, np.abs(a_host - a_dev) / np.maximum(np.abs(a_dev), 1e-300), 0.0)
        return float(rel.max()) if rel.size else 0.0, int(np.argmax(rel)) if rel.size else 0, np.array_equal(np.isnan(a_host), np.isnan(a_dev))

    legs_rel, worst, nan_ok = legs_diff(ll_host, ll_dev)
    legs_retried = 0
    if not os.environ.get("MOPE_DEBUG_MODE") and (not nan_ok or legs_rel > 1e-10):
        sys.stderr.write("[bench] rank %d: device-resident and e2e results differ: max rel %.3e at walker %d (%.17g vs %.17g, log10ab %.4f), "
                         "NaN patterns %s; evaluating both legs again\n" % (D.rank, legs_rel, worst, ll_dev[worst], ll_host[worst], X[worst, -2],
                                                                            "equal" if nan_ok else "DIFFER"))
        legs_retried = 1
        step()
        torch.cuda.synchronize()
        ll_dev2 = out_dev.cpu().numpy()
        ll_host2 = e2e_step()
        sys.stderr.write("[bench] rank %d: second evaluation: value leg changed by %.3e, e2e leg changed by %.3e\n" % (
            D.rank, legs_diff(ll_dev2, ll_dev)[0], legs_diff(ll_host2, ll_host)[0]))
        ll_dev, ll_host = ll_dev2, ll_host2
        legs_rel, worst, nan_ok = legs_diff(ll_host, ll_dev)
        if not nan_ok or legs_rel > 1e-10:
            raise AssertionError("device-resident and e2e results differ (twice)")
    steps_per_eval = tree_n / max(steps, 1)
    return {"ms": ms, "ms_local": ms_local, "e2e_s": e2e_s, "e2e_local_s": e2e_local, "launches_per_step": int(launches),
            "tree_ms_per_step": tree_ms / max(steps, 1), "tree_launches_per_step": steps_per_eval,
            "interp_ms_per_step": interp_ms / max(steps, 1), "interp_launches_per_step": interp_n / max(steps, 1),
            "gemm_ops": int(gemm_ops), "identity_ops": int(identity_ops), "unit_rows": int(unit_rows),
            "warp_segs": int(warp_segs), "warp_segs_active": int(warp_segs_active), "legs_max_rel_diff": legs_rel, "legs_retried": legs_retried,
            "ll_host": ll_host, "clocks": clocks.summary() if clocks else None}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs"), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md (of fallback)"


def committed_traffic(key):
    """DRAM bytes per launch of a kernel from the committed ncu capture of the same workload (profiles/r02_traffic.json);
    ncu cannot run inside the timed process, so this figure is read, not measured, in this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def tree_roofline(m, F, R_exec, R_ref, B, W, peak, kernel):
    """Roofline of the pruning kernel of one configuration (fp64 tensor pipe).
    executed: 2 F^2 x (real rows x GEMM units the kernel ran, counted by the kernel itself);
    useful:   2 F^2 B R W with R = L + 2 n_fam, the branch updates the reference performs for the same evaluation."""
    t = m["tree_ms_per_step"] * 1e-3
    executed = 2.0 * F * F * m["unit_rows"]
    useful = 2.0 * F * F * B * R_ref * W
    return {"bound": "tensor", "kernel": kernel, "achieved": executed / t * 1e-12, "peak": peak, "unit": "TFLOP/s",
            "frac": executed / t * 1e-12 / peak, "executed_flops_per_step": executed,
            "useful_flops_per_step": useful, "useful_tflops": useful / t * 1e-12, "useful_frac": useful / t * 1e-12 / peak,
            "tree_kernel_ms_per_step": m["tree_ms_per_step"], "tree_launches_per_step": m["tree_launches_per_step"],
            "share_of_step": m["tree_ms_per_step"] / m["ms_local"], "interp_ms_per_step": m["interp_ms_per_step"],
            "branch_gemms_per_step": m["gemm_ops"], "identity_branches_per_step": m["identity_ops"],
            "rows_per_walker_executed": R_exec, "rows_per_walker_reference": R_ref}


def interp_roofline(m, F, Fp, n_fixed_blocks, n_rate_blocks):
    """HBM roofline of interp_kernel: per matrix block 4 (shared-matrix branches) or 2 (age-scaled: mutation axis only)
    F x F source matrices read, one Fp x (Fp + 3) block written."""
    hbm, src = hbm_peak()
    if m["interp_ms_per_step"] <= 0:
        return None
    bytes_ = n_fixed_blocks * (4 * F * F * 8 + Fp * (Fp + 3) * 8) + n_rate_blocks * (2 * F * F * 8 + Fp * (Fp + 3) * 8)
    gbs = bytes_ / (m["interp_ms_per_step"] * 1e-3) * 1e-9
    return {"bound": "hbm", "kernel": "interp_kernel", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
            "bytes_per_step": bytes_, "ms_per_step": m["interp_ms_per_step"], "peak_source": src,
            "traffic": committed_traffic("interp_kernel_bytes_per_step_cfg2")}


def run_extra_config(D, name, base_tb, workdir, peak, do_cpu, steps_cap):
    """One of cfg3 / cfg4 / cfg5 at its BASELINE.json size -> dict for the `configs` array (rank 0), None elsewhere."""
    import torch
    from mope_b200 import Inference, synth, generate
    from mope_b200 import dist as mdist
    t_begin = time.perf_counter()
    locus = False
    if name == "cfg3":
        loci, W, tree = 100000, 32, synth.balanced_tree(16, rate_leaves=True, rate_internal=True)
        tb, tag, n_sample = base_tb, "main", 600
        desc = "cfg3: L=100000 loci x 16 tissues, 16 age-scaled leaf branches + 1 age-scaled internal branch (17 of B=30), F=201, W=32 walkers/GPU, walker-sharded"
    elif name == "cfg4":
        loci, W, tree = 1000000, 64, synth.balanced_tree(16, rate_internal=True, bottleneck=True)
        locus = D.world > 1
        from mope_b200.tables import TransitionTables
        b = generate.generate_tables(N=1000, breaks=(0.5, 0.005), gens=[1, 2], us=[1e-9, 1e-8], nbs=CFG4_NBS, bot_us=CFG4_BOT_US,
                                     device=D.device)
        tb = TransitionTables(drift=base_tb.drift, gens=base_tb.gens, us=base_tb.us, N=base_tb.N, freqs=base_tb.freqs,
                              breaks=base_tb.breaks, bot=b.bot, nbs=b.nbs, bot_us=b.bot_us)
        tag, n_sample = "cfg4", 600
        desc = ("cfg4: L=1000000 loci x 16 tissues, one bottleneck (^) branch + one age-scaled internal branch, ascertainment on, F=201, "
                "W=64 walkers, " + ("locus-sharded over %d GPUs (families partitioned, ONE NCCL all-reduce of 64 doubles per evaluation, "
                                    "enqueued on the compute stream)" % D.world if locus else "one GPU (emissions 35 GB resident)"))
    elif name == "cfg5":
        loci, W, tree = 1000, 512, synth.balanced_tree(16)
        tb = generate.generate_tables(N=1000, breaks=(0.5, 1e-9), gens=CFG5_GENS, us=CFG5_US, device=D.device)
        tag, n_sample = "cfg5", 6      # the reference needs ~1 s per (row, walker) at F = 1001: keep the sample to seconds
        desc = "cfg5: F=1001 (WF N=1000, every count its own bin), L=1000 loci x 16 tissues, B=30 fixed-length branches, W=512 walkers/GPU, walker-sharded, 16x8 drift grid"
    else:
        raise ValueError(name)
    ds = synth.make_dataset(loci, 16, tree=tree)
    t_data = time.perf_counter() - t_begin
    if locus:
        inf = mdist.build_locus_sharded([ds.data], [ds.tree], [ds.ages], tb, D.rank, D.world, log_unif_drift=True, device=D.local_rank)
        X = synth.walkers_in_prior_box(inf.lower, inf.upper, W, seed=1)       # the same ensemble on every rank
        L_total = int(round(D.sum_over_ranks(inf.num_loci[0])))
        W_total = W
    else:
        inf = Inference([ds.data], [ds.tree], [ds.ages], tb, log_unif_drift=True, device=D.local_rank, _inputs_are_text=True)
        if name == "cfg4":
            X = synth.walkers_in_prior_box(inf.lower, inf.upper, W, seed=1)
            W_total = W
        else:
            Xall = synth.walkers_in_prior_box(inf.lower, inf.upper, W * D.world, seed=1)
            X = np.ascontiguousarray(Xall[D.rank * W:(D.rank + 1) * W])
            W_total = W * D.world
        L_total = inf.num_loci[0]
    t_setup = time.perf_counter() - t_begin
    # steps: about 3 s of timed work, at least 2
    probe = measure(D, inf, X, 1, 1, locus_sharded=locus, sample_clocks=False)
    steps = int(max(2, min(steps_cap, round(3000.0 / max(probe["ms"], 1e-3)))))
    m = measure(D, inf, X, steps, 1, locus_sharded=locus)
    ped = inf.peds[0]
    F, Fp = tb.F, (tb.F + 15) // 16 * 16
    R_exec = ped.num_loci + 2 * ped.n_asc
    R_ref = ped.num_loci + 2 * ped.n_fam
    B = inf.num_branches[0]
    kernel = ("tree_kernel<13> (F > 208: five N passes per branch, messages through the L2 scratch ring)" if F > 208 else
              "tree_kernel4<26,25,scalar" + (",ascfast>" if R_exec - ped.num_loci >= 64 else ">") +
              (" + canonical launch (family-independent ascertainment subtrees once per walker)" if m["tree_launches_per_step"] > 1.5 else ""))
    roof = tree_roofline(m, F, R_exec, R_ref, B, X.shape[0], peak, kernel)
    n_rate = int((ped.br_mult >= 0).sum())
    if n_rate > 0 and m["warp_segs"] > 0:
        # age-scaled branches: every branch multiplies against each grid generation the tile's rows blend (2 is the minimum: a
        # row blends two generations); a warp none of whose 8 rows touches a segment idles through it
        warps = (R_exec + 7) // 8
        roof["age_scaled"] = {"branches": n_rate,
                              "segments_per_warp_and_branch": m["warp_segs"] / float(X.shape[0] * n_rate * warps),
                              "active_segments_per_warp_and_branch": m["warp_segs_active"] / float(X.shape[0] * n_rate * warps),
                              "idle_warp_segment_frac": 1.0 - m["warp_segs_active"] / float(m["warp_segs"]),
                              "note": "counted by the kernel (mope_b200_last_eval_segment_stats); pure pseudo-row tiles take the "
                                      "messages of age-scaled leaf branches from the stored column sums and run no segment"}
    per_rank = {"ms_per_step": D.gather(m["ms_local"]), "e2e_ms_per_step": D.gather(m["e2e_local_s"] * 1e3),
                "tree_kernel_ms_per_step": D.gather(m["tree_ms_per_step"]), "loci": D.gather(inf.num_loci[0]),
                "rows_per_walker": D.gather(R_exec)}
    value = W_total * L_total / (m["ms"] * 1e-3)
    e2e_value = W_total * L_total / m["e2e_s"]
    ll_host = m["ll_host"]

    pb = None
    if name == "cfg4" and D.rank == 0:
        pb = poisson_binomial_leg(inf, tb, do_cpu and D.world == 1)
    parity = None
    sharded_parity = None
    if locus:
        # the sharded sum against the whole dataset on ONE GPU, 8 walkers (rank 0 builds the unsharded model; the others wait)
        if D.rank == 0:
            full = Inference([ds.data], [ds.tree], [ds.ages], tb, log_unif_drift=True, device=D.local_rank, _inputs_are_text=True)
            ref = full.loglike_batch(X[:8], raise_on_error=False)
            full.close()
            fin = np.isfinite(ref) & np.isfinite(ll_host[:8])
            sharded_parity = {"max_rel_err": float((np.abs(ll_host[:8][fin] - ref[fin]) / np.abs(ref[fin])).max()),
                              "walkers_compared": int(fin.sum()), "against": "the same 1 000 000 loci evaluated unsharded on one GPU"}
        D.barrier()
    inf.close()
    del inf
    torch.cuda.empty_cache()
    if do_cpu and D.world == 1 and D.rank == 0:
        parity = reference_parity(ds, tb, RefTables(tb, workdir, tag) if tag != "main" else run_extra_config.main_ref_tables,
                                  X, n_sample, workdir, D.local_rank, name, max_walkers=8 if name == "cfg5" else 16)
    if D.rank != 0:
        return None
    out = {"name": name, "workload": desc, "n_gpus": D.world, "scaling": "strong" if name == "cfg4" else "weak",
           "walkers_total": W_total, "loci_total": int(L_total), "steps": steps, "warmup": 1,
           "ms_per_step": m["ms"], "value": value, "unit": UNIT,
           "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": m["e2e_s"] * 1e3, "max_rel_diff_vs_value_leg": m["legs_max_rel_diff"],
                   "legs_retried": m["legs_retried"]},
           "gpu_launches_per_step": m["launches_per_step"], "finite_walkers": int(np.isfinite(ll_host).sum()),
           "roofline": roof, "per_rank": per_rank, "parity_vs_reference": parity, "clocks": m["clocks"],
           "setup_s": {"data": t_data, "model_and_tables": t_setup - t_data, "total_wall": time.perf_counter() - t_begin}}
    if sharded_parity is not None:
        out["parity_sharded_vs_single_gpu"] = sharded_parity
    if pb is not None:
        out["poisson_binomial"] = pb
    return out


def poisson_binomial_leg(inf, tb, do_cpu):
    """cfg4's emission variant: Poisson-binomial non-ascertainment vectors P(alt reads < ceil(min_freq * nreads) | f)
    for read sets with nreads ~ Poisson(1000) and per-read phred from {20, 30, 40} w.p. {0.1, 0.6, 0.3}, on the F = 201 grid
    (mope_b200_non_asc_probs; _poisson_binom.get_non_asc_probs is referenced by nothing in the reference, so this is a
    function-level leg, not part of the log-likelihood).  Host read sets in, host vectors out."""
    rng = np.random.RandomState(41)
    n_sets = 2048
    sets = [rng.choice([20.0, 30.0, 40.0], size=max(1, int(n)), p=[0.1, 0.6, 0.3]) for n in rng.poisson(1000, n_sets)]
    n_reads = int(sum(len(q) for q in sets))
    eng = inf.engine
    out = {"read_sets": n_sets, "reads": n_reads, "F": int(tb.F), "runs": []}
    for mf in (0.001, 0.05):
        eng.non_asc_probs(sets[:64], tb.freqs, mf)                       # warm-up
        t0 = time.perf_counter()
        res = eng.non_asc_probs(sets, tb.freqs, mf)
        dt = time.perf_counter() - t0
        K = int(np.ceil(mf * 1000))
        run = {"min_freq": mf, "K_typical": K, "ms": dt * 1e3, "read_sets_per_s": n_sets / dt, "dp_updates_per_s": n_reads * K * tb.F / dt}
        if do_cpu:
            from oracle import mope_oracle as orc
            ref = np.array([orc.get_non_asc_probs(sets[i], tb.freqs, mf) for i in range(2)])
            run["parity_max_rel_err_vs_oracle"] = float((np.abs(res[:2] - ref) / np.maximum(np.abs(ref), 1e-300)).max())
            run["parity_sets_compared"] = 2
        out["runs"].append(run)
    return out


def run_ours(args):
    import torch
    D = Dist()
    rank, local_rank, world = D.rank, D.local_rank, D.world
    from mope_b200 import Inference, synth
    from mope_b200 import dist as mdist

    tb = make_tables(args, D.device)
    ds = make_workload(args)
    W = args.walkers
    locus = (args.shard == "locus" and world > 1)
    if locus:
        inf = mdist.build_locus_sharded([ds.data], [ds.tree], [ds.ages], tb, rank, world, log_unif_drift=True, device=local_rank)
        X = synth.walkers_in_prior_box(inf.lower, inf.upper, W, seed=1)
        W_total, L = W, int(round(D.sum_over_ranks(inf.num_loci[0])))
    else:
        inf = Inference([ds.data], [ds.tree], [ds.ages], tb, log_unif_drift=True, device=local_rank, _inputs_are_text=True)
        # walker-sharded: every rank owns its own W walkers of one global ensemble of W*world
        Xall = synth.walkers_in_prior_box(inf.lower, inf.upper, W * world, seed=1)
        X = np.ascontiguousarray(Xall[rank * W:(rank + 1) * W])
        W_total, L = W * world, inf.num_loci[0]
    eng = inf.engine
    ped = inf.peds[0]
    R = L + 2 * ped.n_asc if not locus else None
    R_local = inf.num_loci[0] + 2 * ped.n_asc
    B = inf.num_branches[0]
    F, Fp = tb.F, (tb.F + 15) // 16 * 16

    m = measure(D, inf, X, args.steps, args.warmup, locus_sharded=locus)
    value = W_total * L / (m["ms"] * 1e-3)
    e2e_value = W_total * L / m["e2e_s"]
    ll_host = m["ll_host"]
    n_keys = len({(ped.varname_of[b], bool(ped.br_bottleneck[i])) for i, b in enumerate(ped.branch_names)})
    h2d = W * inf.ndim * 8 + W * F * 8 + W * n_keys * (88 + 8) + W * 4
    d2h = W * 8
    per_rank = {"ms_per_step": D.gather(m["ms_local"]), "e2e_ms_per_step": D.gather(m["e2e_local_s"] * 1e3),
                "tree_kernel_ms_per_step": D.gather(m["tree_ms_per_step"]), "interp_ms_per_step": D.gather(m["interp_ms_per_step"])}

    line = None
    workdir = tempfile.mkdtemp(prefix="mope_bench_")
    main_ref = RefTables(tb, workdir, "main")
    run_extra_config.main_ref_tables = main_ref
    peak = None
    if rank == 0:
        # ---- roofline of the dominant kernel (the fused pruning kernel), fp64 tensor pipe ----
        # executed flops: 2 F^2 per (real row, GEMM unit) the kernel ran, counted by the kernel; branches that take the
        # reference's identity short-cut are forwarded without a product and do not count
        peak = eng.fp64_peak_tflops(300.0)
        kernel = ("tree_kernel4<26,25,scalar> (fused pruning pass: TMA ring + mma.sync.m8n8k4.f64, warp-owned rows, messages in tensor memory)"
                  if F == 201 else "tree_kernel4 / tree_kernel instance for F=%d" % F)
        roofline = tree_roofline(m, F, R_local, inf.num_loci[0] + 2 * ped.n_fam, B, W, peak, kernel)
        hbm, _src = hbm_peak()
        n_leaves = len(inf.leaf_names[0])
        alg_bytes = (R_local * n_leaves * Fp * 8) * 1.0 + W * n_keys * (5 * F * F * 8)
        roofline.update({
            "peak_source": "fp64 DMMA register-only loop measured live in this run (MEASURED_PEAKS.json has no fp64 entry; "
                           "its dense-bf16 figure does not bound an fp64 kernel)",
            "flops_per_launch": roofline["executed_flops_per_step"] / max(m["tree_launches_per_step"], 1),
            "avg_launch_ms": m["tree_ms_per_step"] / max(m["tree_launches_per_step"], 1),
            "launches_timed": int(round(m["tree_launches_per_step"] * args.steps)),
            "traffic": committed_traffic("tree_kernel4_bytes_per_launch_cfg2") if is_default_headline(args) else None,
            "traffic_unit": "bytes/launch (ncu dram__bytes_read+write of the same workload, committed in profiles/r02_traffic.json; "
                            "not measured in this run -- ncu cannot run inside the timed process)",
            "hbm_frac_if_memory_bound": (alg_bytes / (m["tree_ms_per_step"] * 1e-3) * 1e-9 / hbm) if hbm else None,
            "interp_avg_ms": m["interp_ms_per_step"] / max(m["interp_launches_per_step"], 1),
        })
    # number of matrix blocks the interpolation kernel wrote in the last step: shared-matrix keys that were not identity
    # (age-scaled keys: one block per generation-grid index touched; counted as two-source blocks)
    if rank == 0:
        keys_fixed = [(ped.varname_of[b], bool(ped.br_bottleneck[i])) for i, b in enumerate(ped.branch_names) if ped.br_mult[i] < 0]
        uniq_fixed = sorted(set(keys_fixed))
        N = tb.N
        nblocks = 0
        for x in X:
            for (v, bott) in uniq_fixed:
                ln = 10 ** x[inf.var_indices[v]]
                if bott or not (N * ln < tb.gens[0]):
                    nblocks += 1
        roofline["interp"] = interp_roofline(m, F, Fp, nblocks, 0) if args.tree == "fixed" else None

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        kind = build_cpu_inference(ds, main_ref, workdir)
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
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms"], "higher_is_better": True, "scaling": "strong" if locus else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, F), "walkers_total": W_total, "rows_per_walker": R_local, "branches": B,
                       "drift_grid": "%dx%d" % (tb.drift.shape[0], tb.drift.shape[1]),
                       "parallelism": ("locus-sharded x%d (families partitioned; one NCCL all-reduce of %d doubles per evaluation)" % (world, W)
                                       if locus else "walker-sharded x%d" % world),
                       "l2_policy": "inputs larger than L2 (emissions %.0f MB + per-walker matrices %.0f MB per step)" % (
                           R_local * len(inf.leaf_names[0]) * Fp * 8 / 1e6, W * n_keys * Fp * (Fp + 3) * 8 / 1e6),
                       "value_note": "value = device-resident step: tables, emissions and the walkers' root prior (a function of two of "
                                     "the parameters) already in HBM; e2e includes the root prior (device kernel; SciPy's betainc on the "
                                     "host for walkers with ab < 1e-6, pipelined against the device work)",
                       "finite_walkers": int(np.isfinite(ll_host).sum())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "max_rel_diff_vs_value_leg": m["legs_max_rel_diff"], "legs_retried": m["legs_retried"],
                    "ms_per_step": m["e2e_s"] * 1e3},
            "gpu_launches": int(m["launches_per_step"] * args.steps),
            "clocks": m["clocks"],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "per_rank": per_rank,
        }
    inf.close()
    del inf, eng
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations ----
    wanted = [c.strip() for c in args.configs.split(",")] if args.configs != "all" else ["cfg2", "cfg3", "cfg4", "cfg5"]
    extras = []
    if is_default_headline(args):
        pk = D.max_over_ranks(peak if peak is not None else 0.0)     # every rank needs the same call sequence; rank 0 measured it
        for name in ("cfg3", "cfg4", "cfg5"):
            if name not in wanted:
                continue
            try:
                r = run_extra_config(D, name, tb, workdir, pk, not args.no_cpu_baseline, args.steps)
            except Exception as e:   # an extra configuration must never cost the headline line
                import traceback
                traceback.print_exc(file=sys.stderr)
                r = {"name": name, "error": repr(e)} if rank == 0 else None
                if world > 1:
                    raise
            if r is not None:
                extras.append(r)
    if rank == 0:
        line["configs"] = extras
        print(json.dumps(line))
    import shutil
    shutil.rmtree(workdir, ignore_errors=True)
    D.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Device engine: owns the CUDA buffers (through PyTorch) and drives the C ABI.

PyTorch is plumbing only here -- device memory, streams, pinned host buffers.  All compute is in
csrc/ (hand-written sm_100a kernels) behind include/mope_b200.h.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .model import PedigreeModel, fam_lgamma
from .tables import TransitionTables


class OutOfTableError(ValueError):
    """Parameters outside the transition tables (the reference raises ValueError from
    TransitionData.get_transition_probabilities_2d_not_cached)."""


def status_message(status: int) -> str:
    return "; ".join(msg for bit, msg in _lib.ST_MESSAGES.items() if status & bit) or "ok"


class Engine:
    def __init__(self, tables: TransitionTables, peds: Sequence[PedigreeModel], var_index: Sequence[dict], n_var: int,
                 min_phred_score: Optional[float], min_freq: float, genome_size: float, poisson_like_penalty: float,
                 device: int = 0, workspace_bytes: Optional[int] = None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("mope_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = _lib.lib()
        self.device = int(device)
        self.tables = tables
        self.F = tables.F
        self.n_var = int(n_var)
        self.ndim = 2 * self.n_var + 2
        self.n_ped = len(peds)
        self.asc_sizes = [p.n_asc for p in peds]
        self._keep = []  # host arrays referenced by the descriptors during create

        td = _lib.TablesDesc()
        td.F = self.F
        td.n_gens, td.n_us = tables.drift.shape[0], tables.drift.shape[1]
        td.drift, td.gens, td.us = _lib.dptr(tables.drift), _lib.dptr(tables.gens), _lib.dptr(tables.us)
        if tables.bot is not None:
            td.n_nbs, td.n_bot_us = tables.bot.shape[0], tables.bot.shape[1]
            td.bot, td.nbs, td.bot_us = _lib.dptr(tables.bot), _lib.dptr(tables.nbs), _lib.dptr(tables.bot_us)
        else:
            td.n_nbs = td.n_bot_us = 0
            td.bot = td.nbs = td.bot_us = None
        td.N = float(tables.N)
        td.freqs = _lib.dptr(tables.freqs)
        self._breaks32 = np.ascontiguousarray(tables.breaks, dtype=np.int32)
        td.breaks = _lib.iptr(self._breaks32)

        pds = (_lib.PedigreeDesc * self.n_ped)()
        for k, (p, vidx) in enumerate(zip(peds, var_index)):
            d = pds[k]
            d.n_leaves, d.n_branches = len(p.leaf_names), len(p.branch_names)
            d.n_loci, d.n_fam, d.n_mult, d.n_asc = p.num_loci, p.n_fam, len(p.multipliernames), p.n_asc
            arrs = dict(counts=np.ascontiguousarray(p.counts, dtype=np.float64),
                        coverage=np.ascontiguousarray(p.coverage, dtype=np.float64),
                        row_mult=np.ascontiguousarray(p.row_mult, dtype=np.float64),
                        asc_mult=np.ascontiguousarray(p.asc_mult, dtype=np.float64),
                        fam_count=np.ascontiguousarray(p.fam_count, dtype=np.float64),
                        fam_lgamma=fam_lgamma(p.fam_count))
            iarrs = dict(fam_asc=np.ascontiguousarray(p.fam_asc, dtype=np.int32),
                         br_parent=np.ascontiguousarray(p.br_parent, dtype=np.int32),
                         br_leaf=np.ascontiguousarray(p.br_leaf, dtype=np.int32),
                         br_var=np.array([vidx[p.varname_of[b]] for b in p.branch_names], dtype=np.int32),
                         br_mult=np.ascontiguousarray(p.br_mult, dtype=np.int32),
                         br_bottleneck=np.ascontiguousarray(p.br_bottleneck, dtype=np.int32))
            self._keep.append((arrs, iarrs))
            for name, a in arrs.items():
                setattr(d, name, _lib.dptr(a))
            for name, a in iarrs.items():
                setattr(d, name, _lib.iptr(a))
        md = _lib.ModelDesc()
        md.n_var, md.n_ped, md.peds = self.n_var, self.n_ped, pds
        md.min_phred_score = -1.0 if min_phred_score is None else float(min_phred_score)
        md.min_freq, md.genome_size, md.poisson_like_penalty = float(min_freq), float(genome_size), float(poisson_like_penalty)

        nbytes = C.c_size_t(0)
        _lib.check(self.lib.mope_b200_arena_bytes(C.byref(td), C.byref(md), C.byref(nbytes)), "mope_b200_arena_bytes")
        with torch.cuda.device(self.device):
            self.arena = torch.empty(int(nbytes.value), dtype=torch.uint8, device="cuda:%d" % self.device)
            ctx = C.c_void_p()
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.mope_b200_create(C.byref(td), C.byref(md), self.device, self.arena.data_ptr(),
                                                 self.arena.numel(), stream, C.byref(ctx)), "mope_b200_create")
        self.ctx = ctx
        self._ws = None
        self._ws_ok = set()
        self._ws_limit = workspace_bytes
        self._keep = []

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.mope_b200_destroy(self.ctx)
            self.ctx = None
        self._ws = None
        self.arena = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self.lib.mope_b200_launch_count(self.ctx))

    def _workspace(self, W: int):
        """Device workspace for W walkers (grow-only).  The size query and cudaMemGetInfo run only the first time a
        batch size is seen: cudaMemGetInfo stalls for tens of milliseconds while kernels are in flight."""
        torch = self.torch
        if self._ws is not None and W in self._ws_ok:
            return self._ws
        mn, full = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self.lib.mope_b200_workspace_bytes(self.ctx, int(W), C.byref(mn), C.byref(full)), "mope_b200_workspace_bytes")
        want = int(full.value)
        if self._ws is None or self._ws.numel() < want:
            if self._ws_limit is not None:
                want = max(int(mn.value), min(want, int(self._ws_limit)))
            else:
                free, _total = torch.cuda.mem_get_info(self.device)
                reuse = self._ws.numel() if self._ws is not None else 0
                want = max(int(mn.value), min(want, int(0.85 * (free + reuse))))
            if self._ws is None or self._ws.numel() < want:
                self._ws = None
                with torch.cuda.device(self.device):
                    self._ws = torch.empty(want, dtype=torch.uint8, device="cuda:%d" % self.device)
        self._ws_ok.add(W)
        return self._ws

    # ------------------------------------------------------------------------------------------
    def loglike(self, params: np.ndarray, stat: np.ndarray, want_components: bool = False):
        """params [W, ndim], stat [W, F] host arrays -> (ll [W], status [W][, root_ll [W, n_ped], log_asc [W, sum n_asc]])."""
        torch = self.torch
        params = np.ascontiguousarray(params, dtype=np.float64)
        stat = np.ascontiguousarray(stat, dtype=np.float64)
        W = params.shape[0]
        assert params.shape == (W, self.ndim) and stat.shape == (W, self.F)
        ll = np.empty(W)
        status = np.zeros(W, dtype=np.int32)
        root = np.empty((W, self.n_ped)) if want_components else None
        lasc = np.empty((W, sum(self.asc_sizes))) if want_components else None
        if W == 0:
            return (ll, status, root, lasc) if want_components else (ll, status)
        ws = self._workspace(W)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self.lib.mope_b200_eval(self.ctx, _lib.dptr(params), _lib.dptr(stat), W, ws.data_ptr(), ws.numel(), stream,
                                         _lib.dptr(ll), _lib.iptr(status),
                                         _lib.dptr(root) if want_components else None,
                                         _lib.dptr(lasc) if want_components else None)
        _lib.check(rc, "mope_b200_eval")
        return (ll, status, root, lasc) if want_components else (ll, status)

    def loglike_pipelined(self, params: np.ndarray, stat_fn, n_chunks: int = 4):
        """End-to-end evaluation with the host-side root distributions overlapped with the device work: the batch is
        cut into `n_chunks`; while the device evaluates chunk i (enqueued, not awaited) the host computes stat_fn for
        chunk i+1.  One device->host read of the log-likelihoods at the end.  Returns (ll, status)."""
        torch = self.torch
        out_dev, status = self.loglike_enqueue(params, stat_fn, n_chunks)
        W = out_dev.shape[0]
        with torch.cuda.device(self.device):
            self._out_pin[:W].copy_(out_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return self._out_pin[:W].numpy().copy(), status

    def loglike_enqueue(self, params: np.ndarray, stat_fn, n_chunks: int = 4):
        """As loglike_pipelined, but the log-likelihoods stay on the device: returns (CUDA float64 tensor [W] that the
        current stream will have written, host status [W]) without synchronising.  The tensor is engine-owned scratch,
        valid until the next evaluation."""
        torch = self.torch
        params = np.ascontiguousarray(params, dtype=np.float64)
        W = params.shape[0]
        dev = "cuda:%d" % self.device
        if getattr(self, "_pipe_W", 0) < W:
            self._stat_pin = torch.empty((W, self.F), dtype=torch.float64, pin_memory=True)
            self._stat_dev = torch.empty((W, self.F), dtype=torch.float64, device=dev)
            self._out_dev = torch.empty(W, dtype=torch.float64, device=dev)
            self._out_pin = torch.empty(W, dtype=torch.float64, pin_memory=True)
            self._pipe_W = W
        status = np.zeros(W, dtype=np.int32)
        # growing chunks: only the first (small) chunk's host work is exposed, every later one hides under the
        # device work of its predecessor
        fr = np.array([0.0, 1.0, 4.0, 8.0, 16.0]) / 16.0 if n_chunks == 4 else np.linspace(0, 1, n_chunks + 1)
        bounds = np.unique(np.round(fr * W).astype(int))
        stat_pin_np = self._stat_pin.numpy()
        with torch.cuda.device(self.device):
            for i in range(len(bounds) - 1):
                lo, hi = int(bounds[i]), int(bounds[i + 1])
                if hi <= lo:
                    continue
                stat_pin_np[lo:hi] = stat_fn(params[lo:hi])
                self._stat_dev[lo:hi].copy_(self._stat_pin[lo:hi], non_blocking=True)
                status[lo:hi] = self.loglike_device(params[lo:hi], self._stat_dev[lo:hi], self._out_dev[lo:hi])
        return self._out_dev[:W], status

    def root_prior_device(self, params: np.ndarray, out_dev=None):
        """Root distributions of the walkers on the device (mope_b200_root_prior): params [W, ndim] host array ->
        CUDA float64 tensor [W, F] (engine-owned scratch unless out_dev is given).  Enqueue only."""
        torch = self.torch
        params = np.ascontiguousarray(params, dtype=np.float64)
        W = params.shape[0]
        dev = "cuda:%d" % self.device
        if getattr(self, "_rp_W", 0) < W:
            self._rp_pin = torch.empty((W, self.ndim), dtype=torch.float64, pin_memory=True)
            self._rp_par = torch.empty((W, self.ndim), dtype=torch.float64, device=dev)
            self._rp_out = torch.empty((W, self.F), dtype=torch.float64, device=dev)
            self._rp_W = W
        out = self._rp_out[:W] if out_dev is None else out_dev
        with torch.cuda.device(self.device):
            # the pinned staging row block is reused: the previous copy out of it must have completed
            if getattr(self, "_rp_busy", None) is not None:
                self._rp_busy.synchronize()
            self._rp_pin.numpy()[:W] = params
            self._rp_par[:W].copy_(self._rp_pin[:W], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._rp_busy = ev
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.mope_b200_root_prior(self.ctx, self._rp_par.data_ptr(), W, stream, out.data_ptr()),
                       "mope_b200_root_prior")
        return out

    def loglike_hybrid(self, params: np.ndarray, host_mask: np.ndarray, stat_fn, keep_on_device: bool = False):
        """End-to-end evaluation with the root prior on the device for the walkers where host_mask is False and from
        stat_fn (host, SciPy) where it is True.  The device-prior walkers are enqueued first, so the host work of the others
        runs while the device is busy.  Returns (ll, status)."""
        torch = self.torch
        params = np.ascontiguousarray(params, dtype=np.float64)
        W = params.shape[0]
        dev = "cuda:%d" % self.device
        if getattr(self, "_pipe_W", 0) < W:
            self._stat_pin = torch.empty((W, self.F), dtype=torch.float64, pin_memory=True)
            self._stat_dev = torch.empty((W, self.F), dtype=torch.float64, device=dev)
            self._out_dev = torch.empty(W, dtype=torch.float64, device=dev)
            self._out_pin = torch.empty(W, dtype=torch.float64, pin_memory=True)
            self._pipe_W = W
        status = np.zeros(W, dtype=np.int32)
        idx_d = np.nonzero(~host_mask)[0]
        idx_h = np.nonzero(host_mask)[0]
        nd, nh = idx_d.shape[0], idx_h.shape[0]
        with torch.cuda.device(self.device):
            if nd:
                Xd = params[idx_d]
                self.root_prior_device(Xd, out_dev=self._stat_dev[:nd])
                status[idx_d] = self.loglike_device(Xd, self._stat_dev[:nd], self._out_dev[:nd])
            if nh:
                Xh = params[idx_h]
                self._stat_pin.numpy()[nd:nd + nh] = stat_fn(Xh)
                self._stat_dev[nd:nd + nh].copy_(self._stat_pin[nd:nd + nh], non_blocking=True)
                status[idx_h] = self.loglike_device(Xh, self._stat_dev[nd:nd + nh], self._out_dev[nd:nd + nh])
            if keep_on_device:
                # walker order on the device, no synchronisation: (CUDA tensor [W], host status)
                if nh == 0:
                    return self._out_dev[:W], status
                order = torch.as_tensor(np.concatenate((idx_d, idx_h)), device=dev)
                sorted_out = torch.empty(W, dtype=torch.float64, device=dev)
                sorted_out[order] = self._out_dev[:W]
                return sorted_out, status
            self._out_pin[:W].copy_(self._out_dev[:W], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        ll = np.empty(W)
        got = self._out_pin[:W].numpy()
        ll[idx_d] = got[:nd]
        ll[idx_h] = got[nd:nd + nh]
        return ll, status

    def loglike_resident(self, params_dev, stat_dev=None, lnprior_dev=None, out_dev=None, status_dev=None):
        """Device-resident evaluation (mope_b200_eval_resident): params_dev [W, ndim] CUDA float64 tensor of sign-folded
        parameters; the interpolation plans and (when stat_dev is None) the root distributions are derived on the device.
        Walkers whose lnprior_dev entry is -inf are skipped.  Returns (ll CUDA tensor [W], status CUDA int32 tensor [W]);
        enqueue only, nothing touches the host."""
        torch = self.torch
        assert params_dev.is_cuda and params_dev.dtype == torch.float64 and params_dev.is_contiguous()
        W = params_dev.shape[0]
        dev = params_dev.device
        out_dev = torch.empty(W, dtype=torch.float64, device=dev) if out_dev is None else out_dev
        status_dev = torch.empty(W, dtype=torch.int32, device=dev) if status_dev is None else status_dev
        if W == 0:
            return out_dev, status_dev
        ws = self._workspace(W)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self.lib.mope_b200_eval_resident(self.ctx, params_dev.data_ptr(),
                                                  None if stat_dev is None else stat_dev.data_ptr(),
                                                  None if lnprior_dev is None else lnprior_dev.data_ptr(),
                                                  W, ws.data_ptr(), ws.numel(), stream, out_dev.data_ptr(), status_dev.data_ptr())
        _lib.check(rc, "mope_b200_eval_resident")
        return out_dev, status_dev

    def loglike_device(self, params_host: np.ndarray, stat_dev, out_dev):
        """Device-resident variant: stat_dev [W, F] and out_dev [W] are CUDA float64 tensors.  Enqueue only."""
        torch = self.torch
        params_host = np.ascontiguousarray(params_host, dtype=np.float64)
        W = params_host.shape[0]
        status = np.zeros(W, dtype=np.int32)
        ws = self._workspace(W)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = self.lib.mope_b200_eval_device(self.ctx, _lib.dptr(params_host), stat_dev.data_ptr(), W, ws.data_ptr(),
                                                ws.numel(), stream, out_dev.data_ptr(), _lib.iptr(status))
        _lib.check(rc, "mope_b200_eval_device")
        return status

    # measurement support ------------------------------------------------------------------------
    def set_profiling(self, on: bool) -> None:
        _lib.check(self.lib.mope_b200_set_profiling(self.ctx, int(bool(on))), "mope_b200_set_profiling")

    def last_eval_stats(self):
        """(gemm_ops, identity_ops) of the last evaluation, counted per (walker, branch)."""
        g, i = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.mope_b200_last_eval_stats(self.ctx, C.byref(g), C.byref(i)), "mope_b200_last_eval_stats")
        return g.value, i.value

    def last_eval_unit_rows(self) -> int:
        """Sum over the GEMM units the pruning kernel executed in the last evaluation of the unit's real rows
        (2 * F^2 * this = executed flops); synchronises."""
        v = C.c_int64(0)
        _lib.check(self.lib.mope_b200_last_eval_unit_rows(self.ctx, C.byref(v)), "mope_b200_last_eval_unit_rows")
        return int(v.value)

    def last_eval_segment_stats(self):
        """(warp x generation-segment pairs of age-scaled branches the pruning kernel went through, those it multiplied in)."""
        a, b = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.mope_b200_last_eval_segment_stats(self.ctx, C.byref(a), C.byref(b)), "mope_b200_last_eval_segment_stats")
        return int(a.value), int(b.value)

    def kernel_times(self):
        """(tree_ms, tree_launches, interp_ms, interp_launches) accumulated since the last call; synchronises."""
        tm, im = C.c_double(0), C.c_double(0)
        tn, in_ = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.mope_b200_kernel_times(self.ctx, C.byref(tm), C.byref(tn), C.byref(im), C.byref(in_)),
                   "mope_b200_kernel_times")
        return tm.value, tn.value, im.value, in_.value

    def fp64_peak_tflops(self, millis: float = 300.0) -> float:
        out = C.c_double(0)
        _lib.check(self.lib.mope_b200_fp64_peak(self.ctx, float(millis), self._stream(), C.byref(out)), "mope_b200_fp64_peak")
        return out.value

    # operator-level entry points ------------------------------------------------------------------
    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def transition_matrix(self, x: float, scaled_mut: float, bottleneck: bool = False) -> np.ndarray:
        P = np.empty((self.F, self.F))
        st = C.c_int32(0)
        _lib.check(self.lib.mope_b200_transition_matrix(self.ctx, int(bool(bottleneck)), float(x), float(scaled_mut),
                                                        self._stream(), _lib.dptr(P), C.byref(st)), "mope_b200_transition_matrix")
        if st.value:
            raise OutOfTableError(status_message(st.value))
        return P

    def leaf_likelihoods(self, counts, coverage, freqs, min_phred_score=-1.0) -> np.ndarray:
        counts = np.ascontiguousarray(counts, dtype=np.float64)
        coverage = np.ascontiguousarray(coverage, dtype=np.float64)
        freqs = np.ascontiguousarray(freqs, dtype=np.float64)
        L, S = counts.shape
        out = np.empty((L, S, freqs.shape[0]))
        _lib.check(self.lib.mope_b200_leaf_likelihoods(self.ctx, _lib.dptr(counts), _lib.dptr(coverage), L, S, _lib.dptr(freqs),
                                                       freqs.shape[0], float(min_phred_score), self._stream(), _lib.dptr(out)),
                   "mope_b200_leaf_likelihoods")
        return out

    def branch_update(self, kind: int, child, parent, lengths, scaled_mut: float) -> None:
        """In place: parent[r] *= P_r @ child[r] (kind 0 fixed, 1 rate, 2 bottleneck)."""
        child = np.ascontiguousarray(child, dtype=np.float64)
        assert parent.flags["C_CONTIGUOUS"] and parent.dtype == np.float64 and parent.shape == child.shape
        lengths = np.ascontiguousarray(np.atleast_1d(lengths), dtype=np.float64)
        st = C.c_int32(0)
        _lib.check(self.lib.mope_b200_branch_update(self.ctx, int(kind), _lib.dptr(child), _lib.dptr(parent), child.shape[0],
                                                    _lib.dptr(lengths), float(scaled_mut), self._stream(), C.byref(st)),
                   "mope_b200_branch_update")
        if st.value:
            raise OutOfTableError(status_message(st.value))

    def non_asc_probs(self, quality_sets: List[np.ndarray], freqs, min_empirical_freq: float) -> np.ndarray:
        freqs = np.ascontiguousarray(freqs, dtype=np.float64)
        off = np.zeros(len(quality_sets) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(q) for q in quality_sets])
        q = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64) for s in quality_sets])
                                 if quality_sets else np.zeros(0))
        if q.shape[0] == 0:
            q = np.zeros(1)
        out = np.empty((len(quality_sets), freqs.shape[0]))
        _lib.check(self.lib.mope_b200_non_asc_probs(self.ctx, _lib.dptr(q), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                                    len(quality_sets), _lib.dptr(freqs), freqs.shape[0], float(min_empirical_freq),
                                                    self._stream(), _lib.dptr(out)), "mope_b200_non_asc_probs")
        return out

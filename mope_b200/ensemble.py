"""Affine-invariant ensemble sampler (Goodman & Weare stretch move) -- the part of emcee that mope's
`Inference.run_mcmc` relies on (mope/inference.py:1350-1363: `emcee.EnsembleSampler(num_walkers, ndim, post_clone,
pool=pool, a=chain_alpha)` iterated with `sample(init_pos, iterations, store=False)`).

emcee is neither vendored in the reference nor installed here; this is a restatement of its published stretch move:
the ensemble is split into two halves, every walker of a half proposes  y = c + z (x - c)  with a partner c drawn from
the other half and z ~ g(z) ∝ 1/sqrt(z) on [1/a, a], accepted with probability min(1, z^(ndim-1) p(y)/p(x)).
Each half-step is ONE batched posterior evaluation (what emcee does with `pool.map`), i.e. one device call.
"""
from __future__ import annotations

from typing import Callable, Iterator, Optional, Tuple

import numpy as np


class EnsembleSampler(object):
    def __init__(self, nwalkers: int, ndim: int, log_prob_batch: Callable[[np.ndarray], np.ndarray], a: float = 2.0,
                 seed: Optional[int] = None, randomize_split: bool = True):
        if nwalkers < 2 * ndim or nwalkers % 2:
            raise ValueError("the stretch move needs an even number of walkers, at least 2*ndim (got {} for ndim {})".format(
                nwalkers, ndim))
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_batch = log_prob_batch
        self.random = np.random.RandomState(seed)
        self.randomize_split = randomize_split
        self.naccepted = np.zeros(self.nwalkers)
        self.iterations = 0

    @property
    def acceptance_fraction(self) -> np.ndarray:
        return self.naccepted / max(self.iterations, 1)

    def sample(self, p0: np.ndarray, lnprob0: Optional[np.ndarray] = None, iterations: int = 1
               ) -> Iterator[Tuple[np.ndarray, np.ndarray, tuple]]:
        p = np.array(p0, dtype=np.float64, copy=True)
        if p.shape != (self.nwalkers, self.ndim):
            raise ValueError("initial positions must have shape (nwalkers, ndim)")
        lnprob = np.asarray(self.log_prob_batch(p), dtype=np.float64) if lnprob0 is None else np.array(lnprob0, dtype=np.float64)
        if np.any(np.isnan(lnprob)):
            raise ValueError("The initial lnprob was NaN.")
        half = self.nwalkers // 2
        for _ in range(int(iterations)):
            self.iterations += 1
            idx = self.random.permutation(self.nwalkers) if self.randomize_split else np.arange(self.nwalkers)
            for S, C in ((idx[:half], idx[half:]), (idx[half:], idx[:half])):
                s, c = p[S], p[C]
                zz = ((self.a - 1.0) * self.random.rand(len(S)) + 1) ** 2.0 / self.a
                factors = (self.ndim - 1.0) * np.log(zz)
                rint = self.random.randint(len(C), size=(len(S),))
                q = c[rint] - (c[rint] - s) * zz[:, None]
                newlp = np.asarray(self.log_prob_batch(q), dtype=np.float64)
                newlp = np.where(np.isnan(newlp), -np.inf, newlp)
                lnpdiff = factors + newlp - lnprob[S]
                accepted = np.log(self.random.rand(len(S))) < lnpdiff
                p[S[accepted]] = q[accepted]
                lnprob[S[accepted]] = newlp[accepted]
                self.naccepted[S[accepted]] += 1
            yield p.copy(), lnprob.copy(), self.random.get_state()


class DeviceEnsembleSampler(object):
    """The same stretch move with the whole step resident on the GPU (mope_b200_mcmc_halfstep): positions, proposals,
    sign folding + box prior, root prior, interpolation plans, likelihood and accept/reject never leave HBM; the host only
    enqueues two calls per iteration and reads back the ensemble it prints (pinned ring, a few iterations behind the
    device).  The ensemble is split into fixed halves (emcee 2.x behaviour, `randomize_split=False` of the host sampler).

    Random numbers come from a counter-based generator on the device (Philox4x32-10 keyed by `seed`); `streams`
    (a numpy RandomState) instead replays exactly the draws the host EnsembleSampler would make, which is how the two are
    tested against each other.

    The root distribution is always the device kernel here (Inference(root_prior="device") semantics): walkers with
    ab < 1e-6 see cell masses that are MORE accurate than the reference's CDF differences; their log-likelihood differs
    from the reference's by up to ~3e-10 relative at ab ~ 1e-9."""

    def __init__(self, inf, nwalkers: int, a: float = 2.0, seed: Optional[int] = None, streams=None, depth: int = 4):
        import ctypes as C
        import torch
        from . import _lib
        if nwalkers < 2 * inf.ndim or nwalkers % 2:
            raise ValueError("the stretch move needs an even number of walkers, at least 2*ndim (got {} for ndim {})".format(
                nwalkers, inf.ndim))
        if inf.inverse_bot_priors or inf.post_is_prior:
            raise NotImplementedError("the device sampler covers the box prior (log-uniform or uniform drift); use the host sampler "
                                      "for --inverse-bottleneck-priors / --post-is-prior")
        self.inf, self.eng = inf, inf.engine
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(inf.ndim), float(a)
        self.seed = int(np.random.SeedSequence(seed).generate_state(1, dtype=np.uint64)[0]) if seed is None else int(seed)
        self.streams = streams
        self.depth = int(depth)
        self.iterations = 0
        self.torch, self.C, self._lib = torch, C, _lib
        dev = "cuda:%d" % self.eng.device
        W, nd, half = self.nwalkers, self.ndim, self.nwalkers // 2
        f64 = torch.float64
        self.pos = torch.empty((W, nd), dtype=f64, device=dev)
        self.lnprob = torch.empty(W, dtype=f64, device=dev)
        self.naccept = torch.zeros(W, dtype=torch.int64, device=dev)
        self._scratch = dict(q=torch.empty((half, nd), dtype=f64, device=dev), folded=torch.empty((half, nd), dtype=f64, device=dev),
                             factor=torch.empty(half, dtype=f64, device=dev), lnprior=torch.empty(half, dtype=f64, device=dev),
                             ll=torch.empty(half, dtype=f64, device=dev), status=torch.empty(half, dtype=torch.int32, device=dev))
        # log prior inside the box = const + slope . x  (inference.py:1055-1110 without the inverse-bottleneck option)
        nv = inf.num_varnames
        u, l = inf.upper, inf.lower
        slope = np.zeros(nd)
        if inf.log_unif_drift:
            const = float(np.sum(np.log(1.0 / (u - l))))
        else:
            slope[:nv] = 2.3025850929940459
            const = float(np.sum(0.834032445247956 - np.log(10 ** u[:nv] - 10 ** l[:nv])) + np.sum(np.log(1.0 / (u[nv:2 * nv] - l[nv:2 * nv])))
                          + np.sum(np.log(1.0 / (u[2 * nv:] - l[2 * nv:]))))
        self._lower = torch.as_tensor(np.ascontiguousarray(l, dtype=np.float64), device=dev)
        self._upper = torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64), device=dev)
        self._slope = torch.as_tensor(slope, device=dev)
        self._const = const
        self._ring_pos = torch.empty((self.depth, W, nd), dtype=f64, pin_memory=True)
        self._ring_lp = torch.empty((self.depth, W), dtype=f64, pin_memory=True)
        self.status_or = torch.zeros(half, dtype=torch.int32, device=dev)   # OR of the MOPE_ST_* bits of in-prior proposals over the run (0 = none left the tables)

    @property
    def acceptance_fraction(self) -> np.ndarray:
        return self.naccept.cpu().numpy() / max(self.iterations, 1)

    def _state(self, step: int, zu=None, partner=None, au=None):
        s = self._lib.SamplerState()
        sc = self._scratch
        s.pos, s.lnprob, s.naccept = self.pos.data_ptr(), self.lnprob.data_ptr(), self.naccept.data_ptr()
        s.q, s.folded, s.factor = sc["q"].data_ptr(), sc["folded"].data_ptr(), sc["factor"].data_ptr()
        s.lnprior, s.ll, s.status = sc["lnprior"].data_ptr(), sc["ll"].data_ptr(), sc["status"].data_ptr()
        s.lower, s.upper, s.prior_slope, s.prior_const = self._lower.data_ptr(), self._upper.data_ptr(), self._slope.data_ptr(), self._const
        s.zu = None if zu is None else zu.data_ptr()
        s.partner = None if partner is None else partner.data_ptr()
        s.au = None if au is None else au.data_ptr()
        s.seed, s.step, s.a, s.n_walkers = self.seed & (2 ** 64 - 1), int(step), self.a, self.nwalkers
        return s

    def sample(self, p0: np.ndarray, lnprob0: Optional[np.ndarray] = None, iterations: int = 1):
        torch, C = self.torch, self.C
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        W, nd, half = self.nwalkers, self.ndim, self.nwalkers // 2
        if p0.shape != (W, nd):
            raise ValueError("initial positions must have shape (nwalkers, ndim)")
        if lnprob0 is None:
            mode = self.inf.root_prior
            self.inf.root_prior = "device"
            try:
                lnprob0 = self.inf.log_posterior_batch(p0)
            finally:
                self.inf.root_prior = mode
        lnprob0 = np.asarray(lnprob0, dtype=np.float64)
        if np.any(np.isnan(lnprob0)):
            raise ValueError("The initial lnprob was NaN.")
        dev = self.pos.device
        self.pos.copy_(torch.as_tensor(p0))
        self.lnprob.copy_(torch.as_tensor(lnprob0))
        iterations = int(iterations)
        zu = partner = au = None
        if self.streams is not None:
            # exactly the draws of EnsembleSampler.sample(randomize_split=False): per half-step rand(half), randint(half), rand(half)
            rs = self.streams
            Z = np.empty((iterations, 2, half))
            Pn = np.empty((iterations, 2, half), dtype=np.int32)
            A = np.empty((iterations, 2, half))
            for it in range(iterations):
                for h in range(2):
                    Z[it, h] = rs.rand(half)
                    Pn[it, h] = rs.randint(half, size=(half,))
                    A[it, h] = rs.rand(half)
            zu, partner, au = (torch.as_tensor(x, device=dev) for x in (Z, Pn, A))
        ws = self.eng._workspace(half)
        lib, ctx = self.eng.lib, self.eng.ctx
        events = [None] * self.depth
        with torch.cuda.device(self.eng.device):
            stream = torch.cuda.current_stream()

            def enqueue(it):
                for h in range(2):
                    st = self._state(2 * (self.iterations + it) + h,
                                     None if zu is None else zu[it, h], None if partner is None else partner[it, h],
                                     None if au is None else au[it, h])
                    self._lib.check(lib.mope_b200_mcmc_halfstep(ctx, C.byref(st), h, ws.data_ptr(), ws.numel(), stream.cuda_stream),
                                    "mope_b200_mcmc_halfstep")
                    stt = self._scratch["status"]
                    self.status_or |= torch.where((stt & 128) != 0, torch.zeros_like(stt), stt)   # proposals outside the prior do not count
                slot = it % self.depth
                self._ring_pos[slot].copy_(self.pos, non_blocking=True)
                self._ring_lp[slot].copy_(self.lnprob, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
                events[slot] = ev

            ahead = min(self.depth - 1, iterations)
            for it in range(ahead):
                enqueue(it)
            for it in range(iterations):
                slot = it % self.depth
                events[slot].synchronize()
                p, lp = self._ring_pos[slot].numpy().copy(), self._ring_lp[slot].numpy().copy()
                if it + ahead < iterations:
                    enqueue(it + ahead)          # the slot being refilled is (it + ahead) % depth != slot
                yield p, lp, None
            self.iterations += iterations

#!/usr/bin/env python
"""Generator kernels on one B200: dgemm_tn_kernel (+ transpose) against cuBLAS DGEMM (torch.matmul, comparison only) at the
Wright-Fisher size, and the wall time of a table grid.  usage: python profiles/tools/bench_generator.py [--full]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mope_b200 import generate  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    mat = generate._Mat("cuda")
    out = {}
    for n in (1001, 2001):
        rng = np.random.RandomState(0)
        a, b = mat.put(rng.rand(n, n)), mat.put(rng.rand(n, n))
        ta, tb = a[:, :n].contiguous(), b[:, :n].contiguous()
        ld = a.shape[1]
        bt = mat._new(n)
        c = mat._new(n)
        st = torch.cuda.current_stream().cuda_stream
        gemm_only = lambda: mat.lib.mope_b200_gen_dgemm_tn(a.data_ptr(), bt.data_ptr(), c.data_ptr(), n, n, n, ld, ld, ld, st)  # noqa: E731
        ms_k = timed(gemm_only, 50)
        ms_mm = timed(lambda: mat.mm(a, b), 50)
        ms_cublas = timed(lambda: torch.matmul(ta, tb), 50)
        fl = 2.0 * n ** 3
        out["n%d" % n] = {"dgemm_tn_kernel_ms": ms_k, "dgemm_tn_tflops": fl / ms_k * 1e-9, "mm_with_transpose_ms": ms_mm,
                          "mm_tflops": fl / ms_mm * 1e-9, "cublas_dgemm_ms": ms_cublas, "cublas_tflops": fl / ms_cublas * 1e-9}
    t0 = time.perf_counter()
    full = "--full" in sys.argv
    gens = generate.DEFAULT_GENS
    us = generate.DEFAULT_MUTS if full else generate.DEFAULT_MUTS[::11]
    nbs = generate.DEFAULT_BOTS
    tb = generate.generate_tables(N=1000, breaks=(0.5, 0.005), gens=gens, us=us, nbs=nbs, bot_us=us, device="cuda")
    torch.cuda.synchronize()
    out["grid"] = {"gens": len(gens), "us": len(us), "nbs": len(nbs), "F": int(tb.F), "seconds": time.perf_counter() - t0}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

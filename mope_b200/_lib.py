"""ctypes binding of the C ABI in include/mope_b200.h.  Fails loudly when the CUDA library is missing:
there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libmope_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile mope_b200/csrc into libmope_b200.so in-tree with nvcc for sm_100a."""
    srcs = [os.path.join(CSRC, "mope_b200.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")] + [os.path.join(HERE, "..", "include", "mope_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + srcs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


class TablesDesc(C.Structure):
    _fields_ = [("F", C.c_int32), ("n_gens", C.c_int32), ("n_us", C.c_int32),
                ("drift", C.POINTER(C.c_double)), ("gens", C.POINTER(C.c_double)), ("us", C.POINTER(C.c_double)),
                ("n_nbs", C.c_int32), ("n_bot_us", C.c_int32),
                ("bot", C.POINTER(C.c_double)), ("nbs", C.POINTER(C.c_double)), ("bot_us", C.POINTER(C.c_double)),
                ("N", C.c_double), ("freqs", C.POINTER(C.c_double)), ("breaks", C.POINTER(C.c_int32))]


class PedigreeDesc(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("n_branches", C.c_int32), ("n_loci", C.c_int32), ("n_fam", C.c_int32),
                ("n_mult", C.c_int32), ("n_asc", C.c_int32),
                ("counts", C.POINTER(C.c_double)), ("coverage", C.POINTER(C.c_double)),
                ("row_mult", C.POINTER(C.c_double)), ("asc_mult", C.POINTER(C.c_double)),
                ("fam_asc", C.POINTER(C.c_int32)), ("fam_count", C.POINTER(C.c_double)),
                ("fam_lgamma", C.POINTER(C.c_double)),
                ("br_parent", C.POINTER(C.c_int32)), ("br_leaf", C.POINTER(C.c_int32)), ("br_var", C.POINTER(C.c_int32)),
                ("br_mult", C.POINTER(C.c_int32)), ("br_bottleneck", C.POINTER(C.c_int32))]


class ModelDesc(C.Structure):
    _fields_ = [("n_var", C.c_int32), ("n_ped", C.c_int32), ("peds", C.POINTER(PedigreeDesc)),
                ("min_phred_score", C.c_double), ("min_freq", C.c_double), ("genome_size", C.c_double),
                ("poisson_like_penalty", C.c_double)]


class SamplerState(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("lnprob", C.c_void_p), ("naccept", C.c_void_p), ("q", C.c_void_p), ("folded", C.c_void_p),
                ("factor", C.c_void_p), ("lnprior", C.c_void_p), ("ll", C.c_void_p), ("status", C.c_void_p),
                ("lower", C.c_void_p), ("upper", C.c_void_p), ("prior_slope", C.c_void_p), ("prior_const", C.c_double),
                ("zu", C.c_void_p), ("partner", C.c_void_p), ("au", C.c_void_p),
                ("seed", C.c_uint64), ("step", C.c_uint64), ("a", C.c_double), ("n_walkers", C.c_int32)]


# every symbol include/mope_b200.h declares
SYMBOLS = ["mope_b200_arena_bytes", "mope_b200_create", "mope_b200_destroy", "mope_b200_workspace_bytes",
           "mope_b200_eval", "mope_b200_eval_device", "mope_b200_eval_resident", "mope_b200_mcmc_halfstep", "mope_b200_root_prior", "mope_b200_launch_count", "mope_b200_schedule_stats", "mope_b200_last_eval_stats", "mope_b200_last_eval_unit_rows", "mope_b200_set_profiling",
           "mope_b200_kernel_times", "mope_b200_fp64_peak", "mope_b200_transition_matrix",
           "mope_b200_leaf_likelihoods", "mope_b200_branch_update", "mope_b200_non_asc_probs", "mope_b200_beta_cdf_table",
           "mope_b200_gen_binom_matrix", "mope_b200_gen_transpose", "mope_b200_gen_dgemm_tn", "mope_b200_gen_bin_matrix",
           "mope_b200_sel_branch_update", "mope_b200_sel_loglike", "mope_b200_last_eval_segment_stats",
           "mope_b200_last_error", "mope_b200_version"]

ST_MESSAGES = {
    1: "drift time too large",
    2: "mutsel rate too small",
    4: "mutsel rate too large",
    8: "bottleneck size too small",
    16: "bottleneck size too large",
    32: "scaled_time is not finite",
    64: "pre-computed bottleneck data needed",
    128: "prior is -inf (likelihood not evaluated)",
}

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MOPE_B200_LIB", LIB_PATH)   # kernel-tuning experiments load a variant build
    if not os.path.exists(path):
        raise RuntimeError("mope_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback." % path)
    L = C.CDLL(path)
    dp, vp, ip = C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int32)
    L.mope_b200_arena_bytes.argtypes = [C.POINTER(TablesDesc), C.POINTER(ModelDesc), C.POINTER(C.c_size_t)]
    L.mope_b200_create.argtypes = [C.POINTER(TablesDesc), C.POINTER(ModelDesc), C.c_int, vp, C.c_size_t, vp, C.POINTER(vp)]
    L.mope_b200_destroy.argtypes = [vp]
    L.mope_b200_workspace_bytes.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.mope_b200_eval.argtypes = [vp, dp, dp, C.c_int, vp, C.c_size_t, vp, dp, ip, dp, dp]
    L.mope_b200_eval_device.argtypes = [vp, dp, vp, C.c_int, vp, C.c_size_t, vp, vp, ip]
    L.mope_b200_root_prior.argtypes = [vp, vp, C.c_int, vp, vp]
    L.mope_b200_eval_resident.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_size_t, vp, vp, vp]
    L.mope_b200_mcmc_halfstep.argtypes = [vp, C.POINTER(SamplerState), C.c_int, vp, C.c_size_t, vp]
    L.mope_b200_launch_count.argtypes = [vp]
    L.mope_b200_launch_count.restype = C.c_int64
    L.mope_b200_schedule_stats.argtypes = [C.POINTER(ModelDesc), C.c_int, C.c_int, ip]
    L.mope_b200_last_eval_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.mope_b200_last_eval_unit_rows.argtypes = [vp, C.POINTER(C.c_int64)]
    L.mope_b200_set_profiling.argtypes = [vp, C.c_int]
    L.mope_b200_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_int64), dp, C.POINTER(C.c_int64)]
    L.mope_b200_fp64_peak.argtypes = [vp, C.c_double, vp, dp]
    L.mope_b200_transition_matrix.argtypes = [vp, C.c_int, C.c_double, C.c_double, vp, dp, ip]
    L.mope_b200_leaf_likelihoods.argtypes = [vp, dp, dp, C.c_int, C.c_int, dp, C.c_int, C.c_double, vp, dp]
    L.mope_b200_branch_update.argtypes = [vp, C.c_int, dp, dp, C.c_int, dp, C.c_double, vp, ip]
    L.mope_b200_non_asc_probs.argtypes = [vp, dp, C.POINTER(C.c_int64), C.c_int, dp, C.c_int, C.c_double, vp, dp]
    L.mope_b200_beta_cdf_table.argtypes = [vp, dp, C.c_int, dp, C.c_int, dp, C.c_int]
    L.mope_b200_gen_binom_matrix.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.mope_b200_gen_transpose.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.mope_b200_gen_dgemm_tn.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.mope_b200_gen_bin_matrix.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp]
    L.mope_b200_sel_branch_update.argtypes = [vp, C.c_int, dp, dp, C.c_int, dp, C.c_int, dp, dp, vp, ip]
    L.mope_b200_sel_loglike.argtypes = [vp, C.c_int, dp, dp, C.c_int, ip, dp, vp, dp, dp, ip]
    L.mope_b200_last_eval_segment_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.mope_b200_last_error.restype = C.c_char_p
    L.mope_b200_version.restype = C.c_char_p
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int or name in ("mope_b200_last_error", "mope_b200_version", "mope_b200_launch_count"):
            continue
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mope_b200_last_error().decode("utf-8", "replace")
        if rc == -3:
            raise MemoryError("%s: %s" % (what, msg))
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))

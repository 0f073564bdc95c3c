#!/usr/bin/env python
"""Install the UNMODIFIED-algorithm reference (ammodramus/mope) into baseline/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in mope_b200/ imports what this builds.

mope is Python + Cython and does not import on a modern stack (Cython 3, NumPy 2,
SciPy >= 1.12, pandas 3, no libgsl / h5py / emcee / lru-dict / future).  This script
copies /root/reference/mope into baseline/_ref/mope (git-ignored, like the `pip install
--target baseline/_ref` the bench contract describes), applies the one-token
compatibility patches enumerated in SURVEY.md section 8(c), drops import-only shims
next to it (oracle/ref_shims/), and cythonizes the extensions in place.  The likelihood
arithmetic is untouched; the only numerical stand-in is gsl_ran_binomial_pdf /
gsl_sf_lnchoose (GSL's published formula with lgamma), used by the leaf emissions.

Usage:  python oracle/install_ref.py [--reference /root/reference] [--force]
After it succeeds:  sys.path[:0] = ['baseline/_ref/shims', 'baseline/_ref'];  import mope
"""
from __future__ import annotations

import argparse
import os
import re
import shutil
import subprocess
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEST = os.path.join(ROOT, "baseline", "_ref")

# (file, regex, replacement) -- every edit is a spelling change forced by the toolchain
PATCHES = [
    ("_likes.pyx", r"np\.int_t", "np.int64_t"),
    ("_likes.pyx", r"np\.long_t", "np.int64_t"),
    ("_transition.pyx", r"^from _binom cimport", "from mope._binom cimport"),
    ("_poisson_binom.pyx", r"^from _wf cimport", "from mope._wf cimport"),
    ("transition_data_2d.py", r"dtype = np\.float\)", "dtype = float)"),
    ("transition_data_bottleneck.py", r"dtype = np\.int\)", "dtype = int)"),
    ("generate_transitions.py", r"\bxrange\b", "range"),
    ("generate_transitions.py", r"np\.int\b", "int"),
    ("likelihoods.py", r"^from scipy\.misc import logsumexp", "from scipy.special import logsumexp"),
    ("simulate.py", r"^from scipy\.misc import comb", "from scipy.special import comb"),
    # pandas >= 3 hands out read-only buffers; the typed memoryview in _likes needs a writable one
    ("ascertainment.py", r"counts = asc_ages\['count'\]\.values", "counts = np.array(asc_ages['count'].values)"),
    # inspect.getargspec left the standard library in Python 3.11 (used by run_mcmc to pick emcee's store keyword)
    ("inference.py", r"inspect\.getargspec\(", "inspect.getfullargspec("),
]

# bump when the patch list or the shims change: an existing install with another version is rebuilt
INSTALL_VERSION = "2"

INIT_PY = "inf_data = None\n__version__ = '0.5.1'\n"

CYTHON_MODULES = ["_util", "_binom", "_interp", "_likes", "_transition", "_wf", "_poisson_binom"]


def _patch(dest_pkg: str) -> None:
    for fn, pat, rep in PATCHES:
        path = os.path.join(dest_pkg, fn)
        with open(path) as fin:
            src = fin.read()
        new, n = re.subn(pat, rep, src, flags=re.M)
        if n == 0:
            raise RuntimeError("patch did not apply: %s  %s" % (fn, pat))
        with open(path, "w") as fout:
            fout.write(new)
    with open(os.path.join(dest_pkg, "__init__.py"), "w") as fout:
        fout.write(INIT_PY)


def _build(dest: str) -> None:
    setup_py = os.path.join(dest, "_build_ext.py")
    with open(setup_py, "w") as fout:
        fout.write(textwrap.dedent("""
            import numpy as np
            from setuptools import setup, Extension
            from Cython.Build import cythonize
            mods = %r
            exts = [Extension('mope.' + m, ['mope/' + m + '.pyx'],
                              include_dirs=[np.get_include(), 'shims'],
                              extra_compile_args=['-O3', '-w'],
                              define_macros=[('NPY_NO_DEPRECATED_API', 'NPY_1_7_API_VERSION')])
                    for m in mods]
            setup(name='mope_ref', ext_modules=cythonize(exts, language_level=2, quiet=True),
                  script_args=['build_ext', '--inplace', '-q'])
        """ % (CYTHON_MODULES,)))
    subprocess.check_call([sys.executable, setup_py], cwd=dest)


def install(reference: str = "/root/reference", force: bool = False) -> str:
    src_pkg = os.path.join(reference, "mope")
    if not os.path.isdir(src_pkg):
        raise FileNotFoundError(src_pkg)
    stamp = os.path.join(DEST, ".installed")
    if os.path.exists(stamp) and not force and open(stamp).read().strip() == INSTALL_VERSION:
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(src_pkg, os.path.join(DEST, "mope"))
    for r, _d, files in os.walk(os.path.join(DEST, "mope")):
        for f in files:
            os.chmod(os.path.join(r, f), 0o644)
        os.chmod(r, 0o755)
    # the reference's own example inputs (read by the golden generator and the bench's reference arm)
    shutil.copytree(os.path.join(reference, "examples"), os.path.join(DEST, "examples"))
    for r, _d, files in os.walk(os.path.join(DEST, "examples")):
        for f in files:
            os.chmod(os.path.join(r, f), 0o644)
        os.chmod(r, 0o755)
    shutil.copytree(os.path.join(HERE, "ref_shims"), os.path.join(DEST, "shims"))
    _patch(os.path.join(DEST, "mope"))
    _build(DEST)
    with open(stamp, "w") as fout:
        fout.write(INSTALL_VERSION + "\n")
    return DEST


def activate() -> bool:
    """Put the installed reference on sys.path.  Returns False if it is not installed."""
    if not os.path.exists(os.path.join(DEST, ".installed")):
        return False
    for p in (os.path.join(DEST, "shims"), DEST):
        if p not in sys.path:
            sys.path.insert(0, p)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(install(a.reference, a.force))

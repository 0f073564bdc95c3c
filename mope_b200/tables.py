"""Transition tables: the in-memory state of the reference's TransitionData(memory=True) /
TransitionDataBottleneck(memory=True) (mope/transition_data_2d.py:78-111,
mope/transition_data_bottleneck.py:74-96) as dense sorted grids."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class TransitionTables:
    drift: np.ndarray            # [G, U, F, F]
    gens: np.ndarray             # [G] float64 ascending
    us: np.ndarray               # [U] float64 ascending
    N: float
    freqs: np.ndarray            # [F]
    breaks: np.ndarray           # [F] int
    bot: Optional[np.ndarray] = None      # [NB, UB, F, F]
    nbs: Optional[np.ndarray] = None      # [NB]
    bot_us: Optional[np.ndarray] = None   # [UB]

    def __post_init__(self):
        self.drift = np.ascontiguousarray(self.drift, dtype=np.float64)
        self.gens = np.ascontiguousarray(self.gens, dtype=np.float64)
        self.us = np.ascontiguousarray(self.us, dtype=np.float64)
        self.freqs = np.ascontiguousarray(self.freqs, dtype=np.float64)
        self.breaks = np.asarray(self.breaks)
        if self.breaks.ndim == 0:
            # tables written with --breaks unset store 0 here; the reference's root prior cannot use them
            # (np.add.reduceat raises, likelihoods.py:59)
            raise ValueError("un-binned transition tables (breaks=0) cannot be used for inference")
        self.N = float(self.N)
        G, U, F, F2 = self.drift.shape
        if F != F2 or self.gens.shape != (G,) or self.us.shape != (U,) or self.freqs.shape != (F,):
            raise ValueError("inconsistent drift table shapes")
        if np.any(np.diff(self.gens) <= 0) or np.any(np.diff(self.us) <= 0):
            raise ValueError("grid axes must be strictly ascending")
        if self.bot is not None:
            self.bot = np.ascontiguousarray(self.bot, dtype=np.float64)
            self.nbs = np.ascontiguousarray(self.nbs, dtype=np.float64)
            self.bot_us = np.ascontiguousarray(self.bot_us, dtype=np.float64)
            if self.bot.shape != (self.nbs.shape[0], self.bot_us.shape[0], F, F):
                raise ValueError("inconsistent bottleneck table shapes")
        # the clamp of _util.check_transition_matrix (mope/_util.pyx:140-143) is a no-op on blends of
        # matrices whose entries lie in [0,1]; the device path relies on that for age-scaled branches
        if self.drift.min() < 0.0 or self.drift.max() > 1.0:
            raise ValueError("drift tables hold entries outside [0, 1]")

    @property
    def F(self) -> int:
        return int(self.freqs.shape[0])

    # reference accessors (transition_data_2d.py:330-356)
    def get_bin_frequencies(self):
        return self.freqs

    def get_breaks(self):
        return self.breaks

    def get_N(self):
        return self.N

    def get_min_coal_time(self):
        return self.gens.min() / self.N

    def get_max_coal_time(self):
        return self.gens.max() / self.N

    def get_min_mutsel_rate(self):
        return self.us.min() * 2 * self.N

    def get_max_mutsel_rate(self):
        return self.us.max() * 2 * self.N

    def save_hdf5(self, transitions_file: str, bottleneck_file: Optional[str] = None, breaks_args=None, selection: bool = False) -> None:
        """Write master files in the reference's HDF5 layout (generate_transitions.py:242-323, :420-470): datasets `P<idx>`
        with attributes N, s, u, v, gen (drift) or N, Nb, u (bottleneck), root attributes N, breaks, frequencies (and
        min_bin_size / uniform_weight when `breaks_args = (uniform_weight, min_bin_size)` is given).  Written by the
        built-in HDF5 writer (hdf5_lite): the image has no libhdf5.  `selection`: `us` holds selection coefficients, written
        as attribute s with u = v = 0 (the recipe of `mope generate-commands --selection`, generate_transitions.py:549-551)."""
        from . import hdf5_lite
        root = {"N": int(self.N), "breaks": np.asarray(self.breaks, dtype=np.int64), "frequencies": self.freqs}
        if breaks_args is not None:
            root["uniform_weight"], root["min_bin_size"] = float(breaks_args[0]), float(breaks_args[1])
        dsets, idx = {}, 0
        for j, u in enumerate(self.us):
            for i, g in enumerate(self.gens):
                attrs = {"N": int(self.N), "s": float(u), "u": 0.0, "v": 0.0, "gen": int(g)} if selection else \
                    {"N": int(self.N), "s": 0.0, "u": float(u), "v": float(u), "gen": int(g)}
                dsets["P%d" % idx] = (self.drift[i, j], attrs)
                idx += 1
        hdf5_lite.write_file(transitions_file, root, dsets)
        if bottleneck_file is not None:
            if self.bot is None:
                raise ValueError("no bottleneck tables to write")
            dsets, idx = {}, 0
            for j, u in enumerate(self.bot_us):
                for i, nb in enumerate(self.nbs):
                    dsets["P%d" % idx] = (self.bot[i, j], {"N": int(self.N), "Nb": int(nb), "u": float(u)})
                    idx += 1
            hdf5_lite.write_file(bottleneck_file, root, dsets)

    def save(self, path: str) -> None:
        d = dict(drift=self.drift, gens=self.gens, us=self.us, N=np.float64(self.N), freqs=self.freqs, breaks=self.breaks)
        if self.bot is not None:
            d.update(bot=self.bot, nbs=self.nbs, bot_us=self.bot_us)
        np.savez(path, **d)


def load_npz(path: str) -> TransitionTables:
    z = np.load(path)
    kw = dict(drift=z["drift"], gens=z["gens"], us=z["us"], N=float(z["N"]), freqs=z["freqs"], breaks=z["breaks"])
    if "bot" in z.files:
        kw.update(bot=z["bot"], nbs=z["nbs"], bot_us=z["bot_us"])
    return TransitionTables(**kw)


def _read_h5(path: str):
    """-> (root attrs, {dataset name: (array, attrs)}) of a master file.  Real HDF5 files are parsed by h5py where it is
    installed and by the built-in reader (hdf5_lite) otherwise; anything else is handed to whatever `h5py` module the
    process has (the reference's test installs use a stand-in)."""
    from . import hdf5_lite
    with open(path, "rb") as f:
        head = f.read(8)
    is_hdf5 = head == hdf5_lite.SIGNATURE
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is None or (is_hdf5 and not hasattr(h5py, "h5")):   # no h5py, or a stand-in that cannot read real HDF5
        if not is_hdf5:
            raise ValueError("{}: not an HDF5 file".format(path))
        return hdf5_lite.read_file(path)
    with h5py.File(path, "r") as f:
        dsets = {name: (np.array(f[name][:, :], dtype=np.float64), dict(f[name].attrs)) for name in f}
        return dict(f.attrs), dsets


def _grid_from_h5(path: str, axis_attr: str, xkey: str = "u"):
    """Read a mope master HDF5 file (datasets with attrs N, gen|Nb, u; file attrs frequencies, breaks), with the checks
    of TransitionData._add_dataset / _add_master_file (transition_data_2d.py:43-111)."""
    root, dsets = _read_h5(path)
    mats, N = {}, None
    for name, (mat, a) in dsets.items():
        if N is None:
            N = a["N"]
        elif a["N"] != N:
            raise Exception("Found multiple population sizes in master file's datasets")
        key = (float(a[axis_attr]), float(a[xkey]))
        if key in mats:
            raise Exception("Found duplicate grid point in matrices: {}".format(key))
        mats[key] = np.array(mat, dtype=np.float64)
    freqs = np.array(root["frequencies"])
    breaks = np.array(root["breaks"])
    ax = np.array(sorted({k[0] for k in mats}))
    xs = np.array(sorted({k[1] for k in mats}))
    missing = [(g, x) for g in ax for x in xs if (g, x) not in mats]
    if missing:
        raise ValueError("incomplete transition probability grid")
    F = freqs.shape[0]
    dense = np.empty((ax.shape[0], xs.shape[0], F, F))
    for i, g in enumerate(ax):
        for j, x in enumerate(xs):
            dense[i, j] = mats[(g, x)]
    return dense, ax, xs, float(N), freqs, breaks


def load(transitions_file, bottleneck_file=None, selection=False) -> TransitionTables:
    """`transitions_file` / `bottleneck_file` as given to `mope run`: HDF5 master files, or one .npz
    holding both grids (written by TransitionTables.save), or an already-built TransitionTables.  `selection`: the
    second axis of the drift grid is the selection coefficient `s` of every dataset instead of its mutation rate `u`
    (TransitionData(selection=True), transition_data_2d.py:52-55); `us` then holds the s values."""
    if isinstance(transitions_file, TransitionTables):
        return transitions_file
    if str(transitions_file).endswith(".npz"):
        tb = load_npz(transitions_file)
        if bottleneck_file is not None and str(bottleneck_file).endswith(".npz") and \
                os.path.abspath(bottleneck_file) != os.path.abspath(transitions_file):
            zb = np.load(bottleneck_file)
            tb.bot, tb.nbs, tb.bot_us = zb["bot"], zb["nbs"], zb["bot_us"]
            tb.__post_init__()
        elif bottleneck_file is None:
            tb.bot = tb.nbs = tb.bot_us = None
        return tb
    drift, gens, us, N, freqs, breaks = _grid_from_h5(transitions_file, "gen", "s" if selection else "u")
    kw = dict(drift=drift, gens=gens, us=us, N=N, freqs=freqs, breaks=breaks)
    if bottleneck_file is not None:
        bot, nbs, bus, _, _, _ = _grid_from_h5(bottleneck_file, "Nb")
        kw.update(bot=bot, nbs=nbs, bot_us=bus)
    return TransitionTables(**kw)

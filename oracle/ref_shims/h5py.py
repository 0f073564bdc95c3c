"""Tiny stand-in for h5py backed by one pickle per file (test infrastructure).

Covers exactly what mope's table generator and loaders touch: File(fn, mode) as a
context manager, File.attrs, create_dataset(name, data=...), iteration over dataset
names, File[name] -> dataset with .attrs, .shape and [:, :] slicing.
"""
import pickle
import numpy as np


class _Dataset(object):
    def __init__(self, data):
        self._data = np.array(data)
        self.attrs = {}

    @property
    def shape(self):
        return self._data.shape

    def __getitem__(self, idx):
        return self._data[idx]


class File(object):
    def __init__(self, filename, mode="r"):
        self._filename = filename
        self._mode = mode
        self.attrs = {}
        self._datasets = {}
        if mode == "r":
            with open(filename, "rb") as fin:
                blob = pickle.load(fin)
            self.attrs = blob["attrs"]
            for name, (data, attrs) in blob["datasets"].items():
                ds = _Dataset(data)
                ds.attrs = attrs
                self._datasets[name] = ds

    def create_dataset(self, name, data=None):
        ds = _Dataset(data)
        self._datasets[name] = ds
        return ds

    def __iter__(self):
        return iter(list(self._datasets.keys()))

    def __getitem__(self, name):
        return self._datasets[name]

    def close(self):
        if self._mode != "r":
            blob = {
                "attrs": dict(self.attrs),
                "datasets": {k: (v._data, dict(v.attrs)) for k, v in self._datasets.items()},
            }
            with open(self._filename, "wb") as fout:
                pickle.dump(blob, fout, protocol=4)
            self._mode = "r"

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

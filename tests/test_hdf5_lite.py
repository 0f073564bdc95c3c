"""The built-in HDF5 subset (mope_b200/hdf5_lite.py): reader against a file written by an independent HDF5 implementation,
writer against the reader, and the reference's master-file layout through TransitionTables.save_hdf5 / tables.load."""
import os
import struct

import numpy as np
import pytest

from mope_b200 import hdf5_lite, tables
from tests import helpers as H


def _matlab_sample():
    import scipy.io
    return os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")


def test_reader_parses_an_independent_hdf5_file():
    """MATLAB v7.3 files are HDF5 (512-byte user block, superblock 0, symbol-table group, version-2 contiguous layout,
    attribute message): SciPy ships one in its test data."""
    path = _matlab_sample()
    if not os.path.exists(path):
        pytest.skip("SciPy's HDF5 sample file is not installed")
    root, dsets = hdf5_lite.read_file(path)
    assert list(dsets) == ["testdouble"]
    arr, attrs = dsets["testdouble"]
    np.testing.assert_allclose(arr.ravel(), np.arange(9) * np.pi / 4, rtol=1e-15)
    assert arr.shape == (9, 1) and attrs == {"MATLAB_class": "double"}


def test_writer_round_trip_and_structure(tmp_path):
    rng = np.random.RandomState(0)
    n = 4136 + 5   # more datasets than the default drift grid
    dsets = {"P%d" % i: (rng.rand(3, 3), {"N": 1000, "s": 0.0, "u": 1e-9 * (i + 1), "v": 1e-9 * (i + 1), "gen": i + 1}) for i in range(n)}
    root = {"N": 1000, "breaks": np.arange(11, dtype=np.int64), "frequencies": np.linspace(0, 1, 1001), "min_bin_size": 0.01,
            "uniform_weight": 0.5, "cmd": "mope generate-transitions"}
    path = str(tmp_path / "t.h5")
    hdf5_lite.write_file(path, root, dsets)
    buf = open(path, "rb").read()
    assert buf[:8] == hdf5_lite.SIGNATURE and buf[8] == 0
    assert struct.unpack_from("<Q", buf, 40)[0] == len(buf)          # end-of-file address
    r2, d2 = hdf5_lite.read_file(path)
    assert set(d2) == set(dsets)
    for k, (a, at) in dsets.items():
        assert np.array_equal(d2[k][0], a) and d2[k][0].dtype == np.float64
        assert d2[k][1] == at
    for k, v in root.items():
        assert np.array_equal(r2[k], v)
    assert r2["breaks"].dtype == np.int64 and isinstance(r2["cmd"], str)
    # the names inside the symbol-table node are ordered (the library binary-searches them)
    names = [nm for nm, _ in hdf5_lite._Reader(buf).group_entries(*struct.unpack_from("<QQ", buf, 80))]
    assert names == sorted(names, key=lambda s: s.encode()) and len(names) == n


def test_master_files_in_the_reference_layout(tmp_path):
    tb = H.load_tables()
    T = tables.TransitionTables(drift=tb["drift"], gens=tb["gens"], us=tb["us"], N=float(tb["N"]), freqs=tb["freqs"],
                                breaks=tb["breaks"], bot=tb["bot"], nbs=tb["nbs"], bot_us=tb["bot_us"])
    d, b = str(tmp_path / "drift.h5"), str(tmp_path / "bot.h5")
    T.save_hdf5(d, b, breaks_args=(0.5, 0.02))
    root, dsets = hdf5_lite.read_file(d)
    assert set(root) == {"N", "breaks", "frequencies", "uniform_weight", "min_bin_size"}
    assert set(dsets["P0"][1]) == {"N", "s", "u", "v", "gen"}           # generate_transitions.py:264-268
    assert set(hdf5_lite.read_file(b)[1]["P0"][1]) == {"N", "Nb", "u"}  # :290-292
    back = tables.load(d, b)
    for k in ("drift", "gens", "us", "freqs", "bot", "nbs", "bot_us"):
        assert np.array_equal(getattr(back, k), getattr(T, k)), k
    assert np.array_equal(back.breaks, T.breaks) and back.N == T.N


def test_selection_master_file_round_trip(tmp_path):
    """A selection table set written in the reference's layout (attribute s per dataset, u = v = 0) and read back with the
    `s` attribute as the second grid axis (TransitionData(selection=True), transition_data_2d.py:52-55)."""
    tb = H.load_tables("tables_sel.npz")
    T = tables.TransitionTables(drift=tb["drift"], gens=tb["gens"], us=tb["us"], N=float(tb["N"]), freqs=tb["freqs"], breaks=tb["breaks"])
    path = str(tmp_path / "selection_transitions.h5")
    T.save_hdf5(path, selection=True)
    _root, dsets = hdf5_lite.read_file(path)
    assert {a["u"] for _m, a in dsets.values()} == {0.0} and {a["s"] for _m, a in dsets.values()} == set(tb["us"].tolist())
    back = tables.load(path, selection=True)
    assert np.array_equal(back.drift, T.drift) and np.array_equal(back.us, T.us) and np.array_equal(back.gens, T.gens)
    # read as a neutral set every generation appears once per coefficient with the same u = 0: rejected like the reference does
    # (transition_data_2d.py:66-70, "Found duplicate generation & mutation/selection rate combination")
    with pytest.raises(Exception, match="duplicate"):
        tables.load(path)

"""The pure-Python HDF5 writer (inovesa_b200/h5native.py) against an independent reader written from the
format specification (tests/h5reader.py), the published test vectors of its checksum, and -- where h5py
is importable -- the real library."""
import numpy as np
import pytest

from inovesa_b200.h5native import lookup3, write_hdf5
from inovesa_b200.results import GROUPS, SOFT_LINKS, ResultFile
from h5reader import read_hdf5

f32 = np.float32


def test_checksum_test_vectors():
    # lookup3.c, driver5(): hashlittle of the empty string and of the usual sentence
    assert lookup3(b"") == 0xDEADBEEF
    assert lookup3(b"", 0xDEADBEEF) == 0xBD5B7DDE
    assert lookup3(b"Four score and seven years ago") == 0x17770551
    assert lookup3(b"Four score and seven years ago", 1) == 0xCD628161


def test_round_trip_of_every_supported_thing(tmp_path):
    rng = np.random.default_rng(3)
    data = {
        "/A/data": rng.random((5, 1, 16)).astype(f32),
        "/A/B/deep": rng.integers(-9, 9, (3, 4)).astype(np.int32),
        "/scalar": np.float64(2.5),
        "/Info/u": np.arange(7, dtype=np.uint32),
        "/Info/big": rng.integers(0, 2 ** 40, 5).astype(np.int64),
        "/Empty/data": np.zeros((0, 1, 16), f32),
        "/doubles": rng.random(9),
    }
    attrs = {"/A/data": {"Meter": 1.5e-3, "Second": np.float32(2.0)}, "/A": {"note": "hello", "n": np.int32(7)},
             "/Info": {"vec": np.arange(3, dtype=np.float64), "flag": True}, "/": {"creator": "inovesa_b200"}}
    links = {"/A/axis0": "/Info/u", "/Empty/axis0": "/Info/u"}
    path = str(tmp_path / "t.h5")
    write_hdf5(path, data, attrs, groups=["/OnlyAGroup", "/A/B"], soft_links=links)
    d, a, l = read_hdf5(path)
    assert set(d) == set(data)
    for k, v in data.items():
        v = np.asarray(v)
        assert d[k].dtype == v.dtype and d[k].shape == v.shape and np.array_equal(d[k], v), k
    assert l == links
    # the package's own reader (start distributions from an earlier result file, run.py -i)
    from inovesa_b200.h5native import read_dataset
    for k, v in data.items():
        assert np.array_equal(read_dataset(path, k), np.asarray(v)), k
    with pytest.raises(KeyError):
        read_dataset(path, "/A/missing")
    assert a["/A/data"]["Meter"] == 1.5e-3 and a["/A/data"]["Second"] == np.float32(2.0)
    assert a["/A"]["note"] == b"hello" and a["/A"]["n"] == 7
    assert np.array_equal(a["/Info"]["vec"], [0.0, 1.0, 2.0]) and a["/Info"]["flag"] == 1
    assert a["/"]["creator"] == b"inovesa_b200"
    try:
        import h5py
    except ImportError:
        return
    with h5py.File(path, "r") as f:  # the real library, where there is one
        for k, v in data.items():
            assert np.array_equal(f[k][()], v), k
        assert f["/A/data"].attrs["Meter"] == 1.5e-3
        assert f.get("/A/axis0", getlink=True).path == "/Info/u"
        assert "OnlyAGroup" in f


def test_result_file_in_the_references_layout(tmp_path):
    """ResultFile.to_hdf5 (src/IO/HDF5File.cpp:29-349): groups, datasets with [time][bunch][...] rows, unit
    attributes, axis soft links, /Info/Parameters attributes -- read back from the written file"""
    n, nmax = 32, 256
    ax = np.linspace(-6, 6, n).astype(f32)
    rf = ResultFile(n, nmax, 3, axis_q=ax, axis_p=ax, axis_f=np.arange(nmax, dtype=f32), bl=1e-3, dE=6e5, f_scale=3e11,
                    Ib=1e-3, Qb=3.7e-10, t_sync=1.2e-4, f_rev=2.7e6, volts=12.0, watt_per_hertz=2.0, watts=5.0,
                    impedance=np.ones((nmax, 2), f32))
    rf.add_parameter("GridSize", 32)
    rf.add_parameter("BunchCurrent", 1e-3)
    rng = np.random.default_rng(0)
    for k in range(4):
        rf.append_time_step(0.1 * k, rng.random(n), 1.0, 0.0, rng.random(n), 1.0, 0.0)
        rf.append_csr(rng.random(nmax), 3.0)
        rf.append_wake(rng.random(n))
        rf.append_tracks(rng.random((3, 2)) * (n - 1))
    rf.append_phase_space(0.3, rng.random((n, n)))
    path = str(tmp_path / "result.h5")
    rf.to_hdf5(path)
    d, a, l = read_hdf5(path)
    want = rf.datasets()
    assert set(d) == set(want)
    for k, v in want.items():
        assert np.array_equal(d[k], v) and d[k].dtype == v.dtype, k
    assert d["/BunchProfile/data"].shape == (4, 1, n) and d["/PhaseSpace/data"].shape == (1, 1, n, n)
    assert d["/CSR/Spectrum/data"].shape == (4, 1, nmax // 2) and d["/Particles/data"].shape == (4, 3, 2)
    assert l == SOFT_LINKS
    assert a["/WakePotential/data"]["Volt"] == 12.0 and a["/CSR/Intensity/data"]["Watt"] == 5.0
    assert a["/Info/AxisValues_t"]["Second"] == 1.2e-4 and a["/Impedance/data"]["Ohm"] == 1.0
    assert a["/Info/Parameters"]["GridSize"] == 32 and a["/Info/Parameters"]["BunchCurrent"] == 1e-3
    # the same through the dataset-less route the driver takes, and the file is a few KB of headers + the data
    assert rf.to_hdf5(path, native=True) < 200_000

"""The result file keeps the reference's HDF5 layout (src/IO/HDF5File.cpp:29-349): dataset paths,
row shapes, unit attributes, soft links -- checked without HDF5 (npz + layout substitute)."""
import json
import os

import numpy as np
import pytest

from inovesa_b200.results import DEFLATE, GROUPS, SOFT_LINKS, ResultFile

f32 = np.float32


def make(n=16, nmax=128, npart=3):
    ax = np.linspace(-6, 6, n, dtype=f32)
    return ResultFile(n, nmax, npart, axis_q=ax, axis_p=ax * 2, axis_f=np.arange(nmax, dtype=f32), bl=8.2e-4,
                      dE=6.1e5, f_scale=3.6e11, Ib=5e-5, Qb=1.8e-11, t_sync=1.2e-4, f_rev=2.7e6,
                      impedance=np.ones((nmax, 2), f32))


def test_paths_shapes_and_links_follow_the_reference():
    rf = make()
    n, h = 16, 64
    for k in range(3):
        rf.append_time_step(0.1 * k, np.ones(n), 1.0, 0.0, np.ones(n), 1.0, 0.0)
        rf.append_csr(np.arange(128), 2.5)
        rf.append_wake(np.zeros(n))
        rf.append_tracks([[1.7, 2.2], [15.9, 0.1], [40.0, -3.0]])
    rf.append_phase_space(0.2, np.zeros((n, n)))
    d = rf.datasets()
    want = {"/Info/AxisValues_z": (n,), "/Info/AxisValues_E": (n,), "/Info/AxisValues_f": (h,),
            "/Info/AxisValues_t": (3,), "/Info/BucketNumbers": (1,), "/BunchProfile/data": (3, 1, n),
            "/BunchLength/data": (3, 1), "/BunchPosition/data": (3, 1), "/BunchPopulation/data": (3, 1),
            "/EnergyProfile/data": (3, 1, n), "/EnergySpread/data": (3, 1), "/EnergyAverage/data": (3, 1),
            "/Particles/data": (3, 3, 2), "/WakePotential/data": (3, 1, n), "/CSR/Spectrum/data": (3, 1, h),
            "/CSR/Intensity/data": (3, 1), "/PhaseSpace/axis0": (1,), "/PhaseSpace/data": (1, 1, n, n),
            "/Impedance/data/real": (h,), "/Impedance/data/imag": (h,), "/RFKicks/data": (0, 2)}
    for p, shape in want.items():
        assert d[p].shape == shape, (p, d[p].shape)
    # quirks kept: E axis written from axis 0; spectrum rows hold the first nmax/2 bins;
    # particles stored as axis values at the truncated (and clamped) grid index
    assert np.array_equal(d["/Info/AxisValues_E"], d["/Info/AxisValues_z"])
    assert np.array_equal(d["/CSR/Spectrum/data"][0, 0], np.arange(64, dtype=f32))
    ax = np.linspace(-6, 6, n, dtype=f32)
    assert np.allclose(d["/Particles/data"][0], [[ax[1], 2 * ax[2]], [ax[15], 2 * ax[0]], [ax[15], 2 * ax[0]]])
    lay = rf.layout()
    assert lay["soft_links"]["/PhaseSpace/axis2"] == "/Info/AxisValues_E" and lay["deflate"] == DEFLATE == 6
    assert "/Info/Parameters" in GROUPS and len(SOFT_LINKS) == 19
    assert lay["datasets"]["/BunchLength/data"]["attrs"] == {"Meter": 8.2e-4, "Second": 8.2e-4 / 2.99792458e8}


def test_save_round_trip_and_hdf5_file(tmp_path):
    rf = make()
    rf.add_parameter("grid_size", 16)
    rf.append_time_step(0.0, np.ones(16), 1.0, 0.0, np.ones(16), 1.0, 0.0)
    base = str(tmp_path / "out")
    rf.save(base)
    z = np.load(base + ".npz")
    assert z["BunchProfile.data"].shape == (1, 1, 16)
    lay = json.load(open(base + ".layout.json"))
    assert lay["group_attrs"]["/Info/Parameters"]["grid_size"] == 16
    # through h5py where it exists, else through the pure-Python writer (tests/test_h5native.py reads it back)
    rf.to_hdf5(base + ".h5")
    assert open(base + ".h5", "rb").read(8) == b"\x89HDF\r\n\x1a\n"


def test_unit_factors_equal_the_reference_objects(ref, capi):
    """the Volt / WattPerHertz / Watt attributes (src/IO/HDF5File.cpp:259-276) against the values the
    reference's own ElectricField objects hold (oracle/_ref, src/PS/ElectricField.cpp:23,39-40)"""
    from inovesa_b200.params import kara_params
    from inovesa_b200.results import unit_factors
    for kw in (dict(grid_size=64, bunch_current=100e-6), dict(grid_size=128, steps_per_Ts=4000, bunch_current=2e-3, gap=0.0)):
        p = kara_params(**kw)
        rs = ref.sim(p)
        want = rs.units()
        rs.close()
        got = unit_factors(capi.derive(p))
        assert np.allclose(got, want, rtol=1e-6, atol=0), (got, want)


@pytest.mark.gpu
def test_run_and_record_fills_the_layout(capi, ctx_fma, rng):
    """main()'s loop at its out-step cadence through the simulation object: one row per out-step"""
    from inovesa_b200.params import kara_params
    from inovesa_b200.results import run_and_record
    p = kara_params(grid_size=64, bunch_current=100e-6)
    p.fp_track = 1
    parts = (rng.random((5, 2)) * 40 + 12).astype(f32)
    sim = capi.Simulation(ctx_fma, p, 1, particles=parts)
    d = capi.derive(p)
    rf = run_and_record(sim, p, d, steps=200, outstep=50, save_phase_space=2, wake_impedance=sim.tables["zwake"])
    rf.wake_impedance = sim.tables["zwake"]
    sim.close()
    ds = rf.datasets()
    assert ds["/Info/AxisValues_t"].shape == (5,) and np.allclose(ds["/Info/AxisValues_t"], np.arange(5) * 50 / d.steps)
    assert ds["/BunchProfile/data"].shape == (5, 1, 64) and ds["/Particles/data"].shape == (5, 5, 2)
    assert ds["/PhaseSpace/data"].shape == (3, 1, 64, 64)   # out-steps 0, 2, 4
    assert np.all(ds["/CSR/Intensity/data"] > 0) and np.all(np.abs(ds["/BunchPopulation/data"] - 1) < 1e-4)
    assert np.all(ds["/BunchLength/data"] > 0.9) and rf.parameters["grid_size"] == 64
    # the spectrum rows are the device's updateCSR spectrum (first nmax/2 bins), whose sum over ALL bins
    # times delta_f is the intensity; the wake row belongs to the profile of the same out-step
    assert np.all(ds["/CSR/Spectrum/data"].max(axis=2) > 0)
    delta_f = 1.0 / d.dq / (d.padded_bins - 1)
    assert np.allclose(ds["/CSR/Spectrum/data"].sum(axis=2) * delta_f, ds["/CSR/Intensity/data"], rtol=0.05)
    assert np.abs(ds["/WakePotential/data"][0]).max() > 0          # t = 0 row is not the zero-initialised buffer
    w_now = capi.Context.wake_potential(ctx_fma, ds["/BunchProfile/data"][-1, 0], rf.wake_impedance, d.wakescaling)
    assert np.allclose(ds["/WakePotential/data"][-1, 0], w_now, rtol=1e-4, atol=1e-6 * np.abs(w_now).max())

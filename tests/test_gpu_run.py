"""The main()-equivalent driver (inovesa_b200/run.py) end to end on the device: a config file in the
reference's format, the loop with its out-step cadence, the result file with the reference's layout,
the saved configuration -- and the same physics as driving the simulation object by hand."""
import io
import os
import sys

import numpy as np
import pytest

from inovesa_b200 import run
from inovesa_b200.params import kara_params

pytestmark = pytest.mark.gpu

CFG = """
GridSize = 256
StepsPerTs = 1000
rotations = 0.25
outstep = 50
BunchCurrent = 5e-5
RevolutionFrequency = 2.7157e6
HarmonicNumber = 184
BeamEnergy = 1.3e9
BeamEnergySpread = 4.7e-4
AcceleratingVoltage = 1.4e6
SynchrotronFrequency = 8.5e3
BendingRadius = 5.559
VacuumGap = -1
"""


def test_driver_runs_a_config_and_writes_the_result_layout(tmp_path, capi, ctx_fma):
    cfg = tmp_path / "kara.cfg"
    cfg.write_text(CFG)
    out = tmp_path / "result.h5"
    log = io.StringIO()
    rc = run.main(["-c", str(cfg), "-o", str(out), "--SavePhaseSpace", "2"], out=log)
    text = log.getvalue()
    assert rc == 0, text
    assert "Finished 250 steps" in text
    assert os.path.exists(str(out) + ".cfg") and "GridSize=256" in open(str(out) + ".cfg").read()
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from h5reader import read_hdf5  # the independent reader of tests/ (no HDF5 library in this image)
    data, attrs, links = read_hdf5(str(out))
    # 250 steps with an out-step every 50: rows at 0, 50, ..., 250 (src/main.cpp:961-1033 + the final save);
    # rows are [time][bunch][x] as in the reference's file (src/IO/HDF5File.cpp:60-75)
    assert data["/BunchProfile/data"].shape == (6, 1, 256)
    assert data["/CSR/Spectrum/data"].shape[0] == 6 and np.abs(data["/CSR/Spectrum/data"]).max() > 0
    assert np.allclose(data["/Info/AxisValues_t"], np.arange(6) * 0.05)
    assert data["/PhaseSpace/data"].shape[1:] == (1, 256, 256) and links["/BunchProfile/axis0"] == "/Info/AxisValues_t"
    assert attrs["/Info/Parameters"]["grid_size"] == 256 and attrs["/BunchProfile/data"]["AmperePerNBL"] > 0
    # the same physics as the simulation object driven by hand
    p = kara_params(grid_size=256, steps_per_Ts=1000, gap=-1.0, bunch_current=5e-5)
    sim = capi.Simulation(ctx_fma, p, 1)
    sim.renormalize()
    sim.step(250)
    prof = sim.download()[1][0]
    sim.close()
    assert np.array_equal(data["/BunchProfile/data"][-1, 0], prof)


def test_driver_resumes_from_a_result_file(tmp_path, capi, ctx_fma):
    """-i result.h5 (src/IO/HDF5File.cpp:425-480, src/main.cpp:497-512): the last saved phase space of one run is
    the start distribution of the next; 100 + 100 steps equal 200 steps in one go bit for bit (the initial
    renormalisation of a resumed run is switched off: it would rescale by the charge's round-off)"""
    cfg = tmp_path / "kara.cfg"
    cfg.write_text(CFG.replace("rotations = 0.25", "rotations = 0.1"))
    a, b, c = (str(tmp_path / n) for n in ("a.h5", "b.h5", "c.h5"))
    log = io.StringIO()
    assert run.main(["-c", str(cfg), "-o", a], out=log) == 0, log.getvalue()
    assert run.main(["-c", str(cfg), "-o", b, "-i", a, "--RenormalizeCharge", "-1", "--GridSize", "64"], out=log) == 0, log.getvalue()
    assert run.main(["-c", str(cfg), "-o", c, "-T", "0.2"], out=log) == 0, log.getvalue()
    assert 'Reading in initial distribution from: "%s"' % a in log.getvalue()
    from inovesa_b200.h5native import read_dataset
    end_b = read_dataset(b, "/PhaseSpace/data")[-1, 0]
    end_c = read_dataset(c, "/PhaseSpace/data")[-1, 0]
    assert end_b.shape == (256, 256)  # the file's grid size, not --GridSize
    assert np.array_equal(end_b, end_c)
    assert np.array_equal(read_dataset(b, "/BunchProfile/data")[-1, 0], read_dataset(c, "/BunchProfile/data")[-1, 0])


def test_driver_refuses_what_it_cannot_do(tmp_path):
    log = io.StringIO()
    assert run.main(["-c", "/dev/null", "-i", "start.png"], out=log) == 1
    assert "InitialDistFile" in log.getvalue()
    log = io.StringIO()
    assert run.main(["-c", str(tmp_path / "missing.cfg")], out=log) == 1
    assert "does not exist" in log.getvalue()


def test_driver_runs_a_train_of_bunches(tmp_path, capi, ctx_fma):
    """BunchCurrent with several entries is a filling pattern (src/main.cpp:248-279): the bunches are
    coupled through the wake potential of the padded train and the result file's rows are [time][bunch][x]"""
    cfg = tmp_path / "kara.cfg"
    cfg.write_text(CFG.replace("GridSize = 256", "GridSize = 128").replace("BunchCurrent = 5e-5", ""))
    out = tmp_path / "train.h5"
    log = io.StringIO()
    rc = run.main(["-c", str(cfg), "-o", str(out), "-I", "5e-5", "0", "1e-4", "--RenormalizeCharge", "40"], out=log)
    text = log.getvalue()
    assert rc == 0, text
    assert "2 bunches in 3 buckets" in text and "Finished 250 steps" in text
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from h5reader import read_hdf5
    data, attrs, links = read_hdf5(str(out))
    assert data["/BunchProfile/data"].shape == (6, 2, 128) and data["/WakePotential/data"].shape == (6, 2, 128)
    assert data["/PhaseSpace/data"].shape == (1, 2, 128, 128)
    assert list(data["/Info/BucketNumbers"]) == [2, 0]
    pop = data["/BunchPopulation/data"]
    assert np.allclose(pop[0], [1 / 3, 2 / 3], rtol=1e-5) and np.allclose(pop[-1], [1 / 3, 2 / 3], rtol=1e-4)
    assert abs(attrs["/BunchProfile/data"]["AmperePerNBL"] - 1.5e-4) < 1e-9
    # the same run by hand
    p = kara_params(grid_size=128, steps_per_Ts=1000, gap=-1.0)
    lay = capi.train_layout(p, [5e-5, 0, 1e-4])
    assert lay["spaced_bins"] == data["/Impedance/data/real"].shape[0] * 2
    assert np.abs(data["/WakePotential/data"][-1, 1]).max() > 0


def test_driver_resumes_a_train(tmp_path, capi, ctx_fma):
    """-i result.h5 with a filling pattern: the file's [bunch][x][y] row is the start distribution of the train;
    100 + 100 steps equal 200 in one go bit for bit"""
    cfg = tmp_path / "kara.cfg"
    cfg.write_text(CFG.replace("rotations = 0.25", "rotations = 0.1").replace("BunchCurrent = 5e-5", ""))
    a, b, c = (str(tmp_path / n) for n in ("a.h5", "b.h5", "c.h5"))
    fill = ["-I", "5e-5", "1e-4", "--RenormalizeCharge", "-1"]
    log = io.StringIO()
    assert run.main(["-c", str(cfg), "-o", a] + fill, out=log) == 0, log.getvalue()
    assert run.main(["-c", str(cfg), "-o", b, "-i", a] + fill, out=log) == 0, log.getvalue()
    assert run.main(["-c", str(cfg), "-o", c, "-T", "0.2"] + fill, out=log) == 0, log.getvalue()
    from inovesa_b200.h5native import read_dataset
    end_b = read_dataset(b, "/PhaseSpace/data")[-1]
    end_c = read_dataset(c, "/PhaseSpace/data")[-1]
    assert end_b.shape == (2, 256, 256) and np.array_equal(end_b, end_c)
    log = io.StringIO()  # one bunch in the file, two in the pattern
    single = str(tmp_path / "single.h5")
    assert run.main(["-c", str(cfg), "-o", single, "-I", "5e-5"], out=log) == 0
    assert run.main(["-c", str(cfg), "-i", single] + fill, out=log) == 0 and "2 in the filling pattern" in log.getvalue()

"""The driver's program options against the reference's (src/IO/ProgramOptions.cpp:22-392): names,
short forms, defaults, config-file format, precedence, compatibility aliases, refusals."""
import pytest

from inovesa_b200.options import OPTIONS, OptionError, ProgramOptions, parse_command_line, parse_config_text

CFG = """
# KARA, short-bunch mode
GridSize = 512          # mesh points
StepsPerTs=2000
BunchCurrent = 3e-4
VacuumGap = 0.032
RFVoltage = 1.4e6       # compatibility name
SyncFreq = 8.5e3
LinearRF = false
outstep = 50
"""


def test_defaults_are_the_references():
    o = ProgramOptions()
    # src/IO/ProgramOptions.cpp:29-267
    assert o["GridSize"] == 256 and o["StepsPerTs"] == 1000 and o["rotations"] == 5
    assert o["alpha0"] == 4e-3 and o["RevolutionFrequency"] == 9e6 and o["HarmonicNumber"] == 50
    assert o["BeamEnergy"] == 1.3e9 and o["BeamEnergySpread"] == 4.7e-4 and o["VacuumGap"] == 0.03
    assert o["AcceleratingVoltage"] == 1e6 and o["CutoffFreq"] == 23e9 and o["padding"] == 8.0
    assert o["FPType"] == 3 and o["FPTrack"] == 3 and o["derivation"] == 4 and o["InterpolationPoints"] == 4
    assert o["UseCSR"] is True and o["LinearRF"] is True and o["RoundPadding"] is True
    assert o["BunchCurrent"] == [1e-3] and o["outstep"] == 100 and o["RenormalizeCharge"] == 0
    p = o.ring_params()
    from inovesa_b200.params import default_params
    q = default_params()
    for name, _ in p._fields_:
        assert getattr(p, name) == getattr(q, name), name


def test_config_file_format_and_aliases():
    v = parse_config_text(CFG)
    assert v == {"GridSize": 512, "StepsPerTs": 2000, "BunchCurrent": [3e-4], "VacuumGap": 0.032,
                 "AcceleratingVoltage": 1.4e6, "SynchrotronFrequency": 8.5e3, "LinearRF": False, "outstep": 50}
    with pytest.raises(OptionError):
        parse_config_text("NoSuchOption = 1")
    with pytest.raises(OptionError):
        parse_config_text("config = other.cfg")  # command line only
    with pytest.raises(OptionError):
        parse_config_text("GridSize = many")


def test_command_line_forms():
    v = parse_command_line(["-s", "1024", "--StepsPerTs=4000", "-T2.5", "--BunchCurrent", "1e-4", "-v", "-o", "run.h5",
                            "--alpha1", "-2e-4"])
    assert v == {"GridSize": 1024, "StepsPerTs": 4000, "rotations": 2.5, "BunchCurrent": [1e-4], "verbose": True,
                 "output": "run.h5", "alpha1": -2e-4}
    assert parse_command_line(["--gui", "false", "--steps", "500"]) == {"gui": False, "StepsPerTs": 500}
    assert parse_command_line(["-I", "1e-3", "0", "2e-3"]) == {"BunchCurrent": [1e-3, 0.0, 2e-3]}
    for bad in (["--GridSize"], ["stray"], ["--unknown", "1"]):
        with pytest.raises(OptionError):
            parse_command_line(bad)


def test_command_line_wins_over_config_file_and_missing_files():
    files = {"ring.cfg": CFG}
    o = ProgramOptions().parse(["-c", "ring.cfg", "--GridSize", "2048", "-o", "/dev/null"], read_file=files.get)
    assert o["GridSize"] == 2048 and o["StepsPerTs"] == 2000 and o["output"] == ""
    p = o.ring_params()
    assert p.grid_size == 2048 and p.steps_per_Ts == 2000 and p.linear_rf == 0 and p.V_RF == 1.4e6
    assert abs(p.bunch_current - 3e-4) < 1e-10 and abs(p.f_s - 8.5e3) < 1e-3 and abs(p.gap - 0.032) < 1e-12  # bunch_current is a float
    ProgramOptions().parse([], read_file=files.get)  # no default.cfg: fine (src/IO/ProgramOptions.cpp:371)
    with pytest.raises(OptionError):
        ProgramOptions().parse(["-c", "nope.cfg"], read_file=files.get)
    text = o.save_text()
    assert "GridSize=2048" in text and "LinearRF=false" in text and "config" not in text
    # what was saved parses back to the same values
    again = ProgramOptions().parse(["-c", "saved.cfg"], read_file={"saved.cfg": text}.get)
    assert again.ring_params().grid_size == 2048 and again["StepsPerTs"] == 2000


def test_refusals():
    o = ProgramOptions().parse(["-I", "1e-3", "0", "2e-3"], read_file=lambda _: None)
    assert o.bunch_currents() == ([1e-3, 0.0, 2e-3], [1e-3, 2e-3])
    assert abs(o.ring_params().bunch_current - 3e-3) < 1e-9 and o.unsupported() == []  # a train: Ib = the sum
    o2 = ProgramOptions().parse(["-I", "1e-3", "2e-3", "--tracking", "t.txt", "--RenormalizeCharge", "5"], read_file=lambda _: None)
    assert o2.unsupported() == []  # passive particles and periodic renormalisation work with a train
    with pytest.raises(OptionError):
        ProgramOptions().parse(["-I", "0", "0"], read_file=lambda _: None).ring_params()
    o = ProgramOptions().parse(["-Z", "impedance.dat", "-i", "start.png"], read_file=lambda _: None)
    assert len(o.unsupported()) == 1 and o["Impedance"] == "impedance.dat"  # impedance files are read
    assert all(len(v) == 4 for v in OPTIONS.values())


def test_tracking_file_positions_become_grid_coordinates(tmp_path):
    """src/main.cpp:762-771: `q p` pairs in axis units -> PhaseSpace::x(q), y(p) (inc/PS/PhaseSpace.hpp:272-283)"""
    import numpy as np
    from types import SimpleNamespace
    from inovesa_b200.run import read_tracking_file
    f = tmp_path / "track.txt"
    f.write_text("0 0\n-6 6\n1.5 -2.25\n100 -100\nend\n")
    d = SimpleNamespace(qmin=-6.0, pmin=-6.0, dq=12.0 / 255, dp=12.0 / 255)
    p = read_tracking_file(f, d, 256)
    dq = np.float32(12.0 / 255)
    want = [[np.float32(6.0) / dq, np.float32(6.0) / dq], [0, 255], [np.float32(7.5) / dq, np.float32(3.75) / dq], [255, 0]]
    assert p.dtype == np.float32 and np.array_equal(p, np.asarray(want, np.float32))

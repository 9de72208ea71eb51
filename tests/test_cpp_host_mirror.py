"""The C++ host mirror (include/vfps_b200.hpp: the reference's SourceMap / PhaseSpace /
ElectricField classes above the C ABI) against the reference's own unit-test vectors: the test
program tests/cpp/test_host_mirror.cpp is compiled with g++ against libivs.so and run."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_tests.json")
ANGLES = {"half_pi": "1.57079632679489661923f", "pi": "3.14159265358979323846f"}


def _f(x):
    return "%rf" % float(x)  # repr of a float always carries a '.' or an exponent


def _arr(v):
    return "{" + ", ".join(_f(x) for x in v) + "}"


def write_golden_header(path):
    d = json.load(open(GOLDEN))
    out = ["// generated from tests/golden/reference_tests.json -- do not edit", "#pragma once", "namespace golden {"]
    out.append("struct kick_case { int axis; float offset[4]; float in[16]; float out[16]; float particles_out[9][2]; };")
    out.append("static const kick_case kickmap[] = {")
    for k in d["kickmap"]:
        assert k["it"] == 2
        po = "{" + ", ".join(_arr(p) for p in k["particles_out"]) + "}"
        out.append("  {%d, %s, %s, %s, %s}," % (k["axis"], _arr(k["offset"]), _arr(k["in"]), _arr(k["out"]), po))
    out.append("};")
    out.append("static const float kickmap_tolerance = %s;" % _f(d["kickmap_tolerance"]))
    out.append("struct rot_case { float axes[4]; float angle; int it, clamp, rotmapsize; float in[16]; float out[16];"
               " int has_particle; float particle_in[2]; float particle_out[2]; float particle_tolerance; };")
    out.append("static const rot_case rotationmap[] = {")
    for r in d["rotationmap"]:
        hp = 1 if "particle_in" in r else 0
        out.append("  {%s, %s, %d, %d, %d, %s, %s, %d, %s, %s, %rf}," % (
            _arr(r["axes"]), ANGLES[r["angle"]], r["it"], r["clamp"], r["rotmapsize"], _arr(r["in"]), _arr(r["out"]),
            hp, _arr(r.get("particle_in", [0, 0])), _arr(r.get("particle_out", [0, 0])), float(r.get("particle_tolerance", 0))))
    out.append("};")
    out.append("static const float rotationmap_tolerance = %s;" % _f(d["rotationmap_tolerance"]))
    out.append("struct rot_bad { int it, clamp, rotmapsize; };")
    out.append("static const rot_bad rotationmap_invalid[] = {" + ", ".join(
        "{%d, %d, %d}" % (b["it"], b["clamp"], b["rotmapsize"]) for b in d["rotationmap_invalid_argument"]) + "};")
    ef = d["electricfield"]
    out.append("static const unsigned ef_n = %d, ef_nmax = %d;" % (ef["n"], ef["nmax"]))
    out.append("static const float ef_axes[4] = %s;" % _arr(ef["axes"]))
    out.append("static const float ef_roundtrip_tolerance = %s, ef_causality_limit = %s, ef_change = %s;" % (
        _f(ef["roundtrip_tolerance"]), _f(ef["causality_limit"]), _f(ef["change_magnitude"])))
    out.append("static const float ef_frev = %s, ef_fmax = %s;" % (_f(ef["freespace"]["f_rev"]), _f(ef["freespace"]["f_max"])))
    out.append("static const float phasespace_tolerance = %s;" % _f(d["phasespace"]["tolerance"]))
    ru = d["ruler"]
    out.append("static const unsigned ruler_steps = %d; static const float ruler_min = %s, ruler_max = %s, ruler_zerobin = %s;" % (
        ru["steps"], _f(ru["min"]), _f(ru["max"]), _f(ru["zerobin"])))
    out.append("}  // namespace golden")
    open(path, "w").write("\n".join(out) + "\n")


@pytest.fixture(scope="module")
def exe(tmp_path_factory, capi):
    capi.load()  # builds libivs.so if it is missing
    build = tmp_path_factory.mktemp("cpp")
    write_golden_header(str(build / "golden_vectors.hpp"))
    out = str(build / "test_host_mirror")
    libdir = os.path.join(ROOT, "inovesa_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", str(build), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", out,
                           "-L", libdir, "-livs", "-Wl,-rpath," + libdir])
    return out


def test_mirror_compiles_and_host_side_checks_pass(exe):
    """no GPU needed: Ruler, std::invalid_argument cases, and the loud failure without a device"""
    r = subprocess.run([exe, "--host-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("0 failures")


@pytest.mark.gpu
def test_reference_unit_tests_through_the_mirror(exe):
    """TestKickMap / TestRotationMap / TestElectricField / TestPhaseSpace known answers and ten
    steps of main()'s loop body through the mirrored classes, on cuda:0"""
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout and r.stdout.strip().endswith("0 failures")

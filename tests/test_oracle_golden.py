"""The oracle (our C restatement) against every known-answer vector the reference's own unit
tests hold for the hot path (tests/golden/reference_tests.json, numbers transcribed from
/root/reference/test/*.cpp).  When oracle/_ref is present the reference's own sources compiled
in place are put through the same vectors, which pins the shim build too."""
import json
import os

import numpy as np
import pytest

f32 = np.float32
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tests.json")))
ANGLES = {"half_pi": f32(np.pi / 2), "pi": f32(np.pi)}


def impls(request):
    out = [request.getfixturevalue("oracle"), request.getfixturevalue("oracle_fma")]
    from oracle import pyoracle
    if pyoracle.available("ref", False):
        out.append(request.getfixturevalue("ref"))
    if pyoracle.available("ref", True):
        out.append(request.getfixturevalue("ref_fma"))
    return out


@pytest.mark.parametrize("case", G["kickmap"], ids=lambda c: c["cite"])
def test_kickmap_known_answers(case, request):
    grid = np.array(case["in"], f32).reshape(4, 4)
    want = np.array(case["out"], f32).reshape(4, 4)
    offset = np.array(case["offset"], f32)
    parts = np.array([[0.5 * i, 0.5 * i] for i in range(9)], f32)
    pwant = np.array(case["particles_out"], f32)
    for impl in impls(request):
        if hasattr(impl, "kick_track"):
            got = impl.kick_apply(grid, offset, case["it"], case["axis"])
            pgot = impl.kick_track(4, case["axis"], offset, parts)
        else:
            got, pgot = impl.kick_apply(grid, offset, case["it"], case["axis"], parts)
        assert np.abs(got - want).max() < G["kickmap_tolerance"]
        assert np.abs(pgot - pwant).max() < G["kickmap_tolerance"]


@pytest.mark.parametrize("case", G["rotationmap"], ids=lambda c: c["cite"])
def test_rotationmap_known_answers(case, request):
    grid = np.array(case["in"], f32).reshape(4, 4)
    want = np.array(case["out"], f32).reshape(4, 4)
    a = case["axes"]
    for impl in impls(request):
        got = impl.rotation_apply(grid, a[0], a[1], a[2], a[3], ANGLES[case["angle"]], case["it"],
                                  case["clamp"], case["rotmapsize"])
        assert np.abs(got - want).max() < G["rotationmap_tolerance"]
        if "particle_in" in case:
            p = np.array([case["particle_in"]], f32)
            if hasattr(impl, "rotation_track"):
                pg = impl.rotation_track(4, a[0], a[1], a[2], a[3], ANGLES[case["angle"]], p)
            else:
                _, pg = impl.rotation_apply(grid, a[0], a[1], a[2], a[3], ANGLES[case["angle"]],
                                            case["it"], case["clamp"], case["rotmapsize"], p)
            assert np.abs(pg[0] - np.array(case["particle_out"], f32)).max() < case["particle_tolerance"]


@pytest.mark.parametrize("case", G["rotationmap_invalid_argument"], ids=lambda c: c["cite"])
def test_rotationmap_invalid_argument(case, request):
    grid = np.zeros((4, 4), f32)
    for impl in impls(request):
        with pytest.raises(ValueError):
            impl.rotation_apply(grid, -1, 1, -1, 1, ANGLES["half_pi"], case["it"], case["clamp"],
                                case["rotmapsize"])


def test_ruler_zerobin(oracle):
    r = G["ruler"]
    _, _, zb = oracle.ruler(r["steps"], r["min"], r["max"])
    assert zb == r["zerobin"]


def test_phasespace_gaussian(request):
    """TestPhaseSpace.cpp:38-107: default Gaussian integrates to 1, mean 0, sigma 1"""
    tol = G["phasespace"]["tolerance"]
    for impl in impls(request):
        g = impl.initial_grid(32, -12, 12, -12, 12, 1.0)
        px, py, integral, m = impl.reduce(g, -12, 12, -12, 12)
        assert abs(integral - 1) < tol
        assert abs(m[0]) < tol and abs(m[3]) < tol          # means
        assert abs(m[2] - 1) < tol and abs(m[5] - 1) < tol  # rms


def test_const_impedance(oracle):
    import ctypes
    n = 16
    z = np.zeros((n, 2), f32)
    oracle.lib.orc_z_const(n, 3.0, -2.0, z.reshape(-1))
    assert np.all(z[: n // 2] == np.array([3.0, -2.0], f32))
    assert np.all(z[n // 2:] == 0)


def _fixture_profile(impl):
    E = G["electricfield"]
    a = E["axes"]
    g = impl.initial_grid(E["n"], a[0], a[1], a[2], a[3], 1.0)
    px, _, _, _ = impl.reduce(g, a[0], a[1], a[2], a[3])
    return g, px


def test_wake_roundtrip_unit_impedance(request):
    """TestElectricField.cpp:147-168: Z = 1 => wake potential = profile * scaling (1e-6):
    pins FFT normalisation, sign convention and index mapping"""
    E = G["electricfield"]
    n, nmax, a = E["n"], E["nmax"], E["axes"]
    z = np.zeros((nmax, 2), f32)
    z[: nmax // 2, 0] = 1  # ConstImpedance(1)
    for impl in impls(request):
        g, px = _fixture_profile(impl)
        if hasattr(impl, "sim") and hasattr(impl, "kick_track"):  # Oracle
            scaling = f32(1e-3 / nmax)
            wake = impl.wake_potential(px, z, scaling)
        else:
            wake, _, _, scaling = impl.wake_potential(g, a[0], a[1], a[2], a[3], z, 1e-3)
        # with the reference's conventions only bins < nmax/2 carry Z, i.e. half the spectrum
        # plus its mirror: the round trip returns profile * scaling * nmax up to the Nyquist bin
        assert np.abs(wake - px * scaling * nmax).max() < E["roundtrip_tolerance"]


def test_free_space_wake_is_causal(request):
    """TestElectricField.cpp:175-239: with the free-space CSR impedance a change in the second
    half of the profile leaves the wake potential of the first half unchanged (< 2e-8)"""
    E = G["electricfield"]
    n, nmax, a = E["n"], E["nmax"], E["axes"]
    for impl in (request.getfixturevalue("oracle"), request.getfixturevalue("oracle_fma")):
        oracle = request.getfixturevalue("oracle")
        z = np.zeros((nmax, 2), f32)
        oracle.lib.orc_z_free_space(nmax, E["freespace"]["f_rev"], E["freespace"]["f_max"], z.reshape(-1))
        g, px = _fixture_profile(oracle)
        mod = px.copy()
        c = n // 2
        delta = f32(E["change_magnitude"]) * mod[c]
        mod[c] -= delta; mod[c + 1] += delta; mod[c + 2] += delta; mod[c + 3] -= delta
        # the fixture's scaling: Ib*dt*c/bl/(dp*sd*E0)/nmax with (1e-3,1e-3,3,24/31,1e3,1e9)
        scaling = f32(f32(1e-3 * 1e-3 * 2.99792458e8 / 3.0 / (float(f32(24) / f32(31)) * 1e3 * 1e9)) / f32(nmax))
        w0 = impl.wake_potential(px, z, scaling)
        w1 = impl.wake_potential(mod, z, scaling)
        change = np.abs(w0 - w1)
        assert change.max() > E["causality_limit"]          # the change is visible somewhere
        assert change[:c].max() < E["causality_limit"]       # but not ahead of it




def test_train_wake_potential_golden(oracle):
    """ElectricField::wakePotential with PhaseSpace::nb = 3 (src/PS/ElectricField.cpp:186-263): outputs of the
    reference's own sources (tests/golden/make_train_wake.py), bit for bit"""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_wake.npz"))
    for k in range(3):
        wake, padded = oracle.train_wake_potential(g["profiles"], g["case%d_buckets" % k], int(g["case%d_spacing" % k]),
                                                   g["z"], np.float32(g["case%d_scaling" % k]), full=True)
        assert np.array_equal(wake, g["case%d_wake" % k]) and np.array_equal(padded, g["case%d_padded" % k])

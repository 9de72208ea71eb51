"""Projections / moments / normalisation, the shared-memory FFT wake potential and CSR power,
and particle tracking on the device against the oracle (C ABI, host buffers)."""
import json
import os

import numpy as np
import pytest

from conftest import rel_l1

pytestmark = pytest.mark.gpu
f32 = np.float32
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tests.json")))


@pytest.mark.parametrize("n", [16, 64, 257, 1024])
def test_projections_integral_moments(ctx_strict, oracle, rng, n):
    g = rng.random((n, n)).astype(f32)
    if n == 1024:
        g = oracle.initial_grid(256, -6, 6, -6, 6).repeat(4, 0).repeat(4, 1)
    px, py, integral, m = ctx_strict.reduce(g, -6, 6, -5, 7)
    opx, opy, ointegral, om = oracle.reduce(g, -6, 6, -5, 7)
    # same rounded products; the device adds them in a fixed tree, the reference left to right.
    # Judge both against the exact (float64) sums: the tree is at least as close as the
    # reference's own sequential round-off, and the two agree to that round-off.
    ws = oracle.simpson_weights(n, f32(12) / f32(n - 1)).astype(np.float64)
    ex_px = (g.astype(np.float64) * ws[None, :]).sum(1)
    ex_py = (g.astype(np.float64) * ws[:, None]).sum(0)
    seq_tol = max(2e-6, 6e-8 * n)
    assert np.abs(px / ex_px - 1).max() < 1e-6 and np.abs(py / ex_py - 1).max() < 1e-6
    assert np.abs(opx / ex_px - 1).max() < seq_tol
    assert np.allclose(px, opx, rtol=seq_tol, atol=0) and np.allclose(py, opy, rtol=seq_tol, atol=0)
    assert abs(integral / ointegral - 1) < seq_tol
    assert np.allclose(m[[1, 2, 4, 5]], om[[1, 2, 4, 5]], rtol=10 * seq_tol)
    assert np.allclose(m[[0, 3]], om[[0, 3]], rtol=1e-3, atol=2e-5)


def test_normalize(ctx_strict, oracle, rng):
    n = 128
    g = rng.random((n, n)).astype(f32)
    got = ctx_strict.normalize(g, -6, 6)
    _, _, integral, _ = oracle.reduce(g, -6, 6, -6, 6)
    want = g * (f32(1.0) / f32(integral))
    assert rel_l1(got, want) < 1e-6
    assert abs(ctx_strict.reduce(got, -6, 6, -6, 6)[2] - 1) < 1e-5


def _profile(oracle, n, lo, hi):
    g = oracle.initial_grid(n, lo, hi, lo, hi)
    return oracle.reduce(g, lo, hi, lo, hi)[0]


def test_wake_roundtrip_unit_impedance(ctx_strict, oracle):
    """test/TestElectricField.cpp:147-168 on the device FFT"""
    E = G["electricfield"]
    n, nmax = E["n"], E["nmax"]
    px = _profile(oracle, n, -12, 12)
    z = np.zeros((nmax, 2), f32)
    z[: nmax // 2, 0] = 1
    scaling = f32(1e-3 / nmax)
    wake = ctx_strict.wake_potential(px, z, scaling)
    assert np.abs(wake - px * scaling * nmax).max() < E["roundtrip_tolerance"]


def test_free_space_wake_is_causal(ctx_strict, oracle):
    """test/TestElectricField.cpp:175-239 on the device FFT"""
    E = G["electricfield"]
    n, nmax = E["n"], E["nmax"]
    z = np.zeros((nmax, 2), f32)
    oracle.lib.orc_z_free_space(nmax, E["freespace"]["f_rev"], E["freespace"]["f_max"], z.reshape(-1))
    px = _profile(oracle, n, -12, 12)
    mod = px.copy()
    c = n // 2
    d = f32(E["change_magnitude"]) * mod[c]
    mod[c] -= d; mod[c + 1] += d; mod[c + 2] += d; mod[c + 3] -= d
    scaling = f32(f32(1e-3 * 1e-3 * 2.99792458e8 / 3.0 / (float(f32(24) / f32(31)) * 1e3 * 1e9)) / f32(nmax))
    change = np.abs(ctx_strict.wake_potential(px, z, scaling) - ctx_strict.wake_potential(mod, z, scaling))
    assert change.max() > E["causality_limit"] and change[:c].max() < E["causality_limit"]


@pytest.mark.parametrize("n,nmax", [(32, 256), (256, 2048), (512, 4096), (2048, 16384), (300, 1024)])
def test_wake_potential_and_formfactor(ctx_strict, oracle, rng, n, nmax):
    px = (_profile(oracle, 64, -6, 6).repeat(n // 64 + 1)[:n] * (1 + 0.1 * rng.standard_normal(n))).astype(f32)
    z, _ = oracle.make_impedance(nmax, 3.7e12, 5.559, 2.7157e6, -1)
    z2, _ = oracle.make_impedance(nmax, 3.7e12, 5.559, 2.7157e6, 0.032, 0, 5.8e7, 0.0, 0.008)
    z = (z + z2).astype(f32)
    scaling = f32(3e-8)
    wake, ff = ctx_strict.wake_potential(px, z, scaling, want_formfactor=True)
    owake, off, _ = oracle.wake_potential(px, z, scaling, full=True)
    # fp32 FFT vs the oracle's double-precision FFT rounded to float
    assert np.abs(ff - off).max() <= 3e-6 * np.abs(off).max()
    assert np.abs(wake - owake).max() <= 3e-6 * np.abs(owake).max()


def test_wake_function_is_the_same_operator(ctx_strict, oracle, rng):
    """sum_x' rho[x'] w[x-x'] == wakePotential()[x]: the Toeplitz form used by the fused step"""
    n, nmax = 256, 2048
    px = (_profile(oracle, n, -6, 6) * (1 + 0.2 * rng.standard_normal(n))).astype(f32)
    z, _ = oracle.make_impedance(nmax, 3.7e12, 5.559, 2.7157e6, -1)
    scaling = f32(3e-8)
    w = ctx_strict.wake_function(z, scaling, n).astype(np.float64)
    want = oracle.wake_potential(px, z, scaling)
    got = np.array([np.dot(px.astype(np.float64), w[x - np.arange(n) + n - 1]) for x in range(n)])
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


@pytest.mark.parametrize("n,nmax", [(32, 256), (256, 2048)])
def test_csr_power(ctx_strict, oracle, n, nmax):
    px = _profile(oracle, n, -6, 6)
    z, _ = oracle.make_impedance(nmax, 3.7e12, 5.559, 2.7157e6, -1)
    dq, fs = f32(12) / f32(n - 1), f32(2.99792458e8 / 8.2e-4)
    for cutoff in (23e9, 0.0):
        spec, power = ctx_strict.csr_power(px, z, dq, fs, cutoff)
        ospec, opower = oracle.csr_power(px, z, dq, fs, cutoff)
        assert np.abs(spec - ospec).max() <= 1e-5 * ospec.max()
        assert abs(power / opower - 1) < 1e-5


def test_tracking_bit_exact(ctx_strict, capi, oracle, rng):
    n = 128
    g = rng.random((n, n)).astype(f32)
    off = (rng.standard_normal(n) * 2).astype(f32)
    parts = (rng.random((5000, 2)) * (n - 1)).astype(f32)
    parts[:10] = [[0, 0], [n - 1, n - 1], [0.5, n - 1], [n - 1, 0.5], [1, 1], [n - 2, n - 2.5], [0, 5], [5, 0], [n, n], [63.5, 64.5]]
    for axis in (0, 1):
        assert np.array_equal(ctx_strict.track_kick(parts, off, n, axis), oracle.kick_track(n, axis, off, parts))
    w5 = capi.fokker_planck_weights(n, -6.0, 6.0, 2.2e-3, 3, 4)
    for ft in (1, 2):
        got = ctx_strict.track_fokker_planck(parts, g, w5, ft)
        want = oracle.fp_track(g, -6.0, 6.0, 2.2e-3, 3, 4, ft, parts)
        ok = np.isfinite(want).all(axis=1) & (parts[:, 1] < n - 1) & (parts[:, 0] < n)
        assert np.array_equal(got[ok], want[ok])
    _, dq, zb = oracle.ruler(n, -6, 6)
    assert np.array_equal(ctx_strict.track_rotate(parts, 0.3, zb, zb), oracle.rotation_track(n, -6, 6, -6, 6, 0.3, parts))
    # stochastic mode: reproducible for a seed, right first two moments
    a = ctx_strict.track_fokker_planck(parts, g, w5, 3, e1=2e-3, dp=0.1, seed=7, step=3)
    b = ctx_strict.track_fokker_planck(parts, g, w5, 3, e1=2e-3, dp=0.1, seed=7, step=3)
    c = ctx_strict.track_fokker_planck(parts, g, w5, 3, e1=2e-3, dp=0.1, seed=8, step=3)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    kick = (parts[:, 1] * (1 - 2e-3) - a[:, 1]).astype(np.float64)
    sigma = np.sqrt(2 * 2e-3) / 0.1
    assert abs(kick.mean()) < 4 * sigma / np.sqrt(len(kick)) and abs(kick.std() / sigma - 1) < 0.05

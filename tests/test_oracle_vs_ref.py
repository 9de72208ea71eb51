"""The C restatement (oracle/vfps_oracle.c) against the reference's own sources compiled in
place (oracle/_ref): bit-for-bit in the strict flavour for every function of the path and for
the whole simulation; the FMA flavour is pinned for the map loops.  Skipped where
/root/reference (hence oracle/_ref) is not available -- the committed fixtures of
test_oracle_fixtures.py carry the same comparison to such machines."""
import numpy as np
import pytest

from inovesa_b200.params import kara_params

f32 = np.float32


@pytest.mark.parametrize("n", [16, 64, 257])
def test_maps_bit_exact(oracle, ref, rng, n):
    g = rng.random((n, n)).astype(f32)
    for it in (1, 2, 3, 4):
        for axis in (0, 1):
            off = (rng.standard_normal(n) * 3).astype(f32)
            assert np.array_equal(ref.kick_apply(g, off, it, axis), oracle.kick_apply(g, off, it, axis))
    for der in (3, 4):
        for fpt in (0, 1, 2, 3):
            assert np.array_equal(ref.fp_weights5(n, -6, 6, 2.2e-3, fpt, der),
                                  oracle.fp_weights5(n, -6, 6, 2.2e-3, fpt, der))
            assert np.array_equal(ref.fp_apply(g, -6, 6, 2.2e-3, fpt, der),
                                  oracle.fp_apply(g, -6, 6, 2.2e-3, fpt, der))
    for it in (1, 2, 3, 4):
        for ang in (0.3, np.pi / 2, 2 * np.pi / 1000):
            want = ref.rotation_apply(g, -6, 6, -6, 6, ang, it)
            assert np.array_equal(want, oracle.rotation_apply(g, -6, 6, -6, 6, ang, it))
            assert np.array_equal(want, ref.rotation_apply(g, -6, 6, -6, 6, ang, it, 0, n * n))
    assert np.array_equal(ref.rotation_apply(g, -6, 6, -6, 6, 0.3, 4, 1, n * n),
                          oracle.rotation_apply(g, -6, 6, -6, 6, 0.3, 4, 1, n * n))


def test_maps_bit_exact_fma(oracle_fma, ref_fma, rng):
    """the flavour GCC produces from the reference's own flags on FMA hardware"""
    n = 64
    g = rng.random((n, n)).astype(f32)
    for it in (2, 3, 4):
        for axis in (0, 1):
            off = (rng.standard_normal(n) * 3).astype(f32)
            assert np.array_equal(ref_fma.kick_apply(g, off, it, axis), oracle_fma.kick_apply(g, off, it, axis))
    assert np.array_equal(ref_fma.fp_apply(g, -6, 6, 2.2e-3), oracle_fma.fp_apply(g, -6, 6, 2.2e-3))
    assert np.array_equal(ref_fma.rotation_apply(g, -6, 6, -6, 6, 0.3, 4),
                          oracle_fma.rotation_apply(g, -6, 6, -6, 6, 0.3, 4))


@pytest.mark.parametrize("n", [16, 64, 257])
def test_reductions_and_start_distribution_bit_exact(oracle, ref, rng, n):
    g = rng.random((n, n)).astype(f32)
    ra, oa = ref.reduce(g, -6, 6, -6, 6), oracle.reduce(g, -6, 6, -6, 6)
    for x, y in zip(ra, oa):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    assert np.array_equal(ref.initial_grid(n, -6, 6, -6, 6, 1.0), oracle.initial_grid(n, -6, 6, -6, 6, 1.0))


def test_tracking_bit_exact(oracle, ref, rng):
    n = 64
    g = rng.random((n, n)).astype(f32)
    off = (rng.standard_normal(n) * 2).astype(f32)
    parts = (rng.random((1000, 2)) * (n - 1)).astype(f32)
    for axis in (0, 1):
        _, pr = ref.kick_apply(g, off, 4, axis, parts)
        assert np.array_equal(pr, oracle.kick_track(n, axis, off, parts))
    for ft in (1, 2):
        assert np.array_equal(ref.fp_track(g, -6, 6, 2.2e-3, 3, 4, ft, parts),
                              oracle.fp_track(g, -6, 6, 2.2e-3, 3, 4, ft, parts), equal_nan=True)
    _, pr = ref.rotation_apply(g, -6, 6, -6, 6, 0.3, 4, 0, 0, parts)
    assert np.array_equal(pr, oracle.rotation_track(n, -6, 6, -6, 6, 0.3, parts))


def test_field_and_impedances_bit_exact(oracle, ref):
    n, nmax = 32, 256
    g = ref.initial_grid(n, -12, 12, -12, 12)
    px = ref.reduce(g, -12, 12, -12, 12)[0]
    cases = [(256, 1e12, 47.7, 1e6, -1, 1, 0, 0, 0), (128, 3.7e12, 5.559, 2.7157e6, 0.032, 1, 0, 0, 0),
             (128, 3.7e12, 5.559, 2.7157e6, 0.032, 1, 5.8e7, 0, 0.008),
             (256, 1e12, 5.559, 2.7e6, 0.032, 0, 5.8e7, 0.1, 0.004), (64, 1e12, 5.559, 2.7e6, 0, 1, 1, 0, 0)]
    for c in cases:
        zr, nzr = ref.make_impedance(*c)
        zo, nzo = oracle.make_impedance(*c)
        assert nzr == nzo
        assert np.array_equal(zr, zo), c
    z, _ = ref.make_impedance(nmax, 1e12, 47.7, 1e6, -1)
    wake, ff, pw, used = ref.wake_potential(g, -12, 12, -12, 12, z, 1e-3)
    ow, off, opw = oracle.wake_potential(px, z, used, full=True)
    assert np.array_equal(ow, wake) and np.array_equal(off, ff) and np.array_equal(opw, pw)
    sp, power = ref.csr_power(g, -12, 12, -12, 12, 1e-3, z, 23e9)
    so, po = oracle.csr_power(px, z, f32(24) / f32(31), f32(2.99792458e8 / float(f32(1e-3))), 23e9)
    assert np.array_equal(sp, so) and power == po


def test_train_wake_potential_bit_exact(oracle, ref, rng):
    """PhaseSpace::nb > 1: the bunches' profiles share one padded array at bucket * spacing_bins and one
    transform pair (src/PS/ElectricField.cpp:186-263); three bunches of different charge and shape"""
    n, nmax = 32, 256
    g0 = ref.initial_grid(n, -12, 12, -12, 12)
    grids = np.stack([g0 * f32(0.5), np.roll(g0, 3, axis=0) * f32(0.3), np.roll(g0, -2, axis=0) * f32(0.2)])
    px = np.stack([ref.reduce(g, -12, 12, -12, 12)[0] for g in grids])
    z, _ = ref.make_impedance(nmax, 1e12, 47.7, 1e6, -1)
    for buckets, spacing in (([4, 2, 0], 40), ([2, 1, 0], 32), ([0, 1, 5], 44)):
        wake, pw, used = ref.train_wake_potential(grids, -12, 12, -12, 12, buckets, spacing, z, 1e-3)
        ow, opw = oracle.train_wake_potential(px, buckets, spacing, z, used, full=True)
        assert np.array_equal(opw, pw) and np.array_equal(ow, wake)
        assert np.abs(wake).max() > 0
    # one bunch in bucket 0 is the single-bunch field
    w1 = oracle.train_wake_potential(px[:1], [0], 40, z, f32(1e-3))
    assert np.array_equal(w1[0], oracle.wake_potential(px[0], z, f32(1e-3)))
    with pytest.raises(ValueError):
        oracle.train_wake_potential(px, [6, 2, 0], 40, z, f32(1e-3))


def test_airy_against_scipy(oracle, ref):
    from scipy.special import airy
    for x in list(np.logspace(-3, 2, 40)) + [103.0]:
        want = airy(x)
        for impl in (oracle, ref):
            got, over = impl.airy(float(x))
            assert not over
            assert max(abs(got[i] / want[i] - 1) for i in range(4)) < 1e-11
    assert oracle.airy(106.0)[1] and ref.airy(106.0)[1]  # Bi overflows a double here


@pytest.mark.parametrize("kw", [
    dict(grid_size=64, bunch_current=50e-6),
    dict(grid_size=96, bunch_current=3e-4, gap=0.032, linear_rf=0, alpha1=1e-4, alpha2=1e-3,
         wall_conductivity=5.8e7, collimator=0.004, steps_per_Ts=4000),
    dict(grid_size=64, interpolation=3, derivation=3, gap=0.0),
], ids=["free-space", "pp+rw+coll+sinRF+alpha12", "quadratic-no-wake"])
def test_whole_simulation_bit_exact(oracle, ref, rng, kw):
    p = kara_params(**kw)
    p.fp_track = 2
    parts = (rng.random((64, 2)) * (p.grid_size - 1)).astype(f32)
    rs, os_ = ref.sim(p, None, parts), oracle.sim(p, None, parts)
    try:
        assert rs.derived.as_dict() == os_.derived.as_dict()
        tr, to = rs.tables(), os_.tables()
        for k in tr:
            assert np.array_equal(tr[k], to[k]), k
        for nst in (1, 9, 40):
            rs.step(nst)
            os_.step(nst)
            for a, b in zip(rs.get(), os_.get()):
                assert np.array_equal(a, b)
        (oa, sa, ya), (ob, sb, yb) = rs.observe(), os_.observe()
        assert np.array_equal(oa, ob) and np.array_equal(sa, sb) and np.array_equal(ya, yb)
    finally:
        rs.close()
        os_.close()


def test_dynamic_rf_kick_map_bit_for_bit(oracle, ref):
    """the oracle's DynamicRFKickMap restatement against the reference's own class (sine RF; the queue of
    (phase, amplitude) pairs is given to both: the reference fills it from an unseeded mt19937)"""
    from inovesa_b200.params import kara_params
    p = kara_params(grid_size=64, linear_rf=0, steps_per_Ts=2000)
    d = oracle.derive(p)
    rng = np.random.default_rng(3)
    sync = np.float32(np.arcsin(np.float32(d.V0) / np.float32(d.V_eff)))
    seq = oracle.rf_modulation(30, 1e-3, 2e-4, 5e-3, 0.01, d.revolutionpart, sync, rng.standard_normal((30, 2)))
    a, b, c = oracle.sim(p), ref.sim(p), oracle.sim(p)
    a.set_rf_modulation(seq)
    b.set_rf_modulation(seq)
    for sim in (a, b, c):
        sim.step(30)
    ga, gb, gc = a.get()[0], b.get()[0], c.get()[0]
    for sim in (a, b, c):
        sim.close()
    assert np.array_equal(ga, gb)
    assert not np.array_equal(ga, gc)

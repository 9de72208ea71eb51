"""Parity of the per-map CUDA kernels with the oracle, through the C ABI with host buffers.
STRICT arithmetic: bit-exact.  FMA arithmetic: bit-exact against the oracle's FMA flavour where
that flavour is a plain fused chain (kick maps), otherwise within a few ulp."""
import numpy as np
import pytest

from conftest import rel_l1

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.mark.parametrize("n", [4, 16, 64, 257, 512])
@pytest.mark.parametrize("it", [1, 2, 3, 4])
def test_kick_maps_strict_bit_exact(ctx_strict, oracle, rng, n, it):
    g = rng.random((n, n)).astype(f32)
    for scale in (0.3, 3.0, float(n)):  # sub-cell, multi-cell, and beyond the grid (overflow rows)
        off = (rng.standard_normal(n) * scale).astype(f32)
        assert np.array_equal(ctx_strict.kick_y(g, off, it), oracle.kick_apply(g, off, it, 1))
        assert np.array_equal(ctx_strict.kick_x(g, off, it), oracle.kick_apply(g, off, it, 0))


@pytest.mark.parametrize("it", [2, 3, 4])
def test_kick_maps_fma_bit_exact(ctx_fma, oracle_fma, rng, it):
    n = 256
    g = rng.random((n, n)).astype(f32)
    off = (rng.standard_normal(n) * 1.5).astype(f32)
    # the oracle's FMA flavour fuses the quadratic coefficient 1-f*f as GCC does; the device
    # computes coefficients unfused in both modes, so compare cubic/linear exactly and
    # quadratic to 1 ulp
    got_y, want_y = ctx_fma.kick_y(g, off, it), oracle_fma.kick_apply(g, off, it, 1)
    got_x, want_x = ctx_fma.kick_x(g, off, it), oracle_fma.kick_apply(g, off, it, 0)
    if it == 3:
        assert np.abs(got_y - want_y).max() <= 2.4e-7 and np.abs(got_x - want_x).max() <= 2.4e-7
    else:
        assert np.array_equal(got_y, want_y) and np.array_equal(got_x, want_x)


def test_reference_known_answers_on_device(ctx_strict):
    """test/TestKickMap.cpp:18-172 through the CUDA path"""
    import json, os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tests.json")))
    parts = np.array([[0.5 * i, 0.5 * i] for i in range(9)], f32)
    for case in G["kickmap"]:
        grid = np.array(case["in"], f32).reshape(4, 4)
        off = np.array(case["offset"], f32)
        fn = ctx_strict.kick_x if case["axis"] == 0 else ctx_strict.kick_y
        assert np.abs(fn(grid, off, case["it"]) - np.array(case["out"], f32).reshape(4, 4)).max() < 1e-5
        got = ctx_strict.track_kick(parts, off, 4, case["axis"])
        assert np.abs(got - np.array(case["particles_out"], f32)).max() < 1e-5
    ang = {"half_pi": f32(np.pi / 2), "pi": f32(np.pi)}
    for case in G["rotationmap"]:
        grid = np.array(case["in"], f32).reshape(4, 4)
        a = case["axes"]
        got = ctx_strict.rotate(grid, ang[case["angle"]], a[0], a[1], a[2], a[3], case["it"], case["clamp"])
        assert np.abs(got - np.array(case["out"], f32).reshape(4, 4)).max() < 1e-5
        if "particle_in" in case:
            p = ctx_strict.track_rotate(np.array([case["particle_in"]], f32), ang[case["angle"]], 1.5, 1.5)
            assert np.abs(p[0] - np.array(case["particle_out"], f32)).max() < 1e-3


def test_rotation_clamp_needs_cubic(ctx_strict, capi):
    """src/SM/RotationMap.cpp:83-87 -> std::invalid_argument -> IVS_ERR_INVALID_ARGUMENT"""
    g = np.zeros((4, 4), f32)
    with pytest.raises(capi.IvsError) as e:
        ctx_strict.rotate(g, 0.1, -1, 1, -1, 1, 3, 1)
    assert e.value.code == capi.ERR_INVALID_ARGUMENT
    with pytest.raises(capi.IvsError):
        ctx_strict.rotate(g, 0.1, 1, -1, -1, 1, 4, 0)  # Ruler min >= max


@pytest.mark.parametrize("n", [16, 255, 256])
@pytest.mark.parametrize("derivation", [3, 4])
def test_fokker_planck_strict_bit_exact(ctx_strict, capi, oracle, rng, n, derivation):
    g = rng.random((n, n)).astype(f32)
    for fptype in (0, 1, 2, 3):
        w5 = capi.fokker_planck_weights(n, -6.0, 6.0, 2.23e-3, fptype, derivation)
        assert np.array_equal(ctx_strict.fokker_planck(g, w5), oracle.fp_apply(g, -6.0, 6.0, 2.23e-3, fptype, derivation))


@pytest.mark.parametrize("n", [16, 64, 257])
@pytest.mark.parametrize("it", [1, 2, 3, 4])
def test_rotation_strict_bit_exact(ctx_strict, oracle, rng, n, it):
    g = rng.random((n, n)).astype(f32)
    for ang in (0.3, np.pi / 2, 2 * np.pi / 1000, -1.1):
        assert np.array_equal(ctx_strict.rotate(g, ang, -6, 6, -5, 7, it, 0),
                              oracle.rotation_apply(g, -6, 6, -5, 7, ang, it))
    if it == 4:
        assert np.array_equal(ctx_strict.rotate(g, 0.3, -6, 6, -6, 6, 4, 1),
                              oracle.rotation_apply(g, -6, 6, -6, 6, 0.3, 4, 1, n * n))


def test_fma_mode_is_within_rounding_of_strict(ctx_fma, ctx_strict, capi, rng):
    n = 256
    g = rng.random((n, n)).astype(f32)
    w5 = capi.fokker_planck_weights(n, -6.0, 6.0, 2.23e-3, 3, 4)
    assert rel_l1(ctx_fma.fokker_planck(g, w5), ctx_strict.fokker_planck(g, w5)) < 1e-7
    assert rel_l1(ctx_fma.rotate(g, 0.3, -6, 6, -6, 6, 4, 0), ctx_strict.rotate(g, 0.3, -6, 6, -6, 6, 4, 0)) < 1e-7


def test_generic_source_map(ctx_strict, rng):
    """SourceMap::apply (src/SM/SourceMap.cpp:98-107) with a random table"""
    n, ip = 32, 9
    g = rng.random((n, n)).astype(f32)
    index = rng.integers(0, n * n, (n * n, ip)).astype(np.uint32)
    weight = rng.standard_normal((n * n, ip)).astype(f32)
    want = np.zeros(n * n, f32)
    flat = g.reshape(-1)
    for j in range(ip):
        want = (want + flat[index[:, j]] * weight[:, j]).astype(f32)
    assert np.array_equal(ctx_strict.apply_source_map(g, index, weight).reshape(-1), want)


def test_map_properties_at_full_size(ctx_strict, capi, rng):
    """size-independent properties at the largest benchmark grid (2048^2), where the oracle is
    too slow to be the checker"""
    n = 2048
    g = rng.random((n, n)).astype(f32)
    zero = np.zeros(n, f32)
    # zero offset: cubic weights are exactly (0,1,0,0) -> identity
    assert np.array_equal(ctx_strict.kick_y(g, zero, 4), g)
    assert np.array_equal(ctx_strict.kick_x(g, zero, 4), g)
    # integer offsets are exact shifts with zero fill
    off = np.full(n, 3.0, f32)
    want = np.zeros_like(g)
    want[:, :-3] = g[:, 3:]
    assert np.array_equal(ctx_strict.kick_y(g, off, 4), want)
    want = np.zeros_like(g)
    want[:-3, :] = g[3:, :]
    assert np.array_equal(ctx_strict.kick_x(g, off, 4), want)
    # linearity: K(a) + K(b) == K(a+b) to rounding
    h = rng.random((n, n)).astype(f32)
    off = (rng.standard_normal(n) * 0.7).astype(f32)
    lhs = ctx_strict.kick_y(g, off, 4).astype(np.float64) + ctx_strict.kick_y(h, off, 4)
    assert rel_l1(lhs, ctx_strict.kick_y((g + h).astype(f32), off, 4)) < 1e-6
    # Fokker-Planck of a constant row: damping+diffusion weights sum to 1 + e1 in the bulk
    w5 = capi.fokker_planck_weights(n, -6.0, 6.0, 1e-3, 3, 4)
    out = ctx_strict.fokker_planck(np.ones((n, n), f32), w5)
    assert np.all(out[:, [0, 1, n - 2, n - 1]] == 0)
    assert np.allclose(out[:, 2:n - 2], 1.001, rtol=1e-5)

"""The product's host-side set-up helpers (include/ivs_host.h, csrc/ivs_host.cpp) against the
oracle: derived scalars, kick offsets, Fokker-Planck stencil, Simpson weights, impedance tables
and the start distribution must be BIT-identical, because the device consumes them as they
are.  No GPU involved (these entry points are pure host arithmetic)."""
import numpy as np
import pytest

from inovesa_b200.params import default_params, kara_params

f32 = np.float32


def close_complex(z, zo, rtol=3e-7):
    """|z - zo| <= rtol*|zo| per frequency bin (real and imaginary part can differ hugely in
    size, the error of the double sum scales with the modulus)"""
    mod = np.hypot(zo[:, 0].astype(np.float64), zo[:, 1])
    return bool(np.all(np.abs(z.astype(np.float64) - zo).max(axis=1) <= rtol * mod))

CONFIGS = {
    "reference-defaults": default_params(),
    "config1-free-space-256": kara_params(grid_size=256),
    "config2-parallel-plates-512": kara_params(grid_size=128, gap=0.032, bunch_current=1e-3),
    "config3-csr-rw-collimator": kara_params(grid_size=128, gap=0.032, wall_conductivity=5.8e7,
                                             collimator=0.008, steps_per_Ts=4000),
    "config5-sinRF-alpha12": kara_params(grid_size=192, linear_rf=0, alpha1=2e-4, alpha2=3e-3,
                                         steps_per_Ts=4000),
    "shifted-odd-grid": kara_params(grid_size=129, shift_x=1.5, shift_y=-2.0, interpolation=3,
                                    derivation=3, fp_type=1),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_derived_scalars_and_tables(capi, oracle, name):
    p = CONFIGS[name]
    d, do = capi.derive(p), oracle.derive(p)
    assert d.as_dict() == do.as_dict()
    n = p.grid_size
    sim = oracle.sim(p)
    try:
        t = sim.tables()
        assert np.array_equal(capi.rf_kick_offsets(p, d), t["rf"])
        assert np.array_equal(capi.drift_offsets(p, d), t["drift"])
        assert np.array_equal(capi.fokker_planck_weights(n, d.pmin, d.pmax, d.e1, p.fp_type, p.derivation), t["fp5"])
        zw, nz = capi.wake_impedance(p, d)
        zr = capi.radiation_impedance(p, d)
        if p.gap > 0:
            # parallel plates: three independent Airy implementations (product: series +
            # asymptotic, oracle: quadrature, shim: libstdc++ Bessel) agree to ~1e-10, so the
            # double sums round to the same float except for an occasional last bit
            assert close_complex(zw, t["zwake"]) and close_complex(zr, t["zrad"])
            assert (zw != t["zwake"]).mean() < 0.02
        else:
            assert np.array_equal(zw, t["zwake"])
            assert np.array_equal(zr, t["zrad"])
        assert np.array_equal(capi.initial_grid(d, n, p.zoom), sim.get()[0])
    finally:
        sim.close()
    assert np.array_equal(capi.simpson_weights(n, d.dq), oracle.simpson_weights(n, d.dq))


@pytest.mark.parametrize("fptype", [0, 1, 2, 3])
@pytest.mark.parametrize("derivation", [3, 4])
@pytest.mark.parametrize("n", [16, 255, 256])
def test_fokker_planck_weights(capi, oracle, fptype, derivation, n):
    w = capi.fokker_planck_weights(n, -6.0, 6.0, 2.23e-5, fptype, derivation)
    assert np.array_equal(w, oracle.fp_weights5(n, -6.0, 6.0, 2.23e-5, fptype, derivation))
    # rows the reference forces to zero (src/SM/FokkerPlanckMap.cpp:63-70,109-116)
    edge = [0, 1, n - 2, n - 1] if derivation == 4 else [0, n - 1]
    assert np.all(w[edge] == 0)


def test_impedance_models(capi, oracle):
    cases = [(256, 1e12, 47.7, 1e6, -1, 1, 0, 0, 0), (128, 3.7e12, 5.559, 2.7157e6, 0.032, 1, 0, 0, 0),
             (128, 3.7e12, 5.559, 2.7157e6, 0.032, 1, 5.8e7, 0, 0.008),
             (256, 1e12, 5.559, 2.7e6, 0.032, 0, 5.8e7, 0.1, 0.004), (64, 1e12, 5.559, 2.7e6, 0, 1, 1, 0, 0)]
    for c in cases:
        z, nz = capi.make_impedance(*c)
        zo, nzo = oracle.make_impedance(*c)
        assert nz == nzo
        if c[4] > 0 and c[5]:  # parallel plates, see above
            assert close_complex(z, zo), c
        else:
            assert np.array_equal(z, zo), c


def test_airy(capi):
    """series/asymptotic Airy functions of the product against scipy (independent of the
    quadrature used by the oracle and the libstdc++ Bessel functions used by the shim)"""
    from scipy.special import airy
    for x in [0.0] + list(np.logspace(-3, 2, 80)) + [6.4, 6.5, 6.6, 103.0]:
        want = airy(x)
        got, over = capi.airy(float(x))
        assert not over
        assert max(abs(got[i] / want[i] - 1) for i in range(4)) < 1e-12, x
    assert capi.airy(106.0)[1]


def test_invalid_arguments(capi):
    from inovesa_b200.params import Derived
    import ctypes as C
    lib = capi.load()
    w = np.zeros(40, f32)
    assert lib.ivs_fokker_planck_weights(8, 1.0, -1.0, 1e-3, 3, 4, w) == capi.ERR_INVALID_ARGUMENT
    assert lib.ivs_fokker_planck_weights(8, -1.0, 1.0, 1e-3, 3, 5, w) == capi.ERR_INVALID_ARGUMENT
    assert lib.ivs_simpson_weights(1, 0.1, w) == capi.ERR_INVALID_ARGUMENT


def test_impedance_from_a_file(tmp_path, capi):
    """Impedance(std::string datafile, f_max) (src/Z/Impedance.cpp:51-57,94-111) on the reference's own
    fixture and expectation (test/impedance.dat, test/TestImpedance.cpp:35-48: z[i] = (i - 2, 2.5 - i)),
    the repeated-line-number rule, and `*rv += Impedance(file, fmax)` (src/Z/ImpedanceFactory.cpp:70-75)"""
    import numpy as np
    f = tmp_path / "impedance.dat"
    f.write_text("1\t-2\t 2.5\n2\t-1\t 1.5\n3\t0\t 0.5\n4\t1\t-0.5\n5\t2\t-1.5\n6\t3\t-2.5\n")
    z = capi.read_impedance_file(f)
    assert z.dtype == np.complex64 and len(z) == 6
    assert np.array_equal(z, np.array([complex(i - 2, 2.5 - i) for i in range(6)], np.complex64))
    # a record whose line number repeats the previous one is skipped; no newline at the end of the file
    g = tmp_path / "dup.dat"
    g.write_text("1 1.0 0.0\n1 9.0 9.0\n2 2.0 -1.0\n7 3.5e0 4")
    assert np.array_equal(capi.read_impedance_file(g), np.array([1, 2 - 1j, 3.5 + 4j], np.complex64))
    with pytest.raises(capi.IvsError):
        capi.read_impedance_file(tmp_path / "missing.dat")
    # added on top of the generated wake impedance; a file shorter than the table is refused
    from inovesa_b200.params import kara_params
    p = kara_params(grid_size=16, padding=2.0)
    d = capi.derive(p)
    base, _ = capi.wake_impedance(p, d)
    n = d.padded_bins
    h = tmp_path / "full.dat"
    h.write_text("".join("%d %r %r\n" % (i + 1, 0.25 * i, -0.5 * i) for i in range(n + 3)))
    both, nz = capi.wake_impedance(p, d, h)
    extra = np.stack([0.25 * np.arange(n), -0.5 * np.arange(n)], 1).astype(np.float32)
    assert nz and np.array_equal(both, base + extra)
    with pytest.raises(capi.IvsError):
        capi.wake_impedance(p, d, f)

"""Host side of a train of bunches: the filling pattern of src/main.cpp:248-321 and the wake-function windows
(ivs_train_wake_windows) against the oracle's padded transform pair (CPU only)."""
import numpy as np
import pytest

from inovesa_b200.params import kara_params

f32 = np.float32


def test_train_layout_follows_main(capi):
    p = kara_params(grid_size=64)
    lay = capi.train_layout(p, [1e-4, 0, 2e-4, 3e-4])
    assert lay["buckets"] == [3, 1, 0]  # counted from the end of the pattern (:258-268)
    assert np.allclose(lay["fractions"], [1 / 6, 1 / 3, 1 / 2]) and abs(lay["Ib"] - 6e-4) < 1e-12
    d = lay["d"]
    spacing_ps = (1.0 / d.f_RF) * 2.99792458e8 / d.bl / p.pq_size
    assert lay["spacing_bins"] == int(round(64 * spacing_ps))
    assert lay["spaced_bins"] >= 64 * 4 * spacing_ps and lay["spaced_bins"] & (lay["spaced_bins"] - 1) == 0
    assert max(lay["buckets"]) * lay["spacing_bins"] + 64 <= lay["spaced_bins"]
    # ElectricField::_wakescaling = raw / nFreqs: the longer transform scales it down
    assert np.isclose(lay["wakescaling"] * lay["spaced_bins"], d.wakescaling * d.padded_bins, rtol=1e-6)
    with pytest.raises(ValueError):
        capi.train_layout(p, [0, 0])


def test_train_wake_windows_reproduce_the_padded_transform(capi, oracle, rng):
    n = 64
    p = kara_params(grid_size=n)
    lay = capi.train_layout(p, [1e-4, 0, 2e-4, 3e-4])
    d = lay["d"]
    z, _ = capi.make_impedance(lay["spaced_bins"], d.fmax, d.R_bend, d.f_rev, p.gap)
    w = capi.train_wake_windows(z, n, lay["buckets"], lay["spacing_bins"])
    prof = rng.random((3, n)).astype(f32)
    want = oracle.train_wake_potential(prof, lay["buckets"], lay["spacing_bins"], z, f32(lay["wakescaling"]))
    got = np.zeros((3, n))
    for b in range(3):
        for c in range(3):
            # sum_x' rho_c[x'] w[x - x' + n - 1] = full convolution, entries n-1 .. 2n-2
            got[b] += np.convolve(prof[c].astype(np.float64), w[b, c, :2 * n - 1])[n - 1:2 * n - 1]
    got *= lay["wakescaling"]
    assert np.abs(got - want).max() < 1e-6 * np.abs(want).max()
    # the window of a bunch onto itself is the single-bunch wake function
    z1 = z
    w1 = capi.train_wake_windows(z1, n, [0], lay["spacing_bins"])
    assert np.array_equal(w1[0, 0], w[0, 0]) and np.array_equal(w[1, 1], w[0, 0])
    with pytest.raises(capi.IvsError):
        capi.train_wake_windows(z, n, [9, 0], lay["spacing_bins"])

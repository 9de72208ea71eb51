"""BASELINE.json's configurations AS STATED, on the fused step kernel against the CPU oracle:
parallel-plates CSR at 512^2, parallel plates + resistive wall + collimator at 1024^2 from a
Haissinski start, the v1.1 synchrotron model at 2048^2 with 1e5 tracked particles -- and the
product's OWN host tables (ivs_host.cpp) instead of the oracle's on the device side.

The oracle evaluates the Airy functions of the parallel-plates model by quadrature (minutes per
table), so its tables for these configs are committed in tests/golden/impedance_tables.npz
(made by tests/golden/make_impedance_tables.py) and handed to the oracle simulation."""
import os

import numpy as np
import pytest

from conftest import rel_l1
from inovesa_b200.params import kara_params

pytestmark = pytest.mark.gpu
f32 = np.float32
L1_TOL = 1e-5  # BASELINE.json north_star
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "impedance_tables.npz")

CONFIG2 = dict(grid_size=512, steps_per_Ts=1000, gap=0.032, bunch_current=300e-6)
CONFIG3 = dict(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7, collimator=0.008,
               bunch_current=200e-6)
CONFIG5 = dict(grid_size=2048, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3)


@pytest.fixture(scope="module")
def ztables():
    return np.load(GOLDEN)


def _pair(capi, ctx, oracle, p, checkpoints, *, zw=None, zr=None, grid=None, particles=None, own_tables=True, flags=0):
    """the same run on the oracle and on the device; own_tables: the device side builds its tables
    with the product's host code (rf, drift, Fokker-Planck weights, impedance) instead of taking the
    oracle's"""
    osim = oracle.sim(p, grid, particles, zwake=zw, zrad=zr)
    tabs = None
    if not own_tables:
        t = osim.tables()
        tabs = dict(rf=t["rf"], drift=t["drift"], fp5=t["fp5"], zwake=t["zwake"] if np.any(t["zwake"]) else None)
    sim = capi.Simulation(ctx, p, 1, grid=grid, particles=particles, flags=flags, tables=tabs)
    out, done = [], 0
    try:
        for cp in checkpoints:
            sim.step(cp - done)
            osim.step(cp - done)
            done = cp
            g, pr, w, pt = sim.download()
            og, opr, ow, opt = osim.get()
            out.append(dict(step=cp, l1=rel_l1(g[0], og), prof=rel_l1(pr[0], opr),
                            wake=np.abs(w[0] - ow).max() / max(np.abs(ow).max(), 1e-30),
                            parts=None if pt is None else np.abs(pt[0] - opt).max()))
        assert sim.fallback_steps == 0
    finally:
        sim.close()
        osim.close()
    return out


@pytest.mark.parametrize("mode", ["strict", "fma"])
@pytest.mark.parametrize("path", ["fused", "unfused"])
def test_config1_with_the_products_own_tables(capi, ctx_strict, ctx_fma, oracle, mode, path):
    """config 1 for 1000 steps with NOTHING taken from the oracle on the device side"""
    ctx, orc = (ctx_strict if mode == "strict" else ctx_fma), oracle
    p = kara_params(grid_size=256, bunch_current=50e-6)
    res = _pair(capi, ctx, orc, p, [1, 100, 1000], flags=capi.SIM_UNFUSED if path == "unfused" else 0)
    for r in res:
        assert r["l1"] <= L1_TOL and r["prof"] <= L1_TOL, r


@pytest.mark.parametrize("own", [True, False], ids=["own-tables", "oracle-tables"])
def test_config2_parallel_plates_512(capi, ctx_strict, oracle, ztables, own):
    """configs[1]: 512^2, parallel-plates CSR impedance (src/Z/ParallelPlatesCSR.cpp:25-70), bunch
    current above the microbunching threshold, on the fused streaming kernel.  With its own tables
    the device side also exercises the product's Airy functions (the two parallel-plates tables
    differ in the last bit on < 2 % of the bins)."""
    p = kara_params(**CONFIG2)
    res = _pair(capi, ctx_strict, oracle, p, [1, 30, 60], zw=ztables["config2_512_zwake"],
                zr=ztables["config2_512_zrad"], own_tables=own)
    for r in res:
        assert r["l1"] <= L1_TOL and r["prof"] <= L1_TOL and r["wake"] <= 1e-4, r


@pytest.mark.parametrize("mode", ["strict", "fma"])
def test_config3_haissinski_start_1024(capi, ctx_strict, ctx_fma, oracle, ztables, mode):
    """configs[2]: 1024^2, parallel-plates CSR + resistive wall + collimator (gap > 0,
    src/Z/ImpedanceFactory.cpp:45-68), 4000 steps/T_s, Haissinski-initialised: the start distribution
    is computed by the product's host helper and handed as the same `data` array to the oracle and to
    the device (inc/PS/PhaseSpace.hpp:60).  It is stationary: the bunch profile of the oracle run
    stays put where a Gaussian start moves by per cent."""
    ctx, orc = (ctx_strict if mode == "strict" else ctx_fma), oracle
    p = kara_params(**CONFIG3)
    d = capi.derive(p)
    zw, zr = ztables["config3_1024_zwake"], ztables["config3_1024_zrad"]
    grid, residual = capi.haissinski_grid(p, d)  # the product's own impedance table
    assert residual < 1e-9
    res = _pair(capi, ctx, orc, p, [1, 20, 40], zw=zw, zr=zr, grid=grid)
    for r in res:
        assert r["l1"] <= L1_TOL and r["prof"] <= L1_TOL and r["wake"] <= 1e-4, r
    if mode == "strict":
        drift = {}
        for name, g in (("haissinski", grid), ("gaussian", None)):
            osim = oracle.sim(p, g, zwake=zw, zrad=zr)
            p0 = osim.get()[1].copy()
            osim.step(100)
            drift[name] = rel_l1(osim.get()[1], p0)
            osim.close()
        assert drift["haissinski"] < 5e-4 and drift["gaussian"] > 20 * drift["haissinski"], drift


@pytest.mark.parametrize("mode", ["strict", "fma"])
def test_config5_2048_with_1e5_tracked_particles(capi, ctx_strict, ctx_fma, oracle, mode):
    """configs[4] as stated: 2048^2, sin RF + alpha1/alpha2 drift, 1e5 passively tracked particles
    uniform in q, p in [-3, 3] (numpy default_rng(12345)), FPTrack 1 (deterministic),
    fused step + tracking kernel against the oracle: density L1 <= 1e-5, positions to 4 ulp of an fp32
    coordinate (5e-4 cells at x > 1024: 1e-4 cells is below the resolution of a float there)
    (KickMap::applyTo src/SM/KickMap.cpp:280-303, FokkerPlanckMap::applyTo src/SM/FokkerPlanckMap.cpp:214-268)"""
    ctx, orc = (ctx_strict if mode == "strict" else ctx_fma), oracle
    p = kara_params(**CONFIG5)
    p.fp_track = 1
    n = p.grid_size
    parts = (np.random.default_rng(12345).random((100000, 2)) * (n / 2) + n / 4).astype(f32)
    res = _pair(capi, ctx, orc, p, [1, 6, 12], particles=parts)
    for r in res:
        assert r["l1"] <= L1_TOL and r["prof"] <= L1_TOL, r
        assert r["parts"] <= 5e-4, r

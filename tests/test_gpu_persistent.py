"""The persistent kernel (csrc/ivs_resident.cu): a run that fits the machine at once (single
simulations and small ensembles up to 1024^2: BASELINE configs 1, 2, 3) does the steps of one
ivs_sim_step call inside ONE cooperative launch -- a CTA per SM keeps a block of rows for the whole
run, a grid-wide barrier stands where the kernel boundary was (the loop of src/main.cpp:948-1053
without leaving the device).  Per cell it does the arithmetic of the step kernels operation by
operation, with the same summation trees, so everything it produces equals the one-launch-per-step
path (IVS_SIM_NO_PERSISTENT) bit for bit."""
import numpy as np
import pytest

from inovesa_b200.params import kara_params

pytestmark = pytest.mark.gpu
f32 = np.float32

CASES = [
    (dict(grid_size=256), 1),
    (dict(grid_size=256), 4),                                          # a small ensemble, still one wave
    (dict(grid_size=256, gap=0.0), 1),                                 # no wake map
    (dict(grid_size=512, gap=0.032, bunch_current=300e-6, steps_per_Ts=2000), 1),
    (dict(grid_size=768, steps_per_Ts=2000), 2),
    (dict(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7, collimator=0.008), 1),
    (dict(grid_size=1024, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3), 1),
    (dict(grid_size=512, interpolation=3, derivation=3), 1),           # quadratic taps, 3-point stencil
    (dict(grid_size=512, steps_per_Ts=500), 1),                        # displacements of several cells, wide drift halo
    (dict(grid_size=2048, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3), 1),  # too large: one launch per step
    # ensembles whose members get few blocks of many rows each (64 + halo, 32 + halo)
    (dict(grid_size=256, bunch_current=80e-6), 30),
    (dict(grid_size=512, steps_per_Ts=2000), 9),
    # grid sizes that are multiples of 128 but not of 256: 1, 3, 5, 7 segments per row
    (dict(grid_size=128), 1),
    (dict(grid_size=384, steps_per_Ts=2000), 2),
    (dict(grid_size=640, gap=0.032, bunch_current=200e-6), 1),
    (dict(grid_size=896, steps_per_Ts=4000), 1),
]


@pytest.mark.parametrize("mode", ["strict", "fma"])
@pytest.mark.parametrize("case,members", CASES, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()) if isinstance(c, dict) else str(c))
def test_persistent_run_equals_one_launch_per_step(capi, ctx_strict, ctx_fma, mode, case, members):
    ctx = ctx_strict if mode == "strict" else ctx_fma
    p = kara_params(**case)
    a = capi.Simulation(ctx, p, members)
    b = capi.Simulation(ctx, p, members, flags=capi.SIM_NO_PERSISTENT)
    for k in (1, 2, 7, 24, 101):  # one direct launch, then even and odd numbers of steps in one launch
        l0 = ctx.launches
        a.step(k)
        la = ctx.launches - l0
        b.step(k)
        lb = ctx.launches - l0 - la
        assert lb == k, (k, lb)
        if p.grid_size % 256 == 0:
            assert la == (1 if k > 1 and p.grid_size <= 1024 else k), (k, la)  # the persistent kernel really ran
        else:
            assert la in (1, k, 2 * k), (k, la)  # wherever the plan has what the persistent kernel needs
        for x, y, what in zip(a.download(), b.download(), ("grid", "profile", "wake")):
            assert np.array_equal(x, y), (k, what, int((x != y).sum()))
    assert a.steps_done == b.steps_done == 135
    a.close(); b.close()


def test_persistent_run_is_not_used_where_it_does_not_fit(capi, ctx_fma, rng):
    """more CTAs than the machine holds at once, tracked particles: one launch per step as before"""
    p = kara_params(grid_size=256)
    big = capi.Simulation(ctx_fma, p, 200)  # more members than SMs
    l0 = ctx_fma.launches
    big.step(4)
    assert ctx_fma.launches - l0 == 4
    big.close()
    parts = (rng.random((64, 2)) * 200 + 28).astype(f32)
    trk = capi.Simulation(ctx_fma, p, 1, particles=parts)
    l0 = ctx_fma.launches
    trk.step(4)
    assert ctx_fma.launches - l0 == 8  # step kernel + tracking kernel per step
    trk.close()


def test_persistent_run_recovers_from_displacements_beyond_the_pad(capi, ctx_strict):
    """a step whose wake kick leaves the staged pad ends the persistent run for every CTA at the next
    barrier; the host redoes it through the kernel sequence and starts another run for the rest --
    the result equals the one-launch-per-step path's (which recovers the same way), bit for bit"""
    p = kara_params(grid_size=256, bunch_current=10e-3)
    a = capi.Simulation(ctx_strict, p, 1)
    b = capi.Simulation(ctx_strict, p, 1, flags=capi.SIM_NO_PERSISTENT)
    for k in (20, 33):
        a.step(k); b.step(k)
        for x, y, what in zip(a.download(), b.download(), ("grid", "profile", "wake")):
            assert np.array_equal(x, y), (k, what)
    assert a.fallback_steps == b.fallback_steps > 0
    a.close(); b.close()


def test_persistent_ensemble_stops_as_one_when_a_member_overflows(capi, ctx_fma):
    """the members of an ensemble do not read each other's row sums, but a step that leaves the pad in ONE
    member must end the run for ALL of them at the same step (the host rewinds the whole batch)"""
    p = kara_params(grid_size=256)
    currents = [50e-6, 10e-3, 100e-6, 200e-6]
    a = capi.Simulation(ctx_fma, p, 4, bunch_currents=currents)
    b = capi.Simulation(ctx_fma, p, 4, bunch_currents=currents, flags=capi.SIM_NO_PERSISTENT)
    for k in (20, 7):
        a.step(k); b.step(k)
        for x, y, what in zip(a.download(), b.download(), ("grid", "profile", "wake")):
            assert np.array_equal(x, y), (k, what)
    assert a.fallback_steps == b.fallback_steps > 0
    a.close(); b.close()

"""The column-block step kernel (csrc/ivs_stream8.cu, the default fused step for grids that are
multiples of 256) against the streaming kernel with the shared TMA ring (IVS_SIM_QPT2): the same
maps in the same order with the same summation trees, so every output is identical bit for bit --
density, bunch profile, wake potential, tracked particles.  Parity with the oracle and with the
reference's kernel sequence is in test_gpu_sim.py / test_gpu_configs.py (they run the default)."""
import numpy as np
import pytest

from inovesa_b200.params import kara_params

pytestmark = pytest.mark.gpu
f32 = np.float32

CASES = [
    dict(grid_size=256),
    dict(grid_size=256, gap=0.0),                                     # no wake map
    dict(grid_size=512, gap=0.032, bunch_current=300e-6, steps_per_Ts=2000),
    dict(grid_size=768, steps_per_Ts=2000),                           # 3 warps per CTA
    dict(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7, collimator=0.008),
    dict(grid_size=1536, steps_per_Ts=4000, linear_rf=0),
    dict(grid_size=2048, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3),  # the benchmark workload
    dict(grid_size=2048, steps_per_Ts=1000),                          # displacements of several cells
    dict(grid_size=512, interpolation=3, derivation=3),               # quadratic taps, 3-point stencil
]


def _run(capi, ctx, p, flags, steps, members=1, strips=0, particles=None, currents=None):
    sim = capi.Simulation(ctx, p, members, flags=flags, strips=strips, particles=particles, bunch_currents=currents)
    out = []
    for k in steps:
        sim.step(k)
        out.append(sim.download())
    sim.close()
    return out


@pytest.mark.parametrize("mode", ["strict", "fma"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_column_block_kernel_equals_streaming_kernel(capi, ctx_strict, ctx_fma, mode, case):
    ctx = ctx_strict if mode == "strict" else ctx_fma
    p = kara_params(**case)
    for members, strips in ((1, 0), (3, 0), (2, 5)):
        currents = None if members == 1 else [p.bunch_current * (0.5 + 0.5 * i) for i in range(members)]
        a = _run(capi, ctx, p, capi.SIM_QPT2, [1, 7, 24], members, currents=currents)
        b = _run(capi, ctx, p, capi.SIM_STREAM8, [1, 7, 24], members, strips, currents=currents)
        for (ga, pa, wa, _), (gb, pb, wb, _) in zip(a, b):
            assert np.array_equal(ga, gb), (members, strips)
            assert np.array_equal(pa, pb), (members, strips)
            assert np.array_equal(wa, wb), (members, strips)


def test_short_strips_and_single_rows(capi, ctx_fma):
    """strips of 2 ... 7 rows (what the planner picks for one small run: a strip per SM) and the
    longest ones (128 rows): results do not depend on the strip length"""
    p = kara_params(grid_size=256)
    (want, wp, ww, _), = _run(capi, ctx_fma, p, capi.SIM_QPT2, [9])
    for strips in (128, 64, 37, 9, 3, 2):
        (g, pr, w, _), = _run(capi, ctx_fma, p, capi.SIM_STREAM8, [9], strips=strips)
        assert np.array_equal(g, want) and np.array_equal(pr, wp) and np.array_equal(w, ww), strips


@pytest.mark.parametrize("n,fptrack", [(256, 1), (2048, 1), (256, 3)])
def test_particles_beside_the_next_step_equal_particles_in_line(capi, ctx_fma, rng, n, fptrack):
    """Tracked particles of a fused run step on a low-priority side stream while the next step
    already runs (two alternating wake buffers); IVS_SIM_TRACK_INLINE keeps their kernel in the
    context's stream after every step.  Same kernel, same inputs: identical positions, with CUDA
    graphs, with direct launches and with the step count split in any way."""
    case = dict(grid_size=n) if n == 256 else dict(grid_size=n, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3)
    p = kara_params(**case)
    p.fp_track = fptrack
    parts = (rng.random((5000, 2)) * (n - 60) + 30).astype(f32)
    (gi, pi, wi, qi), = _run(capi, ctx_fma, p, capi.SIM_TRACK_INLINE, [37], members=2, particles=parts)
    for flags, split in ((0, [37]), (capi.SIM_NO_GRAPH, [37]), (0, [1, 16, 3, 17])):
        g, pr, w, q = _run(capi, ctx_fma, p, flags, split, members=2, particles=parts)[-1]
        assert np.array_equal(q, qi), (flags, split)
        assert np.array_equal(g, gi) and np.array_equal(pr, pi) and np.array_equal(w, wi), (flags, split)
    assert np.abs(qi[0] - parts).max() > 1e-3  # they did move

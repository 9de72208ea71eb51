#!/usr/bin/env python
"""Device times of the kernels beside the fused step: RotationMap, the field kernel (wake potential
of one profile), the unfused maps and the out-step observables, with device-resident buffers.
Wall clock around `reps` back-to-back launches on the context's stream (sync on both sides).
Development aid / source of the figures in DESIGN.md section 8.

    python tools/kernel_times.py [reps]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inovesa_b200 import capi  # noqa: E402
from inovesa_b200.params import kara_params  # noqa: E402

f32 = np.float32


def timed(ctx, fn, reps):
    for _ in range(5):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e6


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    capi.load()
    peak = 6557.8
    with capi.Context(0, capi.ARITH_FMA) as ctx:
        L, h = ctx.lib, ctx.h
        rng = np.random.default_rng(1)
        for n in (256, 1024, 2048):
            g = rng.random((n, n), dtype=f32)
            d_in, d_out = ctx.to_device(g), ctx.malloc(g.nbytes)
            off = ((rng.random(n, dtype=f32) - 0.5) * 6).astype(f32)
            d_off = ctx.to_device(off)
            w5 = rng.random((n, 5), dtype=f32)
            d_w5 = ctx.to_device(w5)
            rows = []
            for name, fn in (
                ("rotate (RotationMap, tile kernel)", lambda: ctx.check(L.ivs_rotate(h, d_in, d_out, n, 0.0015, -6.0, 6.0, -6.0, 6.0, 4, 0))),
                ("rotate, angle 2 pi / 1000 * 10", lambda: ctx.check(L.ivs_rotate(h, d_in, d_out, n, 0.0628, -6.0, 6.0, -6.0, 6.0, 4, 1))),
                ("kick_y (KickMap Axis::y)", lambda: ctx.check(L.ivs_kick_y(h, d_in, d_out, n, d_off, 4))),
                ("kick_x (KickMap Axis::x)", lambda: ctx.check(L.ivs_kick_x(h, d_in, d_out, n, d_off, 4))),
                ("fokker_planck", lambda: ctx.check(L.ivs_fokker_planck(h, d_in, d_out, n, d_w5))),
            ):
                us = timed(ctx, fn, reps)
                rows.append((name, us, 8.0 * n * n / us / 1e3))
            for name, us, gbs in rows:
                print("n=%-5d %-36s %8.2f us  %7.1f GB/s (8 B/cell)  %5.1f %% of %.0f" % (n, name, us, gbs, 100 * gbs / peak, peak), flush=True)
            for p in (d_in, d_out, d_off, d_w5):
                ctx.free(p)
        # field kernel: wake potential of one profile (pad -> r2c -> x Z -> c2r), per padded length
        for n in (256, 512, 1024, 2048, 4096):
            nmax = 2 * n  # the reference pads to at least 2 n
            prof = np.exp(-0.5 * ((np.arange(n) - n / 2) / (n / 12)) ** 2).astype(f32)
            z = (rng.random((nmax, 2)) - 0.5).astype(f32)
            d_p, d_z, d_w = ctx.to_device(prof), ctx.to_device(z), ctx.malloc(4 * n)
            us = timed(ctx, lambda: ctx.check(L.ivs_wake_potential(h, d_p, n, d_z, nmax, 1.0, d_w, None)), reps)
            print("field kernel (ivs_wake_potential) n=%-5d nmax=%-5d %8.2f us" % (n, nmax, us), flush=True)
            for p in (d_p, d_z, d_w):
                ctx.free(p)
        # out-step observables of a simulation (sigma_z, sigma_E, CSR power: projections + field kernel)
        for n in (256, 2048):
            sim = capi.Simulation(ctx, kara_params(grid_size=n, steps_per_Ts=4000), 1)
            sim.step(8)
            us = timed(ctx, sim.observe, max(reps // 4, 20))
            print("ivs_sim_observe n=%-5d %8.2f us (with the sync and the 32-byte read-back)" % (n, us), flush=True)
            sim.close()

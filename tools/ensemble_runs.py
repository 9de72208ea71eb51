import sys, time
sys.path.insert(0, "/root/repo")
from inovesa_b200 import capi
from inovesa_b200.params import kara_params
capi.load()
with capi.Context(0, capi.ARITH_FMA) as ctx:
    for n, m in ((256, 30), (256, 37), (512, 9), (512, 18), (1024, 4)):
        for fl, name in ((0, "default"), (capi.SIM_NO_PERSISTENT, "per-step")):
            sim = capi.Simulation(ctx, kara_params(grid_size=n, steps_per_Ts=2000), m, flags=fl)
            sim.step(64); ctx.sync()
            l0 = ctx.launches; t0 = time.perf_counter(); sim.step(2000); ctx.sync(); dt = time.perf_counter() - t0
            print("n %4d members %2d %-8s %6.2f us/step  (%d launches)" % (n, m, name, dt / 2000 * 1e6, ctx.launches - l0), flush=True)
            sim.close()

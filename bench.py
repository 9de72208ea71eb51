#!/usr/bin/env python
"""bench.py -- throughput of the per-step Vlasov-Fokker-Planck update (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path, one rank per GPU
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle/_ref)

metric : phase-space cell-updates/s  (N^2 * members * steps / time of the step loop only)
workload (config.workload): BASELINE.json configs[4] -- 2048x2048 grid, sin RF (RFKickMap) +
  DriftMap with alpha1/alpha2, free-space CSR wake, cubic interpolation, 4000 steps/T_s -- run
  as a batch of `members` independent simulations per GPU (a bunch-current scan, the ensemble
  partitioning of SURVEY.md section 8e; default 26: 17 strips each = 442 CTAs, three full rounds
  over the 148 SMs) so that the state streamed per step (2 x members x 16 MiB) exceeds the 126 MB
  L2: the timing rule "inputs larger than L2" instead of an L2 flush.
  A single L2-resident 2048^2 run is reported next to it as `single_run`.
One "step" = one simulation step of every member (wake potential, wake kick, RF kick, drift,
Fokker-Planck, X projection, tracking of the passive particles): two launches, the fused step
kernel and the tracking kernel.  N > 1 GPUs: every rank runs its own block of the scan (replicas,
weak scaling, no collective on the step path); value = all ranks' cell-updates / max time.
roofline : the step kernel alone (the dominant kernel), timed with CUDA events on the same batch
  without tracked particles; achieved = members * N^2 * 8 B / launch duration against the measured
  HBM copy bandwidth of MEASURED_PEAKS.json; traffic = DRAM bytes of one launch from the ncu
  capture in profiles/.
e2e      : the same metric for one whole job through the C ABI with HOST buffers, everything inside
  the timed region: the start distribution uploaded from pinned host memory, the steps with the
  reference's out-step every 100 (observables + bunch profiles to the host), the final phase space
  downloaded.  `e2e.upload_every_step` is the same with the density re-uploaded before every single
  step (PCIe-bound by construction; round 1's e2e figure).
cpu_baseline / --impl reference : the reference's own CPU code (oracle/_ref) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from inovesa_b200.params import kara_params  # noqa: E402

f32 = np.float32
ALGO_BYTES_PER_CELL = 8.0  # one fp32 read + one fp32 write of the grid per step
# dram__bytes_read.sum + dram__bytes_write.sum of one stream8_step_kernel launch on the benchmark
# batch (26 x 2048^2), from the ncu --set full capture kept in profiles/r2_ncu_raw_metrics.txt
# (updated per round)
TRAFFIC_MEMBERS = 26
TRAFFIC_BYTES_PER_LAUNCH = 843513344  # r2: 448.5 MB read + 395.0 MB written; algorithmic 872.4 MB


def workload_params(name):
    if name == "config5":
        return kara_params(grid_size=2048, steps_per_Ts=4000, linear_rf=0, alpha1=2e-4, alpha2=3e-3,
                           gap=-1.0, bunch_current=50e-6, fp_track=3)
    if name == "config1":
        return kara_params(grid_size=256, steps_per_Ts=1000, gap=-1.0, bunch_current=50e-6)
    if name == "config2":
        return kara_params(grid_size=512, steps_per_Ts=1000, gap=0.032, bunch_current=300e-6)
    if name == "config3":
        return kara_params(grid_size=1024, steps_per_Ts=4000, gap=0.032, wall_conductivity=5.8e7,
                           collimator=0.008, bunch_current=200e-6)
    if name == "config4":
        return kara_params(grid_size=256, steps_per_Ts=1000, gap=-1.0, bunch_current=50e-6)
    raise SystemExit("unknown workload " + name)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        """nvidia-smi takes a while to produce its first line on a fresh box: do not time before it"""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        """statistics of the samples taken in [t_begin, t_end] (the timed region); if the region was
        shorter than the sampling interval, of the samples up to 0.25 s around it"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for _, r in self.rows]
        if t_begin is not None and t_end is not None:
            inside = [r for t, r in self.rows if t_begin <= t <= t_end]
            near = [r for t, r in self.rows if t_begin - 0.25 <= t <= t_end + 0.25]
            rows = inside if len(inside) >= 3 else (near or rows)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    from inovesa_b200 import capi

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    p = workload_params(args.workload)
    n = p.grid_size
    members = args.members
    arith = capi.ARITH_STRICT if args.arith == "strict" else capi.ARITH_FMA
    flags = capi.SIM_UNFUSED if args.unfused else 0
    flags |= {0: 0, 1: capi.SIM_QPT1, 2: capi.SIM_QPT2, 8: capi.SIM_STREAM8}[args.qpt]
    if args.track_inline:
        flags |= capi.SIM_TRACK_INLINE
    ctx = capi.Context(local_rank, arith)
    stream = torch.cuda.ExternalStream(ctx.lib.ivs_ctx_stream(ctx.h), device=local_rank)
    d = capi.derive(p)
    # synthetic input: the reference's Gaussian start distribution; members of the batch differ
    # in bunch current like a current scan (config 4's partitioning applied to this grid size)
    from inovesa_b200.ensemble import member_block, scan_currents
    lo, hi = member_block(members * world, world, rank)  # this rank's block of the scan
    currents = [float(c) for c in scan_currents(0.5 * p.bunch_current, 1.5 * p.bunch_current,
                                                members * world)[lo:hi]]
    rng = np.random.default_rng(12345 + rank)
    parts = None
    if args.particles:
        parts = (rng.random((args.particles, 2)) * (n / 2) + n / 4).astype(f32)  # q,p in [-3,3]
    sim = capi.Simulation(ctx, p, members, particles=parts, flags=flags, bunch_currents=currents,
                          seed=rank, strips=args.strips)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        fn(steps)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed_median(fn, steps, repeats):
        """`repeats` timed blocks of exactly `steps` steps each (CUDA events on the launching stream,
        barrier + synchronize on both sides, max over ranks); the median block is the result"""
        blocks = sorted(timed(fn, steps) for _ in range(repeats))
        return blocks[len(blocks) // 2], blocks

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # warm-up: W steps as asked, then one untimed rehearsal of the timed block itself (the step
    # graphs were captured and instantiated when the simulation was created)
    sim.step(args.warmup)
    sim.step(args.steps)
    if rank == 0:
        sampler.wait_first()
    launches0 = ctx.launches
    t_begin = time.perf_counter()
    ms, ms_blocks = timed_median(sim.step, args.steps, args.repeats)
    t_end = time.perf_counter()
    launches = (ctx.launches - launches0) // args.repeats
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    cells_per_step = float(n) * n * members
    value = cells_per_step * args.steps * world / (ms * 1e-3)

    # kernel-level view for the roofline: the dominant kernel is the fused step kernel.  With tracked
    # particles a step is two launches (step kernel + tracking kernel), so the kernel's own average
    # duration is measured on the same batch without particles, where a step is exactly that one
    # launch (CUDA events on the launching stream, same warm-up, same number of steps).
    peak, peak_src = measured_peak()
    ms_kernel, ms_kernel_blocks = ms, ms_blocks
    if args.particles and not args.unfused:
        simk = capi.Simulation(ctx, p, members, flags=flags, bunch_currents=currents, seed=rank,
                               strips=args.strips)
        simk.step(args.warmup)
        simk.step(args.steps)
        ms_kernel, ms_kernel_blocks = timed_median(simk.step, args.steps, args.repeats)
        simk.close()
    achieved = cells_per_step * ALGO_BYTES_PER_CELL / (ms_kernel * 1e-3 / args.steps) / 1e9

    # end to end through the C ABI with HOST buffers: per step upload the density from pinned
    # host memory (PhaseSpace::syncCLMem cpu2dev), advance one step, read the bunch profile
    # back (what main() consumes every step on the host side of the OpenCL path)
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    host_ptr = capi.P()
    nbytes = int(cells_per_step) * 4
    ctx.check(ctx.lib.ivs_host_alloc(ctx.h, nbytes, capi.C.byref(host_ptr)))
    host_grid = np.ctypeslib.as_array(capi.C.cast(host_ptr, capi.C.POINTER(capi.C.c_float)),
                                      shape=(int(cells_per_step),))
    host_grid[:] = sim.download(profile=False, wake=False, particles=False)[0].reshape(-1)
    prof = np.empty((members, n), f32)

    def e2e_loop(steps):
        for _ in range(steps):
            ctx.check(ctx.lib.ivs_sim_upload_grid(sim.h, host_grid))
            ctx.check(ctx.lib.ivs_sim_step(sim.h, 1))
            ctx.check(ctx.lib.ivs_sim_download(sim.h, None, capi._ptr(prof), None, None))

    e2e_loop(2)
    t0 = time.perf_counter()
    barrier()
    e2e_loop(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_every_step = cells_per_step * e2e_steps * world / e2e_s

    # the same ABI at the reference's out-step cadence (src/main.cpp:961-1033: the host looks at the
    # simulation every `outstep` = 100 steps): 100 steps, then observables and profile to the host
    out_steps, out_rounds = 100, max(1, min(5, args.steps // 100))
    obs = np.zeros((members, 8), f32)

    def outstep_loop(rounds):
        for _ in range(rounds):
            ctx.check(ctx.lib.ivs_sim_step(sim.h, out_steps))
            ctx.check(ctx.lib.ivs_sim_observe(sim.h, obs.reshape(-1)))
            ctx.check(ctx.lib.ivs_sim_download(sim.h, None, capi._ptr(prof), None, None))

    sim.set_radiation()
    outstep_loop(1)
    barrier()
    t0 = time.perf_counter()
    outstep_loop(out_rounds)
    barrier()
    out_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([out_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out_s = float(t.item())
    e2e_outstep = cells_per_step * out_steps * out_rounds * world / out_s

    # the whole job as the reference runs it (src/main.cpp:893-1075), through the C ABI with HOST buffers:
    # the start distribution goes up from pinned host memory once, every `outstep` = 100 steps the
    # observables and the bunch profiles come back, after the last step the phase space comes back
    # (the reference's final HDF5File::append(grid_t1)).  Nothing of it is outside the timed region.
    job_rounds = max(1, args.e2e_job_steps // out_steps)

    def job():
        ctx.check(ctx.lib.ivs_sim_upload_grid(sim.h, host_grid))
        outstep_loop(job_rounds)
        ctx.check(ctx.lib.ivs_sim_download(sim.h, capi._ptr(host_grid), None, None, None))

    barrier()
    t0 = time.perf_counter()
    job()
    barrier()
    job_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([job_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        job_s = float(t.item())
    job_steps = job_rounds * out_steps
    e2e_value = cells_per_step * job_steps * world / job_s
    ctx.check(ctx.lib.ivs_host_free(ctx.h, host_ptr))

    # single runs (no ensemble: the state is L2-resident, a step is latency-bound) at the grid sizes
    # of BASELINE.json's configs 1, 2, 3 and 5
    single_run = None
    if args.single_run:
        single_run = {}
        for key, wl in (("256", "config1"), ("512", "config2"), ("1024", "config3"), ("2048", "config5")):
            p1 = workload_params(wl)
            n1 = p1.grid_size
            k1 = max(args.steps, 256)
            entry = {}
            # default: the steps of one ivs_sim_step call inside one persistent kernel (grid-wide barrier
            # between steps); "per_step_launch": the same through the 16-step CUDA graph, one launch per step
            for name, fl in (("persistent", flags), ("per_step_launch", flags | capi.SIM_NO_PERSISTENT)):
                one = capi.Simulation(ctx, p1, 1, flags=fl)
                one.step(64)
                ms1, _ = timed_median(one.step, k1, 3)
                one.close()
                entry[name] = {"us_per_step": ms1 * 1e3 / k1, "steps_per_s": k1 / (ms1 * 1e-3),
                               "cell_updates_per_s": float(n1) * n1 * k1 / (ms1 * 1e-3)}
            single_run[key] = dict(entry["persistent"], per_step_launch=entry["per_step_launch"])
        single_run["note"] = ("one simulation per grid size (BASELINE configs 1, 2, 3, 5), state L2-resident "
                              "(no flush): the latency-bound view; %d steps per ivs_sim_step call" % max(args.steps, 256))

    # BASELINE.json configs[3]: the bunch-current scan of 256 independent 256^2 runs, 256 / N per GPU
    scan = None
    if args.scan:
        ps = workload_params("config4")
        total = 256
        lo4, hi4 = member_block(total, world, rank)
        cur4 = [float(c) for c in scan_currents(0.05e-3, 3e-3, total)[lo4:hi4]]
        ens = capi.Simulation(ctx, ps, hi4 - lo4, flags=flags, bunch_currents=cur4)
        k4 = max(args.steps, 256)
        ens.step(64)
        ms4, blocks4 = timed_median(ens.step, k4, 5)
        fb = ens.fallback_steps
        ens.close()
        scan = {"workload": "configs[3]: 256 independent 256^2 runs (bunch currents 0.05 ... 3 mA), %d per GPU" % (hi4 - lo4),
                "ensemble_steps_per_s": total * k4 / (ms4 * 1e-3),
                "cell_updates_per_s": 256.0 * 256 * total * k4 / (ms4 * 1e-3),
                "us_per_step": ms4 * 1e3 / k4, "steps": k4, "fallback_steps": fb,
                "frac_of_hbm_peak": 256.0 * 256 * (hi4 - lo4) * ALGO_BYTES_PER_CELL / (ms4 * 1e-3 / k4) / 1e9 / peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(p, args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": "phase-space cell-updates/s", "value": value, "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4]: 2048^2 grid, sin-RF RFKickMap + DriftMap(alpha1,alpha2), "
                                   "free-space CSR wake, cubic, 4000 steps/T_s; batch of %d members/GPU "
                                   "(bunch-current scan)" % members if args.workload == "config5"
                       else args.workload,
                       "grid": n, "members_per_gpu": members, "particles_per_member": args.particles,
                       "arith": args.arith, "step_path": "unfused" if args.unfused else "fused",
                       "l2": "inputs larger than L2 (state %d MiB/GPU)" % (2 * nbytes // 2**20),
                       "timing": "median of %d timed blocks of %d steps each" % (args.repeats, args.steps),
                       "parallelism": "replicas x%d (no collective on the step path)" % world},
            "ensemble_steps_per_s": members * world * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "cell-updates/s",
                    "h2d_bytes_per_step": nbytes // job_steps,
                    "d2h_bytes_per_step": (nbytes + job_rounds * int(prof.nbytes + obs.nbytes)) // job_steps,
                    "steps": job_steps,
                    "what": "one whole job through the C ABI with host buffers, all inside the timed region: "
                            "ivs_sim_upload_grid of the start distribution (pinned H2D, %d MB), %d x [ivs_sim_step(100) + "
                            "ivs_sim_observe + bunch profiles D2H] (the reference's out-step, src/main.cpp:961-1033), "
                            "ivs_sim_download of the final phase space (D2H, %d MB); bytes per step = the job's "
                            "bytes / its %d steps" % (nbytes >> 20, job_rounds, nbytes >> 20, job_steps),
                    "upload_every_step": {"value": e2e_every_step, "unit": "cell-updates/s", "steps": e2e_steps,
                                          "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(prof.nbytes),
                                          "what": "the density re-uploaded from the host before EVERY step (ivs_sim_upload_grid "
                                                  "+ ivs_sim_step(1) + profile D2H): PCIe-bound by construction; no caller of "
                                                  "the reference's loop does this, the state lives on the device there too"},
                    "at_outstep_cadence": {"value": e2e_outstep, "unit": "cell-updates/s",
                                           "what": "state resident; per 100 steps (the reference's out-step): "
                                                   "ivs_sim_step(100) + ivs_sim_observe (sigma_z, sigma_E, CSR "
                                                   "power) + profile D2H"}},
            "gpu_launches": int(launches),
            "repeats": {"n": args.repeats, "ms_blocks": ms_blocks, "kernel_ms_blocks": ms_kernel_blocks,
                        "spread": (ms_blocks[-1] - ms_blocks[0]) / ms},
            "fallback_steps": sim.fallback_steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH if not args.unfused and members == TRAFFIC_MEMBERS and n == 2048 else None,
                         "peak_source": peak_src,
                         "kernel": ("stream8_step_kernel" if args.qpt in (0, 8) else "stream_step_kernel") if not args.unfused else "kick/drift/fokker_planck/project sequence",
                         "us_per_launch": ms_kernel * 1e3 / args.steps,
                         "algorithmic_bytes_per_launch": cells_per_step * ALGO_BYTES_PER_CELL},
            "clocks": clocks,
        }
        # BASELINE.json also asks for the reference's OpenCL path on the B200 "if an ICD is present":
        # building it needs OpenCL headers and clFFT, neither of which exists in this image
        icd = os.path.isdir("/etc/OpenCL/vendors") and bool(os.listdir("/etc/OpenCL/vendors"))
        line["reference_opencl"] = ("not available: %s, and no OpenCL headers / clFFT to build the reference's "
                                    "OpenCL path with" % ("an ICD is registered" if icd else "no OpenCL ICD registered"))
        if single_run:
            line["single_run"] = single_run
        if scan:
            line["scan_256x256"] = scan
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    sim.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ CPU arms

def _ref_worker(args):
    """one replica of the workload on one core through the reference's own code (oracle/_ref)"""
    pdict, steps, warmup, budget_s, fma = args
    from oracle import pyoracle
    from inovesa_b200.params import RingParams
    p = RingParams(**pdict)
    kind = "reference" if pyoracle.available("ref", fma) else "port"
    impl = pyoracle.Ref(fma) if kind == "reference" else pyoracle.Oracle(fma)
    sim = impl.sim(p)
    sim.step(warmup)
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        sim.step(1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    sim.close()
    return kind, done, dt


def _pdict(p):
    return {name: getattr(p, name) for name, _ in p._fields_}


def cpu_baseline_sample(p, seconds):
    """bounded sample for the `cpu_baseline` object of our arm: the reference's CPU path, one
    core (the reference is single-threaded), a few steps of the same grid"""
    kind, done, dt = _ref_worker((_pdict(p), 10**6, 1, seconds, True))
    n = p.grid_size
    return {"value": float(n) * n * done / dt, "unit": "cell-updates/s", "cores": 1, "kind": kind,
            "sample": "%d steps of one %dx%d simulation in %.1f s on 1 core (reference is "
                      "single-threaded; built -O3 -march=x86-64-v3)" % (done, n, n, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    import multiprocessing as mp
    p = workload_params(args.workload)
    n = p.grid_size
    cores = max(1, min(os.cpu_count() or 1, args.cpu_replicas))
    budget = args.reference_seconds
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(_pdict(p), args.steps, args.warmup, budget, True)] * cores)
        wall = time.perf_counter() - t0
    kind = res[0][0]
    total_steps = sum(r[1] for r in res)
    tmax = max(r[2] for r in res)
    value = float(n) * n * total_steps / tmax
    line = {
        "impl": "reference", "metric": "phase-space cell-updates/s", "value": value,
        "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tmax / max(res[0][1], 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: 2048^2 grid, sin-RF + DriftMap(alpha1,alpha2), free-space "
                               "CSR wake, cubic, 4000 steps/T_s" if args.workload == "config5" else args.workload,
                   "grid": n, "replicas": cores},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": kind,
                         "sample": "%d independent replicas (one per core), %d steps each of %d requested "
                                   "(bounded to %.0f s), wall %.1f s" % (cores, res[0][1], args.steps, budget, wall)},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5")
    ap.add_argument("--members", type=int, default=26)
    ap.add_argument("--particles", type=int, default=100000)
    ap.add_argument("--arith", default="fma", choices=["fma", "strict"])
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--single-run", dest="single_run", action="store_true", default=True)
    ap.add_argument("--no-single-run", dest="single_run", action="store_false")
    ap.add_argument("--scan", dest="scan", action="store_true", default=True,
                    help="also time BASELINE configs[3] (256 x 256^2 scan, 256/N members per GPU)")
    ap.add_argument("--no-scan", dest="scan", action="store_false")
    ap.add_argument("--repeats", type=int, default=7, help="timed blocks of --steps steps; the median counts")
    ap.add_argument("--qpt", type=int, default=0, choices=[0, 1, 2, 8], help="tuning: quads of columns per thread")
    ap.add_argument("--strips", type=int, default=0, help="tuning: strips per member (0 = planner)")
    ap.add_argument("--track-inline", action="store_true",
                    help="tuning: tracking kernel in the step stream instead of on the side stream")
    ap.add_argument("--e2e-steps", type=int, default=20, help="steps of the upload-every-step loop")
    ap.add_argument("--e2e-job-steps", type=int, default=1000,
                    help="steps of the end-to-end job (upload once, out-steps every 100, final download)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cpu-replicas", type=int, default=64)
    ap.add_argument("--reference-seconds", type=float, default=120.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""The reference's main() (src/main.cpp:60-1110) on the B200 step: program options (options.py, the
reference's names and config-file format), derived scalars, start distribution, maps, the
simulation loop with its out-step cadence and the result file.

    python -m inovesa_b200.run -c ring.cfg -T 2 -o result.h5 [--GridSize 512 ...]

Writes `<output>.cfg` (the options that were set, as the reference does) and the HDF5 result file
in the reference's layout (results.py: through h5py where it is importable, else through the
pure-Python writer of h5native.py).  Needs a B200 and the built libivs.so; there is no CPU fallback.
"""
import sys
import time

import numpy as np

from . import capi
from .options import OptionError, ProgramOptions
from .results import run_and_record

f32 = np.float32
TWO_PI = 6.283185307179586


def read_tracking_file(path, d, n):
    """start positions of the particles to track: `q p` pairs in axis units (natural bunch lengths /
    energy spreads), turned into grid coordinates by PhaseSpace::x() / y() -- (q - qmin) / dq clamped
    to [0, n - 1], fp32 -- as src/main.cpp:762-771 and inc/PS/PhaseSpace.hpp:272-283 do; reading stops at
    the first token that is not a number, like `while (trackingfile >> q >> p)`"""
    pts = []
    toks = open(path).read().split()
    for k in range(0, len(toks) - 1, 2):
        try:
            q, p = f32(float(toks[k])), f32(float(toks[k + 1]))
        except ValueError:
            break
        x = min(max(f32(0.0), f32(f32(q - f32(d.qmin)) / f32(d.dq))), f32(n - 1.0))
        y = min(max(f32(0.0), f32(f32(p - f32(d.pmin)) / f32(d.dp))), f32(n - 1.0))
        pts.append((x, y))
    return np.asarray(pts, f32).reshape(-1, 2)


def read_start_distribution(path, step, nbunches=1):
    """makePSFromHDF5 / HDF5File::readPhaseSpace (src/IO/HDF5File.cpp:425-480): row (rows + step) % rows of
    /PhaseSpace/data of an earlier result file; the grid size is the file's"""
    try:
        import h5py
        with h5py.File(path, "r") as f:
            ps = np.asarray(f["/PhaseSpace/data"])
    except ImportError:
        from .h5native import read_dataset
        ps = read_dataset(path, "/PhaseSpace/data")
    if ps.ndim not in (3, 4) or ps.shape[0] == 0:
        raise OptionError("%s: /PhaseSpace/data holds no phase space" % path)
    row = ps[(ps.shape[0] + step) % ps.shape[0]]
    if row.ndim == 2:
        row = row[None]
    if row.shape[0] != nbunches:
        raise OptionError("%s: %d bunches in the file, %d in the filling pattern" % (path, row.shape[0], nbunches))
    if row.shape[1] != row.shape[2]:
        raise OptionError("%s: phase space is not square" % path)
    row = np.ascontiguousarray(row, f32)
    return row[0] if nbunches == 1 else row  # a train: [bunch][x][y], every bunch carrying its charge fraction


def rf_noise(opts, params, d, laststep, seed):
    """the queue of (phase, amplitude) pairs of a DynamicRFKickMap, or None for the static map
    (src/main.cpp:332-346,594-621)"""
    phase_noise = max(0.0, opts["RFPhaseSpread"] / 360.0 * TWO_PI)
    ampl_noise = max(0.0, opts["RFAmplitudeSpread"])
    mod_ampl = max(0.0, opts["RFPhaseModAmplitude"] / 360.0 * TWO_PI)
    mod_step = opts["RFPhaseModFrequency"] * d.dt

    def normal(x):  # std::fpclassify(x) == FP_NORMAL
        return x != 0.0 and np.isfinite(x) and abs(x) >= np.finfo(np.float64).tiny

    if not (normal(phase_noise) or normal(ampl_noise) or (normal(mod_ampl) and normal(mod_step))):
        return None
    m = capi.rf_model(params, d)
    return capi.rf_modulation(laststep, phase_noise, ampl_noise, mod_ampl, mod_step, d.revolutionpart,
                              m.syncphase, seed=seed)


def main(argv=None, out=sys.stdout):
    argv = sys.argv[1:] if argv is None else argv
    say = lambda msg: print(msg, file=out, flush=True)  # noqa: E731
    try:
        opts = ProgramOptions().parse(argv)
        bad = opts.unsupported()
        if bad:
            raise OptionError("not supported by this driver: " + "; ".join(bad))
        params = opts.ring_params()
    except OptionError as e:
        say("Error: %s" % e)
        return 1
    grid = None
    if opts["InitialDistFile"]:  # (.h5 / .hdf5: checked by ProgramOptions.unsupported)
        say('Reading in initial distribution from: "%s"' % opts["InitialDistFile"])
        try:
            grid = read_start_distribution(opts["InitialDistFile"], opts["InitialDistStep"], len(opts.bunch_currents()[1]))
        except (OptionError, OSError, KeyError, ValueError) as e:
            say("Error reading %s: %s" % (opts["InitialDistFile"], e))
            return 0  # (EXIT_SUCCESS, src/main.cpp:509-512)
        params.grid_size = grid.shape[-1]
    d = capi.derive(params)
    laststep = int(np.floor(d.steps * opts["rotations"] + 0.5)) if opts["rotations"] > 0 else 0
    ofname = opts["output"]
    if ofname and not ofname.endswith((".h5", ".hdf5")):
        say("Unkown filetype for output.")  # (sic, src/main.cpp:880)
        return 0
    if opts["gui"]:
        say("Warning: no GUI in this build; running without.")
    if opts["InterpolateClamped"]:
        say("Note: the reference's CPU path never clamps kick maps (src/SM/KickMap.cpp:194-271); ignored.")
    n = params.grid_size
    say("Grid %d x %d, %g steps per synchrotron period, %d steps, bunch current %g A" % (
        n, n, d.steps, laststep, params.bunch_current))
    say("Natural bunch length %.4g m, energy spread %.4g eV, padded profile %d bins" % (d.bl, d.dE, d.padded_bins))
    parts = read_tracking_file(opts["tracking"], d, n) if opts["tracking"] not in ("", "/dev/null") else None
    if parts is not None:
        say("Will do particle tracking with %d particles." % len(parts))
        if not len(parts):
            parts = None
    ctx = capi.Context(0, capi.ARITH_FMA)
    if grid is None:
        grid = capi.initial_grid(d, n, params.zoom)
    zfile = opts["Impedance"] or None
    if zfile:
        say('Reading impedance from: "%s"' % zfile)
    filling, bunches = opts.bunch_currents()
    lay = None
    if len(bunches) > 1:
        # a train of bunches (PhaseSpace::nb > 1, src/main.cpp:248-321): the members of the simulation are the
        # bunches, coupled through the wake potential of the whole padded train (capi.set_bunch_train); the
        # grids carry the charge fractions (PhaseSpace::normalize with `filling`, src/main.cpp:528-532)
        lay = capi.train_layout(params, filling)
        say("Filling pattern: %d bunches in %d buckets, %d bins apart, padded profile of the train %d bins" % (
            len(bunches), len(filling), lay["spacing_bins"], lay["spaced_bins"]))
        if grid.ndim == 2:  # (a start distribution read from a result file already holds the bunches)
            w = np.full(n, d.dq / 3.0)
            w[1:-1:2] *= 4.0
            w[2:-1:2] *= 2.0
            charge = float((grid.astype(np.float64) * w[:, None] * w[None, :]).sum())  # Simpson, both axes
            grid = np.stack([grid * f32(fr / charge) for fr in lay["fractions"]]).astype(f32)
        # (the column-block step kernel takes a train at N = 256, 512, ... 2048; other sizes step through the kernel sequence)
        # (tracked particles move in the fields of the first bunch, as the reference's applyTo reads the offsets of
        # bunch 0, src/SM/KickMap.cpp:280-303: the result file keeps member 0's)
        train_flags = 0 if (n % 256 == 0 and n <= 2048) else capi.SIM_UNFUSED
        sim = capi.Simulation(ctx, params, len(bunches), grid=grid, particles=parts, flags=train_flags, impedance_file=zfile)
    else:
        sim = capi.Simulation(ctx, params, 1, grid=grid, particles=parts, impedance_file=zfile)
        if opts["RenormalizeCharge"] >= 0:
            sim.renormalize()  # the initial renormalisation (src/main.cpp:528-532)
    seq = rf_noise(opts, params, d, laststep, seed=1)
    if seq is not None:
        say("Building dynamic, %s RFKickMap..." % ("linear" if params.linear_rf else "nonlinear"))
        sim.set_rf_modulation(seq)
    if lay is None:
        zw, nz = capi.wake_impedance(params, d, zfile)
    else:
        zw, nz = capi.wake_impedance(params, d, zfile, nfreqs=lay["spaced_bins"])  # src/main.cpp:718
        if nz:
            try:
                sim.set_bunch_train(lay["buckets"], lay["spacing_bins"], zw, lay["wakescaling"])
            except capi.IvsError as e:
                if e.code != capi.ERR_UNSUPPORTED or train_flags == capi.SIM_UNFUSED:
                    raise
                sim.close()  # no column-block plan for these maps: the kernel sequence takes any
                sim = capi.Simulation(ctx, params, len(bunches), grid=grid, particles=parts, flags=capi.SIM_UNFUSED, impedance_file=zfile)
                sim.set_bunch_train(lay["buckets"], lay["spacing_bins"], zw, lay["wakescaling"])
            if opts["RenormalizeCharge"] >= 0:
                sim.renormalize()  # every bunch to its charge fraction (src/main.cpp:528-532)
    say("Starting the simulation.")
    t0 = time.time()

    def progress(done, steps):
        if opts["verbose"] or done == steps:
            say("%8.3f T_s  (%d / %d steps, %.1f s)" % (done / d.steps, done, steps, time.time() - t0))

    rf = run_and_record(sim, params, d, laststep, outstep=opts["outstep"], save_phase_space=opts["SavePhaseSpace"],
                        wake_impedance=zw if nz else None, renormalize=opts["RenormalizeCharge"], progress=progress,
                        buckets=None if lay is None else lay["buckets"])
    dt = time.time() - t0
    if sim.fallback_steps:
        say("%d steps were redone through the kernel sequence (displacement beyond the fused kernel's pad)." % sim.fallback_steps)
    say("Finished %d steps in %.2f s (%.0f steps/s)." % (laststep, dt, laststep / max(dt, 1e-9)))
    if ofname:
        open(ofname + ".cfg", "w").write(opts.save_text())
        say('Saved configuiration to "%s.cfg".' % ofname)  # (sic)
        rf.to_hdf5(ofname)  # h5py where it exists, else the pure-Python writer (h5native.py)
        say('Saved results to "%s".' % ofname)
    else:
        say("Will not save results.")
    sim.close()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Result file with the reference's HDF5 layout.

The reference writes its results with HDF5-C++ (/root/reference/src/IO/HDF5File.cpp:29-349 lays
the file out, :360-423 appends one row per out-step).  Neither libhdf5 nor h5py exists in this
image, so the layout -- group and dataset paths, shapes, dtypes, unit attributes, soft-linked axes,
chunking + shuffle + deflate(6) -- is kept here as data and the rows are collected in memory:

  * `ResultFile.save(path)`   writes `<path>.npz` (one array per dataset path) + `<path>.layout.json`
                              (attributes, soft links, chunk shapes): the stated substitute;
  * `ResultFile.to_hdf5(path)` writes the real HDF5 file when h5py is importable (same paths,
                              chunked, shuffle, gzip 6, soft links), else raises ImportError.

Known quirks of the reference's file are kept and marked: `/Info/AxisValues_E` is written from
axis 0 (:168), `/CSR/Spectrum/data` holds nmax/2 of the nmax bins (:43,124-129), particle
coordinates are stored as axis values at the truncated integer grid index (:380-388).
Single bunch (nb = 1) like the rest of this repository.
"""
import json

import numpy as np

f32 = np.float32
DEFLATE = 6  # src/IO/HDF5File.cpp:541-544

# path -> (row shape as a function of the sizes, dtype, unit attributes)
SOFT_LINKS = {  # src/IO/HDF5File.cpp:201-301
    "/BunchProfile/axis0": "/Info/AxisValues_t", "/BunchProfile/axis1": "/Info/AxisValues_z",
    "/BunchLength/axis0": "/Info/AxisValues_t", "/BunchPopulation/axis0": "/Info/AxisValues_t",
    "/BunchPosition/axis0": "/Info/AxisValues_t",
    "/EnergyProfile/axis0": "/Info/AxisValues_t", "/EnergyProfile/axis1": "/Info/AxisValues_E",
    "/EnergyAverage/axis0": "/Info/AxisValues_t", "/EnergySpread/axis0": "/Info/AxisValues_t",
    "/RFKicks/axis0": "/Info/AxisValues_t", "/Particles/axis0": "/Info/AxisValues_t",
    "/WakePotential/axis0": "/Info/AxisValues_t", "/WakePotential/axis1": "/Info/AxisValues_z",
    "/CSR/Spectrum/axis0": "/Info/AxisValues_t", "/CSR/Spectrum/axis1": "/Info/AxisValues_f",
    "/CSR/Intensity/axis0": "/Info/AxisValues_t",
    "/PhaseSpace/axis1": "/Info/AxisValues_z", "/PhaseSpace/axis2": "/Info/AxisValues_E",
    "/Impedance/axis0": "/Info/AxisValues_f",
}
GROUPS = ["/BunchLength", "/BunchPopulation", "/BunchPosition", "/BunchProfile", "/CSR", "/CSR/Intensity",
          "/CSR/Spectrum", "/EnergyProfile", "/EnergySpread", "/EnergyAverage", "/Particles", "/PhaseSpace",
          "/RFKicks", "/Impedance", "/Impedance/data", "/Info", "/Info/Parameters", "/SourceMap",
          "/SourceMap/data", "/WakeFunction", "/WakePotential"]  # :553-580


class ResultFile:
    def __init__(self, n, nmax, n_particles=0, *, axis_q, axis_p, axis_f, bl, dE, f_scale, Ib, Qb,
                 t_sync, f_rev, volts=1.0, watt_per_hertz=1.0, watts=1.0, impedance=None, bucket=0, nb=1,
                 impedance_bins=None):
        """sizes and unit factors as the reference's constructor derives them (:151-318):
        axis_q/axis_p = PhaseSpace axes (n floats), axis_f = ElectricField frequency axis,
        bl/dE = "Meter"/"ElectronVolt" axis scales, t_sync = synchrotron period in s.  nb = PhaseSpace::nb
        (rows are [time][bunch][...]), bucket = the bunches' bucket numbers (src/main.cpp:258-268);
        impedance_bins = entries of the impedance table kept (nFreqs / 2; the spaced length for a train)."""
        self.n, self.nmax, self.np, self.nb = int(n), int(nmax), int(n_particles), int(nb)
        buckets = np.atleast_1d(np.asarray(bucket, np.uint32))
        assert buckets.shape == (self.nb,)
        c = 2.99792458e8
        self.fixed = {
            "/Info/AxisValues_z": np.asarray(axis_q, f32),
            "/Info/AxisValues_E": np.asarray(axis_q, f32),  # quirk: written from axis 0 (:168)
            "/Info/AxisValues_f": np.asarray(axis_f, f32)[: self.nmax // 2],
            "/Info/BucketNumbers": buckets,
            "/Info/Inovesa_v": np.asarray([1, 2, -1], np.int32),
        }
        if impedance is not None:
            z = np.asarray(impedance, f32).reshape(-1, 2)[: (impedance_bins or self.nmax // 2)]
            self.fixed["/Impedance/data/real"] = z[:, 0].copy()
            self.fixed["/Impedance/data/imag"] = z[:, 1].copy()
        n, h, nb = self.n, self.nmax // 2, self.nb
        self.rows = {p: [] for p in (
            "/Info/AxisValues_t", "/PhaseSpace/axis0", "/BunchPopulation/data", "/BunchProfile/data",
            "/BunchLength/data", "/BunchPosition/data", "/EnergyProfile/data", "/EnergySpread/data",
            "/EnergyAverage/data", "/Particles/data", "/RFKicks/data", "/WakePotential/data",
            "/CSR/Spectrum/data", "/CSR/Intensity/data", "/PhaseSpace/data")}
        self.row_shape = {
            "/Info/AxisValues_t": (), "/PhaseSpace/axis0": (), "/BunchPopulation/data": (nb,),
            "/BunchProfile/data": (nb, n), "/BunchLength/data": (nb,), "/BunchPosition/data": (nb,),
            "/EnergyProfile/data": (nb, n), "/EnergySpread/data": (nb,), "/EnergyAverage/data": (nb,),
            "/Particles/data": (self.np, 2), "/RFKicks/data": (2,), "/WakePotential/data": (nb, n),
            "/CSR/Spectrum/data": (nb, h), "/CSR/Intensity/data": (nb,), "/PhaseSpace/data": (nb, n, n)}
        self.attrs = {  # unit factors: value in the file * factor = physical unit (:151-318)
            "/Info/AxisValues_z": {"Meter": bl, "Second": bl / c},
            "/Info/AxisValues_E": {"ElectronVolt": dE},
            "/Info/AxisValues_f": {"Hertz": f_scale},
            "/Info/AxisValues_t": {"Second": t_sync, "Turn": t_sync * f_rev},
            "/PhaseSpace/axis0": {"Second": t_sync, "Turn": t_sync * f_rev},
            "/BunchPopulation/data": {"Ampere": Ib, "Coulomb": Qb},
            "/BunchProfile/data": {"AmperePerNBL": Ib, "CoulombPerNBL": Qb},
            "/BunchLength/data": {"Meter": bl, "Second": bl / c},
            "/BunchPosition/data": {"Meter": bl, "Second": bl / c},
            "/EnergyProfile/data": {"AmperePerNES": Ib, "CoulombPerNES": Qb},
            "/EnergySpread/data": {"ElectronVolt": dE}, "/EnergyAverage/data": {"ElectronVolt": dE},
            "/WakePotential/data": {"Volt": volts},
            "/CSR/Spectrum/data": {"WattPerHertz": watt_per_hertz}, "/CSR/Intensity/data": {"Watt": watts},
            "/PhaseSpace/data": {"AmperePerNBLPerNES": Ib, "CoulombPerNBLPerNES": Qb},
            "/Impedance/data": {"Ohm": 1.0},
        }
        self.parameters = {}
        self._axis_q, self._axis_p = np.asarray(axis_q, f32), np.asarray(axis_p, f32)

    # ---- the reference's append calls (src/IO/HDF5File.cpp:360-423)
    def add_parameter(self, name, value):
        """/Info/Parameters@name (ProgramOptions::save to HDF5 attributes, src/main.cpp:853)"""
        self.parameters[name] = value

    def append_time_step(self, t, profile, sigma_q, mean_q, energy_profile, sigma_p, mean_p, population=1.0):
        """HDF5File::append(ps, t, AppendType::Defaults): one out-step of the 1-D observables"""
        self._put("/Info/AxisValues_t", t)
        self._put("/BunchProfile/data", profile)
        self._put("/BunchLength/data", sigma_q)
        self._put("/BunchPosition/data", mean_q)
        self._put("/EnergyProfile/data", energy_profile)
        self._put("/EnergySpread/data", sigma_p)
        self._put("/EnergyAverage/data", mean_p)
        self._put("/BunchPopulation/data", population)

    def append_phase_space(self, t, grid):
        """HDF5File::append(ps, t, AppendType::PhaseSpace)"""
        self._put("/PhaseSpace/axis0", t)
        self._put("/PhaseSpace/data", grid)

    def append_csr(self, spectrum, power):
        """HDF5File::append(ef): the first nmax/2 bins of the spectrum, and the power"""
        self._put("/CSR/Spectrum/data", np.asarray(spectrum, f32).reshape(self.nb, -1)[:, : self.nmax // 2])
        self._put("/CSR/Intensity/data", power)

    def append_wake(self, wake):
        """HDF5File::append(wkm): the kick map's force"""
        self._put("/WakePotential/data", wake)

    def append_tracks(self, particles):
        """HDF5File::appendTracks: grid coordinates -> axis values at the truncated integer index"""
        p = np.asarray(particles, f32).reshape(self.np, 2)
        ix = np.clip(p[:, 0].astype(np.int64), 0, self.n - 1)
        iy = np.clip(p[:, 1].astype(np.int64), 0, self.n - 1)
        self._put("/Particles/data", np.stack([self._axis_q[ix], self._axis_p[iy]], 1))

    def append_rf_kicks(self, kicks):
        self._put("/RFKicks/data", kicks)

    def _put(self, path, value):
        a = np.asarray(value, f32).reshape(self.row_shape[path])
        self.rows[path].append(a.copy())

    # ---- output
    def datasets(self):
        out = dict(self.fixed)
        for p, rows in self.rows.items():
            shape = (len(rows),) + self.row_shape[p]
            out[p] = np.stack(rows).astype(f32) if rows else np.zeros(shape, f32)
        return out

    def layout(self):
        d = self.datasets()
        return {"groups": GROUPS, "soft_links": SOFT_LINKS, "deflate": DEFLATE, "shuffle": True,
                "datasets": {p: {"shape": list(a.shape), "dtype": str(a.dtype),
                                 "attrs": {k: float(v) for k, v in self.attrs.get(p, {}).items()}}
                             for p, a in d.items()},
                "group_attrs": {"/Impedance/data": self.attrs["/Impedance/data"],
                                "/Info/Parameters": {k: (v if isinstance(v, (int, float, str)) else str(v))
                                                     for k, v in self.parameters.items()}}}

    def save(self, path):
        d = self.datasets()
        np.savez_compressed(path + ".npz", **{p.strip("/").replace("/", "."): a for p, a in d.items()})
        json.dump(self.layout(), open(path + ".layout.json", "w"), indent=1)

    def to_hdf5(self, path, native=None):
        """write the HDF5 file: through h5py where it is importable (chunked, shuffled, deflated, extensible
        like the reference's), else -- or with native=True -- through the pure-Python writer of h5native.py
        (same groups, datasets, attributes and soft links, contiguous storage)"""
        if native is None:
            try:
                import h5py  # noqa: F401
                native = False
            except ImportError:
                native = True
        if native:
            from .h5native import write_hdf5
            attrs = {p: {k: np.float64(v) for k, v in kv.items()} for p, kv in self.attrs.items()}
            attrs["/Info/Parameters"] = {k: (v if isinstance(v, (int, float, str, bool, np.generic)) else str(v))
                                         for k, v in self.parameters.items()}
            return write_hdf5(path, self.datasets(), attrs, groups=GROUPS, soft_links=SOFT_LINKS)
        import h5py
        with h5py.File(path, "w") as f:
            for g in GROUPS:
                f.require_group(g)
            for p, a in self.datasets().items():
                unlimited = p in self.rows
                ds = f.create_dataset(p, data=a, maxshape=((None,) + a.shape[1:]) if unlimited else a.shape,
                                      chunks=True if a.size else None, shuffle=bool(a.size),
                                      compression="gzip" if a.size else None,
                                      compression_opts=DEFLATE if a.size else None)
                for k, v in self.attrs.get(p, {}).items():
                    ds.attrs[k] = np.float64(v)
            f["/Impedance/data"].attrs["Ohm"] = np.float64(1.0)
            for k, v in self.parameters.items():
                f["/Info/Parameters"].attrs[k] = v
            for link, target in SOFT_LINKS.items():
                f[link] = h5py.SoftLink(target)


def unit_factors(derived):
    """(Volt, WattPerHertz, Watt): ElectricField::volts = delta(1) * scale(1, "ElectronVolt") /
    revolutionpart (src/PS/ElectricField.cpp:23, revolutionpart a float), factor4WattPerHertz =
    2 * factor4Ohms * Ib^2 / f_rev with factor4Ohms = 1 (:39, inc/Z/Impedance.hpp:91) and factor4Watts =
    factor4WattPerHertz * c / bl, the "Hertz" scale of the frequency axis (:28-31,40)"""
    d = derived
    volts = float(f32(f32(d.dp) * f32(d.dE_f)) / f32(d.revolutionpart))  # float * float / float, then widened
    wph = 2.0 * d.Ib * d.Ib / d.f_rev
    return volts, wph, wph * float(f32(2.99792458e8 / float(f32(d.bl_f))))


def run_and_record(sim, params, derived, steps, outstep=100, save_phase_space=0, wake_impedance=None,
                   renormalize=0, progress=None, buckets=None):
    """main()'s loop with its out-step cadence (src/main.cpp:948-1053) on a `capi.Simulation` whose members
    are the bunches of the run (one, or a train set up with set_bunch_train; `buckets` = their bucket
    numbers): returns the filled ResultFile.  Time axis in synchrotron periods.
    renormalize > 0: the charge is renormalised before every renormalize-th step (:953-955);
    progress(done, steps): called at every out-step."""
    d, n = derived, params.grid_size
    axis_q = d.qmin + np.arange(n, dtype=f32) * f32(d.dq)
    axis_p = d.pmin + np.arange(n, dtype=f32) * f32(d.dp)
    nmax = d.padded_bins
    axis_f = np.arange(nmax, dtype=f32) * f32(1.0 / d.dq / (nmax - 1))
    rf = ResultFile(n, nmax, sim.np, axis_q=axis_q, axis_p=axis_p, axis_f=axis_f, bl=d.bl, dE=d.dE,
                    f_scale=2.99792458e8 / d.bl, Ib=d.Ib, Qb=d.Qb, t_sync=1.0 / d.fs, f_rev=d.f_rev,
                    volts=unit_factors(d)[0], watt_per_hertz=unit_factors(d)[1], watts=unit_factors(d)[2],
                    impedance=wake_impedance, nb=sim.members, bucket=[0] if buckets is None else buckets,
                    impedance_bins=None if wake_impedance is None else np.asarray(wake_impedance).reshape(-1, 2).shape[0] // 2)
    for name, _ in params._fields_:
        if not name.startswith("reserved"):
            rf.add_parameter(name, getattr(params, name))
    sim.set_radiation()
    done, k = 0, 0
    while True:
        obs = sim.observe()  # per bunch: [integral, mean_q, var_q, sigma_q, mean_p, var_p, sigma_p, P_csr]
        grid, prof, _, parts = sim.download()
        # the CSR spectrum and the wake potential of THIS profile: main() refreshes the wake before
        # the out-step block (src/main.cpp:949-952) and appends updateCSR's spectrum (:1000-1001)
        spectrum, wake = sim.fields()
        t = done / float(d.steps)
        projy = (grid.astype(np.float64) * _simpson(n, d.dq)[None, :, None]).sum(1).astype(f32)
        rf.append_time_step(t, prof, obs[:, 3], obs[:, 1], projy, obs[:, 6], obs[:, 4], obs[:, 0])
        rf.append_csr(spectrum, obs[:, 7])
        rf.append_wake(wake)
        if parts is not None:
            rf.append_tracks(parts[0])
        if save_phase_space and k % save_phase_space == 0:
            rf.append_phase_space(t, grid)
        if progress is not None:
            progress(done, steps)
        if done >= steps:
            break
        nxt = min(outstep, steps - done) if outstep > 0 else steps - done
        if renormalize > 0:
            # grid_t1->integrateAndNormalize() at the top of every renormalize-th step
            end = done + nxt
            while done < end:
                if done % renormalize == 0:
                    sim.renormalize()
                run = min(end, (done // renormalize + 1) * renormalize) - done
                sim.step(run)
                done += run
        else:
            sim.step(nxt)
            done += nxt
        k += 1
    if not save_phase_space or (k % save_phase_space) != 0:
        rf.append_phase_space(done / float(d.steps), sim.download()[0])  # always at the end (:1084-1086)
    return rf


def _simpson(n, delta):
    w = np.empty(n, np.float64)
    w[0] = w[-1] = 1.0
    w[1:-1:2] = 4.0
    w[2:-1:2] = 2.0
    return w * (delta / 3.0)

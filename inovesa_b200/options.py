"""The reference's program options (src/IO/ProgramOptions.cpp:22-305) for the driver in run.py:
same option names, short forms, defaults, config-file format (`name = value`, `#` comments) and
precedence (command line over config file over defaults; the compatibility names `RFVoltage`,
`SyncFreq`, `steps` are aliases, :284-292,365-368).

Options of subsystems that are out of scope here (GUI, OpenCL device, PNG / HDF5 start
distributions) are accepted where the reference ignores them without OpenCL / GUI
support, and refused with a message where they would change the physics.
"""
import os

from .params import default_params

# long name -> (short, type, default, RingParams field or None)
_B = "bool"
OPTIONS = {
    # physical parameters (:28-137)
    "alpha0": (None, float, 4e-3, "alpha0"),
    "alpha1": (None, float, 0.0, "alpha1"),
    "alpha2": (None, float, 0.0, "alpha2"),
    "SynchrotronFrequency": ("f", float, 0.0, "f_s"),
    "RevolutionFrequency": ("F", float, 9e6, "f_rev"),
    "DampingTime": ("d", float, -1.0, "t_damp"),
    "HarmonicNumber": ("H", float, 50.0, "harmonic"),
    "InitialDistFile": ("i", str, "", None),
    "InitialDistStep": (None, int, -1, None),
    "InitialDistZoom": (None, float, 1.0, "zoom"),
    "BunchCurrent": ("I", "floats", [1e-3], None),
    "BendingRadius": ("R", float, -1.0, "R_bend"),
    "BeamEnergy": ("E", float, 1.3e9, "E0"),
    "BeamEnergySpread": ("e", float, 4.7e-4, "sE"),
    "Impedance": ("Z", str, "", None),
    "VacuumGap": ("G", float, 0.03, "gap"),
    "UseCSR": (None, _B, True, "use_csr"),
    "CollimatorRadius": (None, float, 0.0, "collimator"),
    "WallConductivity": (None, float, 0.0, "wall_conductivity"),
    "WallSusceptibility": (None, float, 0.0, "wall_susceptibility"),
    "CutoffFreq": (None, float, 23e9, "f_cutoff"),
    "AcceleratingVoltage": ("V", float, 1e6, "V_RF"),
    "LinearRF": (None, _B, True, "linear_rf"),
    "RFAmplitudeSpread": (None, float, 0.0, None),
    "RFPhaseSpread": (None, float, 0.0, None),
    "RFPhaseModAmplitude": (None, float, 0.0, None),
    "RFPhaseModFrequency": (None, float, 0.0, None),
    # program parameters (:138-210)
    "cldev": (None, int, 0, None),
    "config": ("c", str, "default.cfg", None),
    "ForceOpenGLVersion": (None, int, 2, None),
    "gui": ("g", _B, False, None),
    "output": ("o", str, "", None),
    "outstep": ("n", int, 100, None),
    "SavePhaseSpace": (None, int, 0, None),
    "tracking": (None, str, "", None),
    "verbose": ("v", _B, False, None),
    "run_anyway": (None, _B, False, None),
    # non-physical simulation parameters (:211-267)
    "StepsPerTs": ("N", int, 1000, "steps_per_Ts"),
    "StepsPerRevolution": (None, float, 0.0, "steps_per_Trev"),
    "padding": ("p", float, 8.0, "padding"),
    "RoundPadding": (None, _B, True, "round_padding"),
    "PhaseSpaceSize": ("P", float, 12.0, "pq_size"),
    "PhaseSpaceShiftX": (None, float, 0.0, "shift_x"),
    "PhaseSpaceShiftY": (None, float, 0.0, "shift_y"),
    "RenormalizeCharge": (None, int, 0, None),
    "FPType": (None, int, 3, "fp_type"),
    "FPTrack": (None, int, 3, "fp_track"),
    "GridSize": ("s", int, 256, "grid_size"),
    "rotations": ("T", float, 5.0, None),
    "derivation": (None, int, 4, "derivation"),
    "InterpolationPoints": (None, int, 4, "interpolation"),
    "InterpolateClamped": (None, _B, False, None),
    # compatibility: ignored (:268-283)
    "HaissinskiIterations": (None, int, 0, None),
    "InitialDistParam": (None, int, 0, None),
    "RotationType": (None, int, 0, None),
    "SaveSourceMap": (None, _B, False, None),
}
ALIASES = {"RFVoltage": "AcceleratingVoltage", "SyncFreq": "SynchrotronFrequency", "steps": "StepsPerTs"}
CLI_ONLY = {"config"}                      # not allowed in a config file (:293-298)
IMPLICIT_TRUE = {"gui", "verbose", "run_anyway"}  # flags that may stand without a value on the command line
_SHORT = {v[0]: k for k, v in OPTIONS.items() if v[0]}


class OptionError(ValueError):
    pass


def _convert(name, kind, text):
    try:
        if kind is _B:
            t = text.strip().lower()
            if t in ("1", "true", "yes", "on"):
                return True
            if t in ("0", "false", "no", "off"):
                return False
            raise ValueError(text)
        if kind == "floats":
            return [float(x) for x in text.replace(",", " ").split()]
        if kind is int:
            return int(float(text)) if any(c in text for c in ".eE") else int(text)
        return kind(text)
    except ValueError:
        raise OptionError("the argument ('%s') for option '--%s' is invalid" % (text, name))


def _canonical(name):
    name = ALIASES.get(name, name)
    if name not in OPTIONS:
        raise OptionError("unrecognised option '%s'" % name)
    return name


def parse_config_text(text, where="config file"):
    """boost::program_options::parse_config_file: `name = value` lines, `#` starts a comment;
    a multitoken option (BunchCurrent) may be repeated"""
    out = {}
    for ln, raw in enumerate(text.splitlines(), 1):
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.startswith("["):
            raise OptionError("%s:%d: sections are not used by Inovesa's options" % (where, ln))
        if "=" not in line:
            raise OptionError("%s:%d: expected 'name = value'" % (where, ln))
        key, val = (t.strip() for t in line.split("=", 1))
        name = _canonical(key)
        if name in CLI_ONLY:
            raise OptionError("%s:%d: option '%s' is only valid on the command line" % (where, ln, key))
        kind = OPTIONS[name][1]
        v = _convert(name, kind, val)
        if kind == "floats" and name in out:
            out[name] = out[name] + v
        else:
            out[name] = v
    return out


def parse_command_line(argv):
    """`--Name value`, `--Name=value`, `-x value`, `-xvalue`; multitoken BunchCurrent takes every
    following token that is a number"""
    out = {}
    i = 0
    while i < len(argv):
        tok = argv[i]
        i += 1
        if tok.startswith("--"):
            key, eq, val = tok[2:].partition("=")
            name = _canonical(key)
        elif tok.startswith("-") and len(tok) >= 2 and tok[1] in _SHORT:
            name, eq, val = _SHORT[tok[1]], bool(tok[2:]), tok[2:]
        else:
            raise OptionError("too many positional options have been specified on the command line ('%s')" % tok)
        kind = OPTIONS[name][1]
        if not eq:
            nxt = argv[i] if i < len(argv) else None
            takes = nxt is not None and not (nxt.startswith("-") and not _is_number(nxt))
            if name in IMPLICIT_TRUE and (not takes or not _is_boolish(nxt)):
                out[name] = True
                continue
            if not takes:
                raise OptionError("the required argument for option '--%s' is missing" % name)
            val = nxt
            i += 1
        v = _convert(name, kind, val)
        if kind == "floats":
            while i < len(argv) and _is_number(argv[i]):
                v.append(float(argv[i]))
                i += 1
            out[name] = out.get(name, []) + v if name in out else v
        else:
            out[name] = v
    return out


def _is_number(t):
    try:
        float(t)
        return True
    except ValueError:
        return False


def _is_boolish(t):
    return t.strip().lower() in ("0", "1", "true", "false", "yes", "no", "on", "off")


class ProgramOptions:
    """values of all options after `parse(argv)`; `ring_params()` gives the struct the host
    helpers and the simulation object take"""

    def __init__(self):
        self.values = {k: (list(v[2]) if isinstance(v[2], list) else v[2]) for k, v in OPTIONS.items()}
        self.given = set()

    def parse(self, argv, read_file=None):
        """command line first, then the config file it names (values given on the command line
        win, :322-375); a missing default.cfg is not an error, any other missing file is"""
        cli = parse_command_line(list(argv))
        cfgname = cli.get("config", self.values["config"])
        cfg = {}
        if cfgname and cfgname != "/dev/null":
            if read_file is None:
                exists = os.path.isfile(cfgname)
                text = open(cfgname).read() if exists else None
            else:
                text = read_file(cfgname)
            if text is None:
                if cfgname != "default.cfg":
                    raise OptionError('Config file "%s" does not exist.' % cfgname)
            else:
                cfg = parse_config_text(text, cfgname)
        for src in (cfg, cli):
            for k, v in src.items():
                self.values[k] = v
                self.given.add(k)
        for k in ("output", "InitialDistFile"):
            if self.values[k] == "/dev/null":
                self.values[k] = ""
        return self

    def __getitem__(self, k):
        return self.values[_canonical(k)]

    def bunch_currents(self):
        """(filling pattern, positive entries): src/main.cpp:249-268"""
        filling = list(self.values["BunchCurrent"])
        return filling, [c for c in filling if c > 0]

    def ring_params(self):
        """the ring parameters; bunch_current = the whole beam's current Ib (src/main.cpp:274), i.e. the one
        bunch's, or the sum over a filling pattern (a train of bunches: run.py, capi.train_layout)"""
        _, bunches = self.bunch_currents()
        if not bunches:
            raise OptionError("BunchCurrent names no bunch")
        kw = {field: self.values[name] for name, (_, _, _, field) in OPTIONS.items() if field}
        for k in ("use_csr", "linear_rf", "round_padding"):
            kw[k] = 1 if kw[k] else 0
        kw["bunch_current"] = float(sum(bunches))
        return default_params(**kw)

    def unsupported(self):
        """options that would change the result and that this driver cannot honour"""
        bad = []
        f = self.values["InitialDistFile"]
        if f and not f.endswith((".h5", ".hdf5")):
            bad.append("InitialDistFile (start distribution from a PNG or text file; result files (.h5) are read)")
        return bad

    def save_text(self):
        """ProgramOptions::save (:393-430): every option that is not a default, `name=value` per line"""
        lines = []
        for k in sorted(self.given):
            if k in CLI_ONLY:
                continue
            v = self.values[k]
            if isinstance(v, list):
                lines += ["%s=%r" % (k, x) for x in v]
            elif isinstance(v, bool):
                lines.append("%s=%s" % (k, "true" if v else "false"))
            else:
                lines.append("%s=%s" % (k, v))
        return "\n".join(lines) + "\n"

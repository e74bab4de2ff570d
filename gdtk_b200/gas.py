"""Host-side gas models: what Eilmer's prep stage needs to build a FlowState.

Mirrors, for set-up purposes only (``FlowState:new{p=..., T=..., velx=...}``
in a job script), the reference's

* ``IdealGas``             src/gas/ideal_gas.d:30-144
* ``ThermallyPerfectGas``  src/gas/therm_perf_gas.d:230-249,394-447 with
  ``CEAThermoCurve``       src/gas/thermo/cea_thermo_curves.d:24-181

and reads the reference's Lua gas-model files (``setGasModel('file.lua')``)
with a small table parser, so job scripts keep naming the same files.
The per-cell thermodynamic update of the time-stepping path runs on the GPU;
nothing here is on the hot path.
"""
import math
import re

from . import _abi

R_UNIVERSAL = 8.31451  # src/gas/physical_constants.d:13


# --------------------------------------------------------------------------
# A minimal reader for the Lua tables written by `prep-gas`.

_TOKEN = re.compile(r"""
    (?P<ws>\s+|--[^\n]*)            |
    (?P<num>[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?) |
    (?P<str>'[^']*'|"[^"]*")        |
    (?P<name>[A-Za-z_][A-Za-z_0-9]*)|
    (?P<sym>[{}=,.;\[\]])
""", re.X)


def _tokens(text):
    pos = 0
    out = []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError(f"cannot parse Lua gas file near: {text[pos:pos + 40]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        out.append((kind, m.group(kind)))
    return out


class _Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def value(self):
        kind, v = self.next()
        if kind == "num":
            return float(v)
        if kind == "str":
            return v[1:-1]
        if kind == "name":
            if v == "true":
                return True
            if v == "false":
                return False
            if v == "nil":
                return None
            return v
        if (kind, v) == ("sym", "{"):
            return self.table()
        raise ValueError(f"unexpected token {v!r}")

    def table(self):
        fields = {}
        items = []
        while True:
            kind, v = self.peek()
            if (kind, v) == ("sym", "}"):
                self.next()
                break
            if kind == "name" and self.peek(1) == ("sym", "="):
                self.next()
                self.next()
                fields[v] = self.value()
            elif (kind, v) == ("sym", "["):
                self.next()
                key = self.value()
                assert self.next() == ("sym", "]")
                assert self.next() == ("sym", "=")
                fields[key] = self.value()
            else:
                items.append(self.value())
            if self.peek() in (("sym", ","), ("sym", ";")):
                self.next()
        if items and not fields:
            return items
        if items:
            for n, it in enumerate(items):
                fields[n + 1] = it
        return fields

    def chunk(self):
        env = {}
        while self.i < len(self.t):
            path = [self.next()[1]]
            while self.peek() == ("sym", "."):
                self.next()
                path.append(self.next()[1])
            assert self.next() == ("sym", "="), f"expected '=' after {'.'.join(path)}"
            val = self.value()
            d = env
            for p in path[:-1]:
                d = d.setdefault(p, {})
            d[path[-1]] = val
        return env


def read_lua_tables(path):
    """Return the global assignments of a prep-gas Lua file as nested dicts/lists."""
    with open(path) as f:
        return _Parser(_tokens(f.read())).chunk()


# --------------------------------------------------------------------------

class GasState:
    """rho, u, p, T, a, massf, rho_s (src/gas/gas_state.d:18-49)."""

    def __init__(self, nsp=1):
        self.rho = self.u = self.p = self.T = self.a = 0.0
        self.massf = [1.0] + [0.0] * (nsp - 1)
        self.rho_s = [0.0] * nsp


class IdealGas:
    """src/gas/ideal_gas.d"""
    kind = _abi.GAS_IDEAL
    n_species = 1

    def __init__(self, mMass=0.02896, gamma=1.4, name="air"):
        self.species_names = [name]
        self.mol_masses = [mMass]
        self.gamma = gamma
        self.Rgas = R_UNIVERSAL / mMass            # ideal_gas.d:64-68
        self.Cv = self.Rgas / (gamma - 1.0)
        self.Cvinv = 1.0 / self.Cv
        self.Cp = self.Rgas * gamma / (gamma - 1.0)

    def update_thermo_from_pT(self, Q):            # ideal_gas.d:89-97
        if Q.T <= 0.0 or Q.p <= 0.0:
            raise ValueError("Temperature and/or pressure was negative for update_thermo_from_pT.")
        Q.rho = Q.p / (Q.T * self.Rgas)
        Q.u = self.Cv * Q.T

    def update_sound_speed(self, Q):               # ideal_gas.d:136-143
        Q.a = math.sqrt(self.gamma * self.Rgas * Q.T)

    def fill_config(self, cfg):
        cfg.gas_model = self.kind
        cfg.n_species = 1
        cfg.ideal_mol_mass = self.mol_masses[0]
        cfg.ideal_gamma = self.gamma


class CEAThermoCurve:
    """src/gas/thermo/cea_thermo_curves.d:24-181 (Cp and h only)."""

    def __init__(self, R, T_break_points, T_blend_ranges, segments):
        self.R = R
        self.T_breaks = list(T_break_points)
        self.T_blends = list(T_blend_ranges)
        self.coeffs = [list(s) for s in segments]
        self.T_low, self.T_high = self.T_breaks[0], self.T_breaks[-1]
        self.Cp_low, self.Cp_high = self.eval_Cp(self.T_low), self.eval_Cp(self.T_high)
        self.h_low, self.h_high = self.eval_h(self.T_low), self.eval_h(self.T_high)

    def _coeffs(self, T):
        tb, bl, co = self.T_breaks, self.T_blends, self.coeffs
        if T < tb[1] - 0.5 * bl[0]:
            return co[0]
        if T > tb[-2] + 0.5 * bl[-1]:
            return co[-1]
        for i in range(1, len(tb) - 1):
            lo, hi = tb[i] - 0.5 * bl[i - 1], tb[i] + 0.5 * bl[i - 1]
            if lo <= T <= hi:
                wB = (1. / bl[i - 1]) * (T - lo)
                wA = 1.0 - wB
                return [wA * a + wB * b for a, b in zip(co[i - 1], co[i])]
            if T > hi and T < tb[i + 1] - 0.5 * bl[i]:
                return co[i]
        raise ValueError("Coefficients for CEA curve could not be determined.")

    def eval_Cp(self, T):
        if T < self.T_low:
            return self.Cp_low
        if T > self.T_high:
            return self.Cp_high
        a = self._coeffs(T)
        Cp_on_R = a[0] / (T * T) + a[1] / T + a[2] + a[3] * T
        Cp_on_R += a[4] * T * T + a[5] * T * T * T + a[6] * T * T * T * T
        return self.R * Cp_on_R

    def eval_h(self, T):
        logT = math.log(T)
        if T < self.T_low:
            return self.h_low - self.Cp_low * (self.T_low - T)
        if T > self.T_high:
            return self.h_high + self.Cp_high * (T - self.T_high)
        a = self._coeffs(T)
        h_on_RT = -a[0] / T + a[1] * logT + a[2] * T + a[3] * T * T / 2.0
        h_on_RT += a[4] * T * T * T / 3.0 + a[5] * T * T * T * T / 4.0 + a[6] * T * T * T * T * T / 5.0 + a[7]
        return self.R * h_on_RT


class ThermallyPerfectGas:
    """src/gas/therm_perf_gas.d (thermo part)."""
    kind = _abi.GAS_THERMALLY_PERFECT

    def __init__(self, species_names, db):
        self.species_names = list(species_names)
        self.n_species = len(species_names)
        if self.n_species > _abi.MAX_SPECIES:
            raise ValueError("too many species")
        self.mol_masses = [db[s]["M"] for s in species_names]
        self.R = [R_UNIVERSAL / m for m in self.mol_masses]      # therm_perf_gas.d:84
        self.thermo = []
        self.curves = []
        for s, R in zip(species_names, self.R):
            tc = db[s]["thermoCoeffs"]
            nseg = int(tc["nsegments"])
            segs = [tc[f"segment{i}"] for i in range(nseg)]
            self.thermo.append((nseg, tc["T_break_points"], tc["T_blend_ranges"], segs))
            self.curves.append(CEAThermoCurve(R, tc["T_break_points"], tc["T_blend_ranges"], segs))

    def _mass_average(self, Q, phi):               # gas_model.d:418-428
        result = 0.0
        for mf, x in zip(Q.massf, phi):
            result += mf * x
        return result

    def update_thermo_from_pT(self, Q):            # therm_perf_gas.d:230-234
        Rmix = 0.0
        for mf, R in zip(Q.massf, self.R):         # perf_gas_mix_eos.d:88-96
            Rmix += mf * R
        denom = Rmix * Q.T
        Q.rho = Q.p / denom                        # perf_gas_mix_eos.d:55-62
        vals = [c.eval_h(Q.T) - R * Q.T for c, R in zip(self.curves, self.R)]
        Q.u = self._mass_average(Q, vals)          # therm_perf_gas_mix_eos.d:61-67

    def update_sound_speed(self, Q):               # therm_perf_gas.d:394-404
        Cp = self._mass_average(Q, [c.eval_Cp(Q.T) for c in self.curves])
        Cv = self._mass_average(Q, [c.eval_Cp(Q.T) - R for c, R in zip(self.curves, self.R)])
        R = self._mass_average(Q, self.R)
        Q.a = math.sqrt((Cp / Cv) * (R * Q.T))

    def fill_config(self, cfg):
        cfg.gas_model = self.kind
        cfg.n_species = self.n_species
        for i, (M, (nseg, tb, bl, segs)) in enumerate(zip(self.mol_masses, self.thermo)):
            sp = cfg.species[i]
            sp.mol_mass = M
            sp.nsegments = nseg
            if nseg > _abi.MAX_SEGMENTS:
                raise ValueError("too many thermo segments")
            for j, v in enumerate(tb):
                sp.T_break_points[j] = v
            for j, v in enumerate(bl):
                sp.T_blend_ranges[j] = v
            for j, seg in enumerate(segs):
                for k in range(9):
                    sp.coeffs[j][k] = seg[k]


def set_gas_model(path):
    """``setGasModel(fname)`` of the job-script API: build the host gas model from a Lua file."""
    if path.endswith(".json"):
        import json
        with open(path) as f:
            env = json.load(f)
    else:
        env = read_lua_tables(path)
    if env.get("model") == "IdealGas":
        t = env["IdealGas"]
        return IdealGas(mMass=t["mMass"], gamma=t["gamma"], name=t.get("speciesName", "gas"))
    model = env.get("physical_model", env.get("model"))
    if model in ("thermally-perfect-gas", "ThermallyPerfectGas"):
        return ThermallyPerfectGas(env["species"], env["db"])
    raise ValueError(f"gas model {model!r} is outside the accelerated path (ideal and thermally-perfect only)")


class FlowState:
    """``FlowState:new{p=, T=, velx=, vely=, velz=, massf=}`` (src/eilmer/flowstate.d:600-700)."""

    def __init__(self, gmodel, p, T, velx=0.0, vely=0.0, velz=0.0, massf=None):
        nsp = gmodel.n_species
        Q = GasState(nsp)
        Q.p, Q.T = float(p), float(T)
        if massf is not None:
            if isinstance(massf, dict):
                Q.massf = [float(massf.get(n, 0.0)) for n in gmodel.species_names]
            else:
                Q.massf = [float(x) for x in massf]
        gmodel.update_thermo_from_pT(Q)
        gmodel.update_sound_speed(Q)
        Q.rho_s = [mf * Q.rho for mf in Q.massf]     # flowstate.d:653
        self.gas = Q
        self.vel = (float(velx), float(vely), float(velz))
        self.nsp = nsp

    def as_prims(self):
        """Values in EB200_PRIM order."""
        g = self.gas
        out = [g.rho, g.u, g.p, g.T, g.a, self.vel[0], self.vel[1], self.vel[2]]
        if self.nsp > 1:
            out += list(g.massf) + list(g.rho_s)
        return out

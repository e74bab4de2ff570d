"""Pin the oracle's gas-model functions against the reference's own unit-test values.

Golden values are the asserts of the reference's D unittests:
  src/gas/ideal_gas.d:224-239, src/gas/therm_perf_gas.d:565-593,
  src/gas/thermo/cea_thermo_curves.d:194-206, src/gas/thermo/perf_gas_mix_eos.d:97-116,
  src/gas/thermo/therm_perf_gas_mix_eos.d:177-199
with the same tolerances.  The same values are asserted for the host-side (prep) gas
models of gdtk_b200.gas.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from conftest import GAS_DATA
from gdtk_b200 import Config, FlowState, set_gas_model
from gdtk_b200.gas import GasState, ThermallyPerfectGas


def _sim(oracle, gmodel, **kw):
    cfg = Config(dimensions=2, flux_calculator="ausmdv", **kw)
    s = cfg.to_struct(gmodel)
    h = oracle.init(C.byref(s))
    assert h >= 0, oracle.error()
    return h


def _update(oracle, h, mode, rho=0.0, u=0.0, p=0.0, T=0.0, massf=()):
    q = np.array([rho, u, p, T, 0.0] + list(massf), dtype=np.float64)
    rc = oracle.gas_update(h, mode, q.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return dict(rho=q[0], u=q[1], p=q[2], T=q[3], a=q[4])


def close(a, b, rel):
    return math.isclose(a, b, rel_tol=rel)


def test_ideal_gas_kat(oracle):
    gm = set_gas_model(os.path.join(GAS_DATA, "ideal-air-gas-model.json"))
    assert close(gm.Rgas, 287.086, 1.0e-4)
    h = _sim(oracle, gm)
    r = _update(oracle, h, 0, p=1.0e5, T=300.0)
    assert close(r["rho"], 1.16109, 1.0e-4)
    assert close(r["u"], 215314.0, 1.0e-4)
    assert close(r["a"], 347.241, 1.0e-4)
    fs = FlowState(gm, p=1.0e5, T=300.0)
    # host model and oracle must agree to the last bit (same IEEE operations)
    assert (fs.gas.rho, fs.gas.u, fs.gas.a) == (r["rho"], r["u"], r["a"])
    # rhou inverse
    r2 = _update(oracle, h, 1, rho=r["rho"], u=r["u"])
    assert close(r2["T"], 300.0, 1e-14) and close(r2["p"], 1.0e5, 1e-14)
    oracle.finalize(h)


def test_thermally_perfect_gas_kat(oracle):
    gm = set_gas_model(os.path.join(GAS_DATA, "therm-perf-5-species-air.json"))
    assert gm.species_names == ["N2", "O2", "NO", "N", "O"]
    h = _sim(oracle, gm)
    mf = [0.2] * 5
    r = _update(oracle, h, 0, p=1.0e6, T=2000.0, massf=mf)
    assert close(r["u"], 11_801_825.6, 1.0e-6)
    assert close(r["rho"], 1.2840117, 1.0e-6)
    fs = FlowState(gm, p=1.0e6, T=2000.0, massf=mf)
    assert fs.gas.u == r["u"] and fs.gas.rho == r["rho"] and fs.gas.a == r["a"]
    r = _update(oracle, h, 1, rho=2.0, u=14.0e6, T=2000.0, massf=mf)
    assert close(r["p"], 3_373_757.4, 1.0e-6)
    assert close(r["T"], 4331.944, 1.0e-6)
    r = _update(oracle, h, 2, rho=1.5, T=10_000.0, massf=mf)
    assert close(r["p"], 5_841_068.3, 1.0e-6)
    assert close(r["u"], 20_340_105.9, 1.0e-6)
    r = _update(oracle, h, 3, rho=10.0, p=5.0e6, massf=mf)
    assert close(r["u"], 11_164_648.5, 1.0e-6)
    assert close(r["T"], 1284.012, 1.0e-6)
    oracle.finalize(h)


def test_newton_exit_state_quirk(oracle):
    """after update_thermo_from_rhou Q.u holds u(previous iterate),
    within Cv*1e-6 K of the target but not equal to it."""
    gm = set_gas_model(os.path.join(GAS_DATA, "therm-perf-5-species-air.json"))
    h = _sim(oracle, gm)
    mf = [0.767, 0.233, 0.0, 0.0, 0.0]
    r = _update(oracle, h, 1, rho=0.5, u=2.5e6, T=2500.0, massf=mf)
    assert r["u"] != 2.5e6
    assert abs(r["u"] - 2.5e6) < 1.0e-2        # |dx| < 1e-6 K times Cv ~ 1e3
    oracle.finalize(h)


def test_cea_curve_kat(oracle):
    gm = set_gas_model(os.path.join(GAS_DATA, "O-thermo.json"))
    h = _sim(oracle, gm)
    out = C.c_double(0.0)
    assert oracle.cea_eval(h, 0, 0, 500.0, C.byref(out)) == 0
    assert close(out.value, 1328.627, 1.0e-6)
    assert close(gm.curves[0].eval_Cp(500.0), 1328.627, 1.0e-6)
    assert oracle.cea_eval(h, 0, 1, 3700.0, C.byref(out)) == 0
    assert close(out.value, 20_030_794.683, 1.0e-6)
    assert close(gm.curves[0].eval_h(3700.0), 20_030_794.683, 1.0e-6)
    assert oracle.cea_eval(h, 0, 2, 10_000.0, C.byref(out)) == 0
    assert close(out.value, 14_772.717, 1.0e-3)
    oracle.finalize(h)


def test_perfect_gas_mix_eos_kat(oracle):
    # R = [297, 260] (N2, O2): build a two-species model whose molecular masses give those R.
    db = {n: {"M": 8.31451 / R, "thermoCoeffs": None} for n, R in (("N2", 297.0), ("O2", 260.0))}
    ref = set_gas_model(os.path.join(GAS_DATA, "therm-perf-5-species-air.json"))
    for n in db:
        db[n]["thermoCoeffs"] = {"nsegments": ref.thermo[0][0], "T_break_points": ref.thermo[0][1],
                                 "T_blend_ranges": ref.thermo[0][2],
                                 **{f"segment{i}": s for i, s in enumerate(ref.thermo[0][3])}}
    gm = ThermallyPerfectGas(["N2", "O2"], db)
    h = _sim(oracle, gm)
    r = _update(oracle, h, 2, rho=1.2, T=300.0, massf=[0.78, 0.22])
    assert close(r["p"], 103_989.6, 1.0e-6)
    r = _update(oracle, h, 0, p=103_989.6, T=300.0, massf=[0.78, 0.22])
    assert close(r["rho"], 1.2, 1.0e-6)
    r = _update(oracle, h, 3, rho=1.2, p=103_989.6, massf=[0.78, 0.22])
    assert close(r["T"], 300.0, 1.0e-6)
    oracle.finalize(h)


def test_tpg_mix_eos_kat(oracle):
    gm = set_gas_model(os.path.join(GAS_DATA, "O2-N2-H2.json"))
    h = _sim(oracle, gm)
    mf = [0.2, 0.7, 0.1]
    r = _update(oracle, h, 2, rho=1.0, T=1000.0, massf=mf)
    assert close(r["u"], 1_031_849.875, 1.0e-6)
    Q = GasState(3)
    Q.massf, Q.T, Q.p = mf, 1000.0, 1.0e5
    gm.update_thermo_from_pT(Q)
    assert close(Q.u, 1_031_849.875, 1.0e-6)
    # inverse: start the Newton iteration a little off, at 1500 K
    r = _update(oracle, h, 1, rho=1.0, u=r["u"], T=1500.0, massf=mf)
    assert close(r["T"], 1000.0, 1.0e-6)
    oracle.finalize(h)

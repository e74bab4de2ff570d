"""Regenerates tests/golden/gas/*.json from the reference's sample gas files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_gas_fixtures.py
Only the numeric tables the hot path needs are kept (M and thermoCoeffs per species).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gdtk_b200.gas import read_lua_tables  # noqa: E402

SRC = "/root/reference/src/gas/sample-data"
OUT = os.path.join(ROOT, "tests", "golden", "gas")


def tpg(fname, out):
    env = read_lua_tables(os.path.join(SRC, fname))
    species = env["species"]
    db = {s: {"M": env["db"][s]["M"], "thermoCoeffs": env["db"][s]["thermoCoeffs"]} for s in species}
    with open(os.path.join(OUT, out), "w") as f:
        json.dump({"physical_model": "thermally-perfect-gas", "species": species, "db": db}, f, indent=1)


def ideal(fname, out):
    env = read_lua_tables(os.path.join(SRC, fname))
    t = env["IdealGas"]
    with open(os.path.join(OUT, out), "w") as f:
        json.dump({"model": "IdealGas", "IdealGas": {"speciesName": t["speciesName"], "mMass": t["mMass"], "gamma": t["gamma"]}}, f, indent=1)


def o_atom(out):
    env = read_lua_tables(os.path.join(SRC, "O-thermo.lua"))
    db = {"O": {"M": 0.0159994, "thermoCoeffs": env["CEA_coeffs"]}}     # R = 8.31451/0.0159994 in the unittest
    with open(os.path.join(OUT, out), "w") as f:
        json.dump({"physical_model": "thermally-perfect-gas", "species": ["O"], "db": db}, f, indent=1)


def o2n2h2(out):
    env = read_lua_tables(os.path.join(SRC, "O2-N2-H2.lua"))
    species = env["species"]
    db = {s: {"M": env[s]["M"], "thermoCoeffs": env[s]["cea_thermo"]} for s in species}
    with open(os.path.join(OUT, out), "w") as f:
        json.dump({"physical_model": "thermally-perfect-gas", "species": species, "db": db}, f, indent=1)


if __name__ == "__main__":
    ideal("ideal-air-gas-model.lua", "ideal-air-gas-model.json")
    tpg("therm-perf-5-species-air.lua", "therm-perf-5-species-air.json")
    o_atom("O-thermo.json")
    o2n2h2("O2-N2-H2.json")
    print("wrote fixtures to", OUT)

"""Writes tests/golden/flow_*.npz: conserved quantities and dt histories of small runs of the CPU
oracle (oracle/oracle.c).  They freeze the oracle's output so that (a) any later change of the
oracle shows up as a diff, (b) the GPU tests can compare against committed vectors as well as
against a live oracle run.  NOTE: these vectors are produced by the restatement, not by a build of
the reference (no D compiler in the build container); the reference-derived golden
values are the gas-model unit-test numbers in tests/test_oracle_gas.py and the integration KATs
in tests/test_oracle_kats.py.

    python tests/golden/make_flow_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "sod3d": ("sod", dict(dims=3, ncells=40, nj=3, nk=2, nblocks=2), 25),
    "cone20_ausmdv": ("cone20", dict(nx0=6, nx1=14, ny=16), 40),
    "box3d_sheared_roe": ("box3d", dict(n=8, nb=2, sheared=True, flux_calculator="roe"), 6),
    "box3d_ausm_plus_up": ("box3d", dict(n=12, nb=1, flux_calculator="ausm_plus_up"), 6),
    "ffs_hanel": ("ffs", dict(nx=60, ny=20, flux_calculator="hanel"), 20),
    "cone20_adaptive_default": ("cone20", dict(nx0=6, nx1=14, ny=16, flux_calculator="adaptive_hanel_ausmdv"), 60),
    "ramp3d_default_euler": ("ramp3d", dict(), 120),
    "cone20_adaptive_efm": ("cone20", dict(nx0=6, nx1=14, ny=16, flux_calculator="adaptive"), 60),
    "box3d_ldfss2_rk3": ("box3d", dict(n=8, nb=1, flux_calculator="ldfss2", gasdynamic_update_scheme="tvd-rk3"), 5),
    # round 2: Eilmer 5's formulas, walls without ghost cells (one-sided stencils, wall fluxes, static inflow profile),
    # Roe's flux for a thermally perfect mixture
    "box3d_lmr_variant": ("box3d", dict(n=8, nb=2, solver_variant="lmr"), 6),
    "vortex_walls_without_ghost_cells": ("vortex", dict(gfactor=1), 40),
    "tpg_roe": ("tpg_box3d", dict(n=8, nb=1, flux_calculator="roe"), 4),
}
# not bit-comparable between glibc and CUDA: exp() in efm, pow() in the wall flux, log() in the thermally perfect gas
NOT_BITWISE = ("efm", "vortex", "tpg")


def run(lib, name):
    from gdtk_b200 import cases
    from util import run_case
    fac, kw, nsteps = CASES[name]
    sim, U, P = run_case(getattr(cases, fac), lib, nsteps, **kw)
    out = {"dt_history": np.array(sim.dt_history)}
    for bid, arrs in U.items():
        for q, a in enumerate(arrs):
            out[f"U_b{bid}_q{q}"] = a
    sim.close()
    return out


if __name__ == "__main__":
    from conftest import build_oracle
    from gdtk_b200 import _abi
    lib = _abi.load_library(build_oracle(), "orc_")
    only = sys.argv[1:]                      # names given on the command line: write only those
    for name in CASES:
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"flow_{name}.npz"), **run(lib, name))
        print("wrote", name)

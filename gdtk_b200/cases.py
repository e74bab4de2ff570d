"""Synthetic job descriptions for the configurations BASELINE.json names.

Each factory returns (config, gmodel, blocks) ready for ``Simulation``; they play the
role of the Lua job scripts under examples/eilmer in the reference:

* cone20        examples/eilmer/2D/sharp-cone-20-degrees/sg/cone20.lua  (C1)
* ffs           examples/eilmer/2D/forward-facing-step/ffs.lua          (C2)
* box3d         3D ideal-air box / sheared ramp, N x N x N blocks        (C3, C4)
* sod           examples/eilmer/3D/sod-shock-tube/sg (2D or 3D tube)
* tpg_box3d     thermally-perfect 5-species air box                      (C5)
"""
import math
import os

import numpy as np

from . import _abi
from .gas import FlowState, IdealGas, set_gas_model
from .grids import (box_grid_2d, box_grid_3d, quad_patch_grid, split_grid, connect_block_array, roberts_function, hex_volume_grid,
                    uniform_box_geometry)
from .sim import (Config, FluidBlock, InFlowBC_Supersonic, OutFlowBC_Simple, OutFlowBC_SimpleExtrapolate,
                  WallBC_WithSlip, identify_block_connections)
from .geometry import NG, geometry_3d, geometry_2d


def ideal_air():
    """examples' ideal-air-gas-model.lua: mMass 0.02896, gamma 1.4."""
    return IdealGas(mMass=0.02896, gamma=1.4, name="air")


def cone20(flux_calculator="ausmdv", nx0=10, nx1=30, ny=40, fbarray=False, gmodel=None, inflow=None, initial=None, **cfg_kw):
    """Mach 1.5 flow over a 20-degree cone, 2D axisymmetric, 2 blocks (C1).

    Geometry, states and settings of cone20.lua:26-66.  Differences, both forced by
    scope: config.flux_calculator is set explicitly (the reference default is the
    adaptive hanel/ausmdv blend) and the second patch uses the same
    straight-edged Coons patch as the first one instead of gridType="ao"."""
    gm = gmodel or ideal_air()
    cfg = Config(dimensions=2, axisymmetric=True, flux_calculator=flux_calculator,
                 max_time=5.0e-3, max_step=3000, cfl_value=0.5, extrema_clipping=False)
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    initial = initial or FlowState(gm, p=5955.0, T=304.0, velx=0.0)
    inflow = inflow or FlowState(gm, p=95.84e3, T=1103.0, velx=1000.0)
    a, b, c = (0.0, 0.0), (0.2, 0.0), (1.0, 0.29118)
    d, e, f = (1.0, 1.0), (0.2, 1.0), (0.0, 1.0)
    grid0 = quad_patch_grid(a, b, e, f, nx0, ny)
    grid1 = quad_patch_grid(b, c, d, e, nx1, ny)
    if fbarray:
        # sg-mpi/cone20.lua:48-51: FBArray{grid0, njb=2} and FBArray{grid1, nib=3, njb=2}, eight blocks
        blocks = []
        for grid, nib, njb, state in ((grid0, 1, 2, inflow), (grid1, 3, 2, initial)):
            for ib, jb, kb, sub in split_grid(grid, nib, njb):
                blk = FluidBlock(sub, state, id=len(blocks))
                if grid is grid0 and ib == 0:
                    blk.bcList["west"] = InFlowBC_Supersonic(inflow)
                if grid is grid1 and ib == nib - 1:
                    blk.bcList["east"] = OutFlowBC_Simple()
                blocks.append(blk)
        identify_block_connections(blocks, 2)
        return cfg, gm, blocks
    blk0 = FluidBlock(grid0, inflow, id=0)
    blk1 = FluidBlock(grid1, initial, id=1)
    identify_block_connections([blk0, blk1], 2)
    blk0.bcList["west"] = InFlowBC_Supersonic(inflow)
    blk1.bcList["east"] = OutFlowBC_Simple()
    return cfg, gm, [blk0, blk1]


def vortex_flow(gm):
    """The inviscid compressible vortex of udf-vortex-flow.lua:7-46 as a function of position -> FlowState
    (the file's own constants: R = 287 J/kg/K and gamma = 1.4 define p and T, the gas model turns them into a state)."""
    Rgas, g = 287.0, 1.4
    r_i, p_i, M_i, rho_i = 1.0, 100.0e3, 2.25, 1.0
    T_i = p_i / (Rgas * rho_i)
    a_i = np.sqrt(g * Rgas * T_i)
    u_i = M_i * a_i

    def state(x, y, z=0.0):
        r = np.sqrt(x * x + y * y)
        theta = np.arctan2(y, x)
        u = u_i * r_i / r
        t1 = r_i / r
        t2 = 1.0 + 0.5 * (g - 1.0) * M_i * M_i * (1.0 - t1 * t1)
        rho = rho_i * t2 ** (1.0 / (g - 1.0))
        p = p_i * (rho / rho_i) ** g
        T = p / (rho * Rgas)
        return FlowState(gm, p=float(p), T=float(T), velx=float(np.sin(theta) * u), vely=float(-np.cos(theta) * u))
    return state


def vortex(gfactor=4, nib=4, flux_calculator="adaptive_hanel_ausmdv", **cfg_kw):
    """examples/eilmer/2D/vortex-supersonic/vtx.lua: supersonic vortex in a 90-degree bend, inner and outer walls
    WallBC_WithSlip1 (no ghost cells: one-sided reconstruction and the wall flux), the exact vortex in the ghost
    cells of the inflow plane (UserDefinedBC), OutFlowBC_Simple at the exit; 4 blocks (FBArray nib = 4), the default
    flux calculator, 20 ms.  vtx-test.rb expects 2761 +- 3 steps and, against the exact solution, L2(p) = 800 +- 100 Pa
    and L2(T) = 0.405 +- 0.10 K (volume-weighted, flowsolution.d:306-346)."""
    from .sim import UserDefinedBC, WallBC_WithSlip1
    gm = ideal_air()
    cfg = Config(dimensions=2, flux_calculator=flux_calculator, max_time=20.0e-3, max_step=6000, dt_init=1.0e-6, cfl_value=0.5)
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    initial = FlowState(gm, p=1000.0, T=348.43, velx=0.0, vely=0.0)
    R_inner, R_outer = 1.0, 1.384
    nx, ny = int(20 * gfactor), int(10 * gfactor)
    # makePatch{north=Arc c->e, east=Line d->e, south=Arc b->d, west=Line b->c}: with arcs about the origin and
    # radial lines the Coons patch is the polar map (r along i: angle from 90 degrees down to 0, s along j: radius)
    r = (np.arange(nx + 1) / nx)[None, :]
    sv = (np.arange(ny + 1) / ny)[:, None]
    theta = 0.5 * np.pi * (1.0 - r)
    rad = R_inner + (R_outer - R_inner) * sv
    grid = (rad * np.cos(theta), rad * np.sin(theta))
    exact = vortex_flow(gm)
    blocks = []
    for ib, jb, kb, sub in split_grid(grid, nib, 1):
        blk = FluidBlock(sub, initial, id=len(blocks))
        blk.bcList["north"] = WallBC_WithSlip1()
        blk.bcList["south"] = WallBC_WithSlip1()
        if ib == 0:
            blk.bcList["west"] = UserDefinedBC(exact)
        if ib == nib - 1:
            blk.bcList["east"] = OutFlowBC_Simple()
        blocks.append(blk)
    identify_block_connections(blocks, 2)
    return cfg, gm, blocks


def ramp3d(flux_calculator="adaptive_hanel_ausmdv", **cfg_kw):
    """Mach 1.5 flow over a 10-degree ramp, 3D, two blocks (examples/eilmer/3D/simple-ramp/sg/ramp.lua):
    10x4x40 cells ahead of the ramp, 30x4x40 over it, k-lines clustered towards the wedge surface with
    RobertsFunction(end0=true, end1=false, beta=1.2), Euler update, the default flux calculator.  The
    reference's test expects 862 +- 3 steps to t = 5 ms (ramp-test.rb:33)."""
    gm = ideal_air()
    cfg = Config(dimensions=3, flux_calculator=flux_calculator, gasdynamic_update_scheme="euler",
                 max_time=5.0e-3, max_step=1000, dt_init=1.0e-6, cfl_value=0.5)
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    initial = FlowState(gm, p=5955.0, T=304.0, velx=0.0)
    inflow = FlowState(gm, p=95.84e3, T=1103.0, velx=1000.0)

    def box(x0, xs, ys=0.1, zs=1.0):
        return [[x0, 0, 0], [x0 + xs, 0, 0], [x0 + xs, ys, 0], [x0, ys, 0],
                [x0, 0, zs], [x0 + xs, 0, zs], [x0 + xs, ys, zs], [x0, ys, zs]]
    cluster_k = roberts_function(True, False, 1.2)
    grid0 = hex_volume_grid(box(0.0, 0.2), 11, 5, 41, cf_t=cluster_k)
    c1 = box(0.2, 0.8)
    c1[1][2] = c1[2][2] = 0.8 * math.tan(math.pi * 10.0 / 180.0)
    grid1 = hex_volume_grid(c1, 31, 5, 41, cf_t=cluster_k)
    blk0 = FluidBlock(grid0, initial, id=0)
    blk1 = FluidBlock(grid1, initial, id=1)
    blk0.bcList["west"] = InFlowBC_Supersonic(inflow)
    blk1.bcList["east"] = OutFlowBC_Simple()
    identify_block_connections([blk0, blk1], 3)
    return cfg, gm, [blk0, blk1]


def sod(dims=3, ncells=100, nj=2, nk=2, flux_calculator="ausmdv", nblocks=1, east_bc=None, west_bc=None, **cfg_kw):
    """Sod's shock tube along x (examples/eilmer/3D/sod-shock-tube/sg/sod.lua: L=1.0,
    high p=1e5,T=348.4 | low p=1e4,T=278.8, ideal air, t=0.6 ms)."""
    gm = ideal_air()
    cfg = Config(dimensions=dims, flux_calculator=flux_calculator, max_time=0.6e-3, max_step=5000,
                 dt_init=1.0e-6, cfl_value=0.5)
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    hi = FlowState(gm, p=1.0e5, T=348.4)
    lo = FlowState(gm, p=1.0e4, T=278.8)

    def init(x, y, z):
        return hi if x < 0.5 else lo
    if dims == 3:
        grid = box_grid_3d((0.0, 0.0, 0.0), (1.0, 0.1, 0.1), ncells, nj, nk)
        parts = split_grid(grid, nblocks, 1, 1)
    else:
        grid = box_grid_2d(0.0, 1.0, 0.0, 0.1, ncells, nj)
        parts = split_grid(grid, nblocks, 1)
    blocks = {}
    for n, (ib, jb, kb, sub) in enumerate(parts):
        blocks[(ib, jb, kb)] = FluidBlock(sub, init, id=n)
    connect_block_array(blocks, dims)
    if east_bc is not None:            # open the tube at an end (the reference's tube is closed: walls)
        blocks[(nblocks - 1, 0, 0)].bcList["east"] = east_bc
    if west_bc is not None:
        blocks[(0, 0, 0)].bcList["west"] = west_bc
    return cfg, gm, list(blocks.values())


class PerturbedState:
    """Lazy initial state of one block: base FlowState with rho (and consistently p, rho_s)
    perturbed by d(rho)/rho = 1e-3 sin(2 pi x) sin(2 pi y) sin(2 pi z) plus seeded uniform noise
    1e-6.  Cell centres are given analytically (origin + spacing) or read from
    the block geometry."""

    def __init__(self, base, seed, origin=None, spacing=None, amplitude=1.0e-3, noise=1.0e-6):
        self.base, self.seed, self.origin, self.spacing = base, seed, origin, spacing
        self.amplitude, self.noise = amplitude, noise

    def padded_arrays(self, geom):
        shp = (geom.NK, geom.NJ, geom.NI)
        if self.origin is not None:
            h, o = self.spacing, self.origin
            x = ((np.arange(geom.NI) - NG + 0.5) * h[0] + o[0])[None, None, :]
            y = ((np.arange(geom.NJ) - NG + 0.5) * h[1] + o[1])[None, :, None]
            z = ((np.arange(geom.NK) - geom.kg + 0.5) * h[2] + o[2])[:, None, None] if geom.dims == 3 else 0.0
        else:
            x, y, z = geom.pos
        pert = self.amplitude * np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y) * (np.sin(2 * np.pi * z) if geom.dims == 3 else 1.0)
        rng = np.random.default_rng(self.seed)
        fac = 1.0 + (pert + self.noise * (rng.random(shp) - 0.5))
        vals = self.base.as_prims()
        prims = [np.full(shp, v) for v in vals]
        prims[0] = prims[0] * fac
        prims[2] = prims[2] * fac      # p = rho R T at unchanged T and u
        nsp = self.base.nsp
        if nsp > 1:
            for i in range(nsp):
                prims[8 + nsp + i] = prims[8 + nsp + i] * fac   # rho_s = massf * rho
        return prims


def box3d(n=64, nb=2, flux_calculator="ausmdv", sheared=False, uniform_fast=True, seed=1234,
          gmodel=None, inflow=None, wall_bc=None, perturb=True, west_bc=None, closed=False, **cfg_kw):
    """3D ideal-air box (C3/C4): unit cube, n^3 cells in nb^3 blocks; inflow west, simple
    outflow east, slip walls elsewhere; initial state = inflow + smooth perturbation.
    sheared=True tilts the k-lines by 10 degrees (general-metric path, cf.
    examples/eilmer/3D/simple-ramp).  Uniform blocks share one geometry object."""
    gm = gmodel or ideal_air()
    cfg = Config(dimensions=3, flux_calculator=flux_calculator, max_step=10, max_time=1.0,
                 dt_init=1.0e-3, cfl_value=0.5)
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    inflow = inflow or FlowState(gm, p=95.84e3, T=1103.0, velx=1000.0)
    if n % nb:
        raise ValueError("n must be divisible by nb")
    m = n // nb
    h = 1.0 / n
    blocks = {}
    bid = 0
    shared = None
    amp = {} if perturb else {"amplitude": 0.0, "noise": 0.0}      # perturb=False: the uniform stream
    for ib in range(nb):
        for jb in range(nb):
            for kb in range(nb):
                x0, y0, z0 = ib * m * h, jb * m * h, kb * m * h
                if sheared or not uniform_fast:
                    X, Y, Z = box_grid_3d((x0, y0, z0), (x0 + m * h, y0 + m * h, z0 + m * h), m, m, m)
                    if sheared:
                        X = X + math.tan(math.radians(10.0)) * Z
                    geom = geometry_3d(X, Y, Z)
                    init = PerturbedState(inflow, seed + bid, **amp)
                else:
                    if shared is None:
                        shared = uniform_box_geometry(3, m, m, m, h, h, h)
                    geom = shared
                    init = PerturbedState(inflow, seed + bid, origin=(x0, y0, z0), spacing=(h, h, h), **amp)
                blk = FluidBlock(geom, init, id=bid)
                blocks[(ib, jb, kb)] = blk
                bid += 1
    connect_block_array(blocks, 3)
    for (ib, jb, kb), blk in blocks.items():
        if closed:                       # slip walls all round (the default of a face without a boundary condition)
            continue
        if ib == 0:
            blk.bcList["west"] = west_bc(gm, inflow) if west_bc is not None else InFlowBC_Supersonic(inflow)
        if ib == nb - 1:
            blk.bcList["east"] = OutFlowBC_Simple()
        if wall_bc is not None:          # the four side walls with another wall class (WallBC_WithSlip1: no ghost cells)
            for name, at_wall in (("south", jb == 0), ("north", jb == nb - 1), ("bottom", kb == 0), ("top", kb == nb - 1)):
                if at_wall:
                    blk.bcList[name] = wall_bc()
    cfg.block_index = {blk.id: key for key, blk in blocks.items()}
    return cfg, gm, list(blocks.values())


def sheared_inflow_profile(gm, inflow):
    """A UserDefinedBC for box3d's inflow plane whose ghost-cell FlowStates vary with y and z (a static profile:
    exercises the per-ghost-cell table and its ordering in 3D)."""
    from .sim import UserDefinedBC
    p0, T0, u0 = inflow.gas.p, inflow.gas.T, inflow.vel[0]

    def state(x, y, z):
        return FlowState(gm, p=p0 * (1.0 + 0.02 * y), T=T0 * (1.0 + 0.01 * z), velx=u0 * (1.0 + 0.03 * y * z), vely=5.0 * z, velz=-3.0 * y)
    return UserDefinedBC(state)


def tpg_subset_model(species):
    """A thermally perfect mixture of some of the species of the 5-species air file (the same CEA curves), as a
    gas model plus free-stream mass fractions: lets the tests run other species counts than five."""
    import json
    import tempfile
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gas")
    with open(os.path.join(root, "therm-perf-5-species-air.json")) as f:
        d = json.load(f)
    d["species"] = list(species)
    d["db"] = {k: d["db"][k] for k in species}
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(d, f)
        path = f.name
    try:
        gm = set_gas_model(path)
    finally:
        os.unlink(path)
    air = {"N2": 0.767 * 0.99, "O2": 0.233 * 0.99, "NO": 0.004, "N": 0.003, "O": 0.003}
    tot = sum(air[k] for k in species)
    return gm, {k: air[k] / tot for k in species}


def tpg_box3d(n=32, nb=2, gas_file=None, species=None, **kw):
    """C5: thermally-perfect 5-species air (N2, O2, NO, N, O), frozen chemistry, T = 3000 K.
    species: a subset of those five (other species counts)."""
    if species is not None:
        gm, massf = tpg_subset_model(species)
    else:
        gas_file = gas_file or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "tests", "golden", "gas", "therm-perf-5-species-air.json")
        gm = set_gas_model(gas_file)
        massf = {"N2": 0.767 * 0.99, "O2": 0.233 * 0.99, "NO": 0.004, "N": 0.003, "O": 0.003}
    inflow = FlowState(gm, p=95.84e3, T=3000.0, velx=3000.0, massf=massf)
    return box3d(n=n, nb=nb, gmodel=gm, inflow=inflow, **kw)


def tpg_air(gas_file=None):
    """The 5-species thermally perfect air of C5 and its free-stream composition."""
    gas_file = gas_file or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        "tests", "golden", "gas", "therm-perf-5-species-air.json")
    gm = set_gas_model(gas_file)
    massf = {"N2": 0.767 * 0.99, "O2": 0.233 * 0.99, "NO": 0.004, "N": 0.003, "O": 0.003}
    return gm, massf


def tpg_ffs(nx=120, ny=40, **kw):
    """The forward-facing step with the thermally perfect 5-species air (2D, uniform-Cartesian blocks)."""
    gm, massf = tpg_air()
    probe = FlowState(gm, p=101.325e3, T=2500.0, massf=massf)
    inflow = FlowState(gm, p=101.325e3, T=2500.0, velx=3.0 * probe.gas.a, massf=massf)
    return ffs(nx=nx, ny=ny, gmodel=gm, inflow=inflow, **kw)


def tpg_cone20(**kw):
    """cone20's geometry (2D axisymmetric, general-metric blocks) with the thermally perfect 5-species air."""
    gm, massf = tpg_air()
    initial = FlowState(gm, p=5955.0, T=2000.0, massf=massf)
    inflow = FlowState(gm, p=95.84e3, T=3000.0, velx=1500.0, massf=massf)
    return cone20(gmodel=gm, inflow=inflow, initial=initial, **kw)


def ffs(nx=384, ny=128, flux_calculator="ausmdv", uniform_fast=True, gmodel=None, inflow=None, i_step=None, **cfg_kw):
    """Mach-3 forward-facing step (C2): domain [0,3]x[0,1], step at x=0.6, height 0.2
    (examples/eilmer/2D/forward-facing-step/ffs.lua:21-32), three blocks like the example:
    blk0 [0,0.6]x[0,0.2], blk1 [0,0.6]x[0.2,1], blk2 [0.6,3]x[0.2,1].  nx, ny are the cell
    counts of the bounding grid (4096 x 1024 for the benchmark); dx = 3/nx, dy = 1/ny.  i_step: the cell index of
    the step face (default: the nearest to x = 0.6); the benchmark picks the nearest EVEN index, so that every block
    has an even padded width and its rows can be staged by TMA (16-byte strides)."""
    gm = gmodel or ideal_air()
    cfg = Config(dimensions=2, flux_calculator=flux_calculator, max_step=10, max_time=1.0,
                 dt_init=1.0e-3, cfl_value=0.5)
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    if inflow is None:
        T0 = 300.0
        a0 = math.sqrt(gm.gamma * gm.Rgas * T0)
        inflow = FlowState(gm, p=101.325e3, T=T0, velx=3.0 * a0)
    dx, dy = 3.0 / nx, 1.0 / ny
    i_step, j_step = (int(i_step) if i_step is not None else int(round(0.6 / dx))), int(round(0.2 / dy))
    specs = [(0, i_step, 0, j_step), (0, i_step, j_step, ny), (i_step, nx, j_step, ny)]
    blocks = []
    for bid, (i0, i1, j0, j1) in enumerate(specs):
        if uniform_fast:
            geom = uniform_box_geometry(2, i1 - i0, j1 - j0, 1, dx, dy)
        else:
            X, Y = box_grid_2d(i0 * dx, i1 * dx, j0 * dy, j1 * dy, i1 - i0, j1 - j0)
            geom = geometry_2d(X, Y)
        blocks.append(FluidBlock(geom, inflow, id=bid))
    from .sim import ExchangeBC_FullFace
    blocks[0].bcList["north"] = ExchangeBC_FullFace(1, _abi.SOUTH)
    blocks[1].bcList["south"] = ExchangeBC_FullFace(0, _abi.NORTH)
    blocks[1].bcList["east"] = ExchangeBC_FullFace(2, _abi.WEST)
    blocks[2].bcList["west"] = ExchangeBC_FullFace(1, _abi.EAST)
    blocks[0].bcList["west"] = InFlowBC_Supersonic(inflow)
    blocks[1].bcList["west"] = InFlowBC_Supersonic(inflow)
    blocks[2].bcList["east"] = OutFlowBC_Simple()
    return cfg, gm, blocks

"""A prepared Eilmer job on disk: what ``e4shared --prep`` leaves and ``e4shared --run`` reads.

    <dir>/config/<job>.config      JSON written by src/eilmer/output.lua (write_config_file) with one
                                   "block_N" object per FluidBlock (fluidblock.lua:183-260) whose
                                   "boundary_<face>" objects list the ghost-cell and flux effects
                                   (bc.lua:706-760 and the tojson methods of the effects)
    <dir>/config/<job>.times       tindx, time, dt of every written solution (simcore_io.d:126-131)
    <dir>/grid/t0000/<job>.grid.bBBBB.t0000.gz    and    <dir>/flow/tNNNN/<job>.flow.bBBBB.tNNNN.gz

``load_job`` turns that into the (Config, gas model, FluidBlocks, history points) the Simulation class
takes, for the options and boundary conditions of the accelerated path; anything else raises.
``write_job`` is the inverse (what the preparation stage would write for a job built in Python), so that
the two can be tested against each other and a job can be handed to the reference's post-processor.
No file of this kind is in the reference tree, so the key names follow the Lua writers cited above.
"""
import json
import os

from . import io
from .gas import FlowState, set_gas_model
from .sim import (Config, ExchangeBC_FullFace, FluidBlock, InFlowBC_Supersonic, OutFlowBC_FixedP, OutFlowBC_FixedPT,
                  OutFlowBC_SimpleExtrapolate, OutFlowBC_SimpleFlux, WallBC_WithSlip, WallBC_WithSlip1)

FACES = ["west", "east", "south", "north", "bottom", "top"]          # _abi face order

# options that must have these values for a job to be on the accelerated path
REQUIRED = {"viscous": False, "reacting": False, "MHD": False, "grid_motion": "none", "turbulence_model": "none",
            "udf_source_terms": False, "solver_mode": None,
            # options of the explicit update that this path does not implement: refuse them instead of ignoring them
            "with_local_time_stepping": False, "with_super_time_stepping": False, "with_super_time_stepping_flexible_stages": False,
            "residual_smoothing": False, "adjust_invalid_cell_data": False, "nsolidblock": 0, "n_ghost_cell_layers": 2}


def _flowstate_from_json(gm, d):
    massf = d.get("massf")
    kw = dict(p=d["p"], T=d["T"], velx=d.get("velx", 0.0), vely=d.get("vely", 0.0), velz=d.get("velz", 0.0))
    if massf is not None and gm.n_species > 1:
        kw["massf"] = list(massf)
    return FlowState(gm, **kw)


def _flowstate_to_json(fs, nsp):
    d = {"p": fs.gas.p, "T": fs.gas.T}
    d["massf"] = list(fs.gas.massf) if nsp > 1 else [1.0]
    d.update({"quality": 1.0, "velx": fs.vel[0], "vely": fs.vel[1], "velz": fs.vel[2], "mu_t": 0.0, "k_t": 0.0, "S": 0.0})
    return d


def _bc_from_json(gm, face, b):
    pre = b.get("pre_recon_action", [])
    post = b.get("post_conv_flux_action", [])
    types = [e["type"] for e in pre]
    if b.get("is_wall_with_viscous_effects"):
        raise ValueError(f"boundary {face}: viscous walls are not on this path")
    if not b.get("ghost_cell_data_available", True):
        if not types and not post:
            return WallBC_WithSlip1()            # bc.lua:783-806
        raise ValueError(f"boundary {face}: a boundary without ghost-cell data and with effects {types} + {[e['type'] for e in post]} is not on this path")
    if types == ["internal_copy_then_reflect"] and not post:
        return WallBC_WithSlip()
    if types == ["flowstate_copy"] and not post:
        return InFlowBC_Supersonic(_flowstate_from_json(gm, pre[0]["flowstate"]))
    if types and types[0] == "extrapolate_copy":
        if pre[0].get("x_order", 0) != 0:
            raise ValueError(f"boundary {face}: extrapolate_copy with x_order != 0 is not on this path")
        if types == ["extrapolate_copy"]:
            if [e["type"] for e in post] == ["simple_outflow_flux"]:
                return OutFlowBC_SimpleFlux()
            if not post:
                return OutFlowBC_SimpleExtrapolate()
        if types == ["extrapolate_copy", "fixed_pressure"] and not post:
            return OutFlowBC_FixedP(pre[1]["p_outside"])
        if types == ["extrapolate_copy", "fixed_pressure_temperature"] and not post:
            return OutFlowBC_FixedPT(pre[1]["p_outside"], pre[1]["T_outside"])
    if types == ["full_face_copy"] and not post:
        e = pre[0]
        return ExchangeBC_FullFace(int(e["other_block"]), FACES.index(e["other_face"]), int(e.get("orientation", 0)))
    raise ValueError(f"boundary {face}: effects {types} + {[e['type'] for e in post]} are not on this path")


def _bc_to_json(bc, nsp):
    pre, post, kind = [], [], "wall_with_slip"
    ghost = True
    if isinstance(bc, WallBC_WithSlip1):
        ghost = False
    elif isinstance(bc, WallBC_WithSlip):
        pre = [{"type": "internal_copy_then_reflect"}]
    elif isinstance(bc, InFlowBC_Supersonic):
        kind = "inflow_supersonic"
        pre = [{"type": "flowstate_copy", "flowstate": _flowstate_to_json(bc.flowState, nsp), "x0": 0.0, "y0": 0.0, "z0": 0.0, "r": 0.0}]
    elif isinstance(bc, OutFlowBC_SimpleFlux):
        kind = "outflow_simple_flux"
        pre, post = [{"type": "extrapolate_copy", "x_order": 0}], [{"type": "simple_outflow_flux"}]
    elif isinstance(bc, OutFlowBC_SimpleExtrapolate):
        kind = "outflow_simple_extrapolate"
        pre = [{"type": "extrapolate_copy", "x_order": 0}]
    elif isinstance(bc, OutFlowBC_FixedPT):
        kind = "outflow_fixed_p_and_t"
        pre = [{"type": "extrapolate_copy", "x_order": 0},
               {"type": "fixed_pressure_temperature", "p_outside": bc.p_outside, "T_outside": bc.T_outside}]
    elif isinstance(bc, OutFlowBC_FixedP):
        kind = "outflow_fixed_p"
        pre = [{"type": "extrapolate_copy", "x_order": 0}, {"type": "fixed_pressure", "p_outside": bc.p_outside}]
    elif isinstance(bc, ExchangeBC_FullFace):
        kind = "exchange_over_full_face"
        pre = [{"type": "full_face_copy", "other_block": bc.otherBlock, "other_face": FACES[bc.otherFace],
                "orientation": bc.orientation, "reorient_vector_quantities": False,
                "Rmatrix": [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]}]
    else:
        raise ValueError(f"cannot write boundary condition {type(bc).__name__}")
    return {"label": "", "type": kind, "group": "", "is_wall_with_viscous_effects": False, "ghost_cell_data_available": ghost,
            "convective_flux_computed_in_bc": bool(post), "is_design_surface": False, "num_cntrl_pts": 0,
            "pre_recon_action": pre, "post_conv_flux_action": post, "pre_spatial_deriv_action_at_bndry_faces": [],
            "pre_spatial_deriv_action_at_bndry_cells": [], "post_diff_flux_action": []}


def load_job(job_dir, job, tindx=0, **overrides):
    """Returns (config, gmodel, blocks, history_points, sim_time).  ``overrides`` set Config attributes that are
    not Eilmer options (strict_fp, ...)."""
    with open(os.path.join(job_dir, "config", f"{job}.config"), encoding="utf-8") as f:
        J = json.load(f)
    for key, want in REQUIRED.items():
        if key in J and want is not None and J[key] != want:
            raise ValueError(f"config.{key} = {J[key]!r}: not on the accelerated path (needs {want!r})")
    cfg = Config()
    for key, val in J.items():
        if hasattr(cfg, key) and not isinstance(val, (dict, list)):
            setattr(cfg, key, val)
    # the reference's writer never emits cfl_value: it writes the schedule (output.lua:173-198)
    if "cfl_schedule_values" in J:
        values, times = list(J["cfl_schedule_values"]), list(J.get("cfl_schedule_times", []))
        if len(values) != len(times) or not values:
            raise ValueError("cfl_schedule_values and cfl_schedule_times must be two non-empty lists of equal length")
        cfg.cfl_schedule = [(float(t), float(v)) for t, v in zip(times, values)]
        cfg.cfl_value = float(values[0])
    for key, val in overrides.items():
        setattr(cfg, key, val)
    gas_file = J["gas_model_file"]
    gm = set_gas_model(gas_file if os.path.isabs(gas_file) else os.path.join(job_dir, gas_file))
    dims = int(J["dimensions"])
    nfaces = 6 if dims == 3 else 4
    blocks = []
    sim_time = 0.0
    for n in range(int(J["nfluidblock"])):
        B = J[f"block_{n}"]
        if B.get("type", "fluid_block") != "fluid_block" or B.get("grid_type") != "structured_grid":
            raise ValueError(f"block {n}: only structured fluid blocks are on this path")
        if not B.get("active", True):
            raise ValueError(f"block {n}: inactive blocks are not supported")
        if float(B.get("omegaz", 0.0)) != 0.0:
            raise ValueError(f"block {n}: rotating frames (omegaz != 0) are not on this path")
        grid = io.read_grid(io.job_file(job_dir, job, "grid", n, 0))
        flow = io.read_flow(io.job_file(job_dir, job, "flow", n, tindx))
        sim_time = flow["sim_time"]
        nsp = len(getattr(gm, "species_names", None) or ["air"])
        blk = FluidBlock(io.grid_arrays(grid), io.FlowFromFile(flow, nsp=nsp), id=n)
        blk.label = B.get("label", "")
        nkc, njc, nic = flow["data"]["rho"].shape
        if (B["nic"], B["njc"], B.get("nkc", 1)) != (nic, njc, nkc):
            raise ValueError(f"block {n}: nic/njc/nkc of the config do not match the flow file")
        for face in FACES[:nfaces]:
            blk.bcList[face] = _bc_from_json(gm, face, B[f"boundary_{face}"])
        blocks.append(blk)
    hist = []
    for n in range(int(J.get("nhcell", 0))):
        ib, cell = J[f"history-cell-{n}"]
        nic, njc = J[f"block_{ib}"]["nic"], J[f"block_{ib}"]["njc"]
        hist.append((ib, cell % nic, (cell // nic) % njc, cell // (nic * njc)))
    return cfg, gm, blocks, hist, sim_time


def write_job(job_dir, job, cfg, gm, gas_model_file, blocks, sim, history_points=()):
    """Write config/<job>.config, grid/t0000, flow/t0000 and the first line of <job>.times for a set-up Simulation
    (its blocks carry geometry and the initial FlowStates are on the device)."""
    dims = cfg.dimensions
    nfaces = 6 if dims == 3 else 4
    nsp = len(getattr(gm, "species_names", None) or ["air"])
    J = {"title": cfg.title, "base_file_name": job, "grid_format": "gziptext", "flow_format": "gziptext", "new_flow_format": False,
         "gas_model_file": gas_model_file, "nfluidblock": len(blocks), "nfluidblockarrays": 0,
         "viscous": False, "reacting": False, "MHD": False, "grid_motion": "none", "turbulence_model": "none", "udf_source_terms": False,
         "n_ghost_cell_layers": 2}
    skip = {"strict_fp", "force_general_path", "force_generic_kernel", "no_tma", "no_push", "block_index", "title", "viscous", "reacting"}
    skip |= {"cfl_value", "cfl_schedule"}
    for key, val in vars(cfg).items():
        if key not in skip and isinstance(val, (bool, int, float, str)):
            J[key] = val
    # like write_config_file (output.lua:173-198): the schedule, never cfl_value
    sched = cfg.cfl_schedule or [(0.0, cfg.cfl_value)]
    J["cfl_schedule_length"] = len(sched)
    J["cfl_schedule_values"] = [float(p[1]) for p in sched]
    J["cfl_schedule_times"] = [float(p[0]) for p in sched]
    for key in ("with_local_time_stepping", "with_super_time_stepping", "residual_smoothing", "adjust_invalid_cell_data"):
        J[key] = False
    J["nsolidblock"] = 0
    for b in blocks:
        g = b.geom
        B = {"type": "fluid_block", "label": getattr(b, "label", "") or "", "active": True, "fluidBlockArrayId": -1, "omegaz": 0.0,
             "may_be_turbulent": False, "grid_type": "structured_grid", "nic": g.nic, "njc": g.njc, "nkc": g.nkc}
        for face in FACES[:nfaces]:
            B[f"boundary_{face}"] = _bc_to_json(b.bcList.get(face) or WallBC_WithSlip(), nsp)
        J[f"block_{b.id}"] = B
    J["nhcell"] = len(history_points)
    for n, (ib, i, j, k) in enumerate(history_points):
        g = next(b for b in blocks if b.id == ib).geom
        J[f"history-cell-{n}"] = [ib, i + g.nic * (j + g.njc * k)]
    os.makedirs(os.path.join(job_dir, "config"), exist_ok=True)
    with open(os.path.join(job_dir, "config", f"{job}.config"), "w", encoding="utf-8") as f:
        json.dump(J, f, indent=1)
    os.makedirs(os.path.join(job_dir, "grid", "t0000"), exist_ok=True)
    for b in blocks:
        grid = b.grid
        if not isinstance(grid, (tuple, list)):
            raise ValueError(f"block {b.id} was built from metrics only (no vertex grid): it cannot be written as a job")
        io.write_grid(io.job_file(job_dir, job, "grid", b.id, 0), *grid, label=getattr(b, "label", "") or "", dimensions=dims)
    io.write_solution_files(job_dir, job, sim, 0)


def run_job(job_dir, job, lib=None, tindx_start=0, n_solutions=None, **overrides):
    """integrate_in_time for a prepared job: load, run to config.max_time / max_step, write flow/tNNNN every
    config.dt_plot of simulated time (n_solutions, when given, spaces that many solutions evenly instead) and at
    the end, the history files and the .times entries.  A restart (tindx_start > 0) takes its time step from
    config/<job>.times like init_simulation (simcore.d:257-268).  Returns the Simulation (closed by the caller)."""
    from .sim import Simulation
    cfg, gm, blocks, hist, t0 = load_job(job_dir, job, tindx_start, **overrides)
    sim = Simulation(cfg, gm, blocks, lib=lib) if lib is not None else Simulation(cfg, gm, blocks)
    sim.time = t0
    if tindx_start > 0:
        try:
            sim.dt_global = io.read_times(job_dir, job)[tindx_start][1]
        except (OSError, KeyError):
            pass
    keys = [sim.set_history_point(*h) for h in hist]
    t_end = cfg.max_time
    dt_plot = (t_end - t0) / max(1, n_solutions) if n_solutions else float(getattr(cfg, "dt_plot", t_end - t0))
    dt_plot = max(dt_plot, 1.0e-300)
    n = 0
    while sim.time < t_end and sim.step < cfg.max_step:
        n += 1
        sim.run(max_time=min(t_end, t0 + n * dt_plot))
        io.write_solution_files(job_dir, job, sim, tindx_start + n)
        if n_solutions and n >= n_solutions:
            break
    os.makedirs(os.path.join(job_dir, "hist"), exist_ok=True)
    for key in keys:
        ib, i, j, k = key
        g = next(b for b in blocks if b.id == ib).geom
        cell = i + g.nic * (j + g.njc * k)
        io.write_history_file(os.path.join(job_dir, "hist", f"{job}-blk-{ib}-cell-{cell}.dat.0"), sim, key)
    return sim

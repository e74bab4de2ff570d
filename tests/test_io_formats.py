"""Eilmer's grid and flow file formats (gdtk_b200/io.py) against the reference's own sample output
(tests/golden/ref_sample_data = src/eilmer/sample-data of the reference): the files pin the readers,
the cell geometry and the ideal-gas state; the job they describe (cone20 as the reference prepared
it, block 1 on its area-orthogonality grid) is then run through the oracle."""
import os

import numpy as np

from gdtk_b200 import Simulation, cases, io
from gdtk_b200.gas import set_gas_model
from gdtk_b200.geometry import NG, geometry_2d
from gdtk_b200.sim import FluidBlock, InFlowBC_Supersonic, OutFlowBC_Simple, identify_block_connections

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_sample_data")


def sample(kind, blk):
    return os.path.join(DATA, f"cone20.{kind}.b{blk:04d}.t0000.gz")


def test_sample_grids_are_read():
    g0, g1 = io.read_grid(sample("grid", 0)), io.read_grid(sample("grid", 1))
    assert g0["X"].shape == (1, 41, 11) and g1["X"].shape == (1, 41, 31) and g0["dimensions"] == 2
    assert g0["X"][0, 0, -1] == 0.2 and g1["X"][0, 0, 0] == 0.2          # the blocks meet at x = 0.2
    assert abs(g1["Y"][0, 0, -1] - 0.29118) < 1e-12 and g1["X"][0, 0, -1] == 1.0    # end of the cone surface
    assert np.all(g1["Y"][0, -1, :] == 1.0)


def test_cell_centres_and_axisymmetric_volumes_match_the_reference_output():
    """pos.x, pos.y and volume in the flow files were computed by the reference from the grid files
    (compute_primary_cell_geometric_data).  Both files carry 13 digits, and the reference worked from
    the unrounded vertices: differences of coordinates rounded at 5e-13 over cell sizes of 0.02 leave
    about 1e-10 of the cell volume."""
    for blk in (0, 1):
        X, Y = io.grid_arrays(io.read_grid(sample("grid", blk)))
        geom = geometry_2d(X, Y, True)
        f = io.read_flow(sample("flow", blk))["data"]
        sl = (0, slice(NG, NG + geom.njc), slice(NG, NG + geom.nic))
        for mine, ref in ((geom.pos[0][sl], f["pos.x"][0]), (geom.pos[1][sl], f["pos.y"][0]), (geom.vol[sl], f["volume"][0])):
            assert np.max(np.abs(mine - ref) / np.abs(ref)) < 1.0e-10


def test_ideal_gas_relations_hold_in_the_reference_output():
    """The sample flow files were written with an earlier ideal-air file (R = 8.31451/0.028964, not the
    0.02896 of the lua file next to them), so only the relations between the columns are checked:
    a^2 = gamma p / rho, e = p / (rho (gamma - 1)), R = p / (rho T) constant."""
    gm = set_gas_model(os.path.join(DATA, "ideal-air-gas-model.lua"))
    for blk in (0, 1):
        f = io.read_flow(sample("flow", blk))["data"]
        p, T, rho = f["p"], f["T[0]"], f["rho"]
        assert np.max(np.abs(np.sqrt(gm.gamma * p / rho) - f["a"]) / f["a"]) < 2e-12
        assert np.max(np.abs(p / (rho * (gm.gamma - 1.0)) - f["e[0]"]) / f["e[0]"]) < 2e-12
        R = p / (rho * T)
        assert abs(R.max() - R.min()) / R.max() < 2e-12 and abs(R.mean() - 8.31451 / 0.028964) < 1e-6


def test_grid_and_flow_round_trip(tmp_path, oracle):
    cfg, gm, blocks = cases.sod(dims=3, ncells=12, nj=3, nk=2, nblocks=2)
    sim = Simulation(cfg, gm, blocks, lib=oracle)
    sim.run(max_step=5, max_time=1.0)
    X, Y, Z = blocks[1].grid
    io.write_grid(tmp_path / "g.gz", X, Y, Z, label="blk-1", tags=["a", "b", "c", "d", "e", "f"])
    g = io.read_grid(tmp_path / "g.gz")
    assert g["label"] == "blk-1" and g["tags"] == ["a", "b", "c", "d", "e", "f"] and g["dimensions"] == 3
    assert np.array_equal(g["X"], X) and np.array_equal(g["Y"], Y) and np.array_equal(g["Z"], Z)
    io.write_flow(tmp_path / "f.gz", sim, 1, sim.time, label="blk-1")
    f = io.read_flow(tmp_path / "f.gz")
    assert f["sim_time"] == sim.time and f["names"] == io.flow_variable_list(gm) and f["dimensions"] == 3
    P = [sim.interior(1, a) for a in sim.download_flow(1)]
    for name, q in (("rho", 0), ("u", 1), ("p", 2), ("T", 3), ("a", 4), ("vel.x", 5)):
        assert np.array_equal(f["data"][name], P[q])                # "%.18e" is lossless for doubles
    # and back into a simulation
    blk = FluidBlock(blocks[1].grid, io.FlowFromFile(f), id=0)
    s2 = Simulation(cfg, gm, [blk], lib=oracle)
    assert np.array_equal(s2.interior(0, s2.download_flow(0)[0]), P[0])
    sim.close(); s2.close()


def test_cone20_from_the_reference_files(oracle):
    """The job as the reference prepared it: its grids (block 1 is the area-orthogonality grid that
    cases.cone20 replaces by a Coons patch) and its initial flow files, the default flux calculator.
    cone20-test.rb expects 833 +- 3 steps."""
    gm = set_gas_model(os.path.join(DATA, "ideal-air-gas-model.lua"))
    cfg, _, _ = cases.cone20(flux_calculator="adaptive_hanel_ausmdv")
    blocks = [FluidBlock(io.grid_arrays(io.read_grid(sample("grid", b))), io.FlowFromFile(io.read_flow(sample("flow", b))), id=b)
              for b in (0, 1)]
    from gdtk_b200.gas import FlowState
    blocks[0].bcList["west"] = InFlowBC_Supersonic(FlowState(gm, p=95.84e3, T=1103.0, velx=1000.0))
    blocks[1].bcList["east"] = OutFlowBC_Simple()
    identify_block_connections(blocks, 2)
    sim = Simulation(cfg, gm, blocks, lib=oracle)
    steps = sim.run()
    assert abs(steps - 833) < 3
    sim.close()


def test_history_points(tmp_path, oracle):
    """setHistoryPoint of cone20.lua (ib=1, i=20, j=0 and ib=0, i=5, j=5): samples every dt_history, probe equals the
    downloaded block, and the history file has one line per sample in the flow-file layout."""
    cfg, gm, blocks = cases.cone20(dt_history=1.0e-4)
    sim = Simulation(cfg, gm, blocks, lib=oracle)
    k1, k0 = sim.set_history_point(1, 20, 0), sim.set_history_point(0, 5, 5)
    sim.run(max_step=200, max_time=1.0)
    n = len(sim.history[k1])
    assert n == int(sim.time / 1.0e-4) and n == len(sim.history[k0]) and n >= 5
    times = [r[0] for r in sim.history[k1]]
    assert all(t2 > t1 for t1, t2 in zip(times, times[1:])) and times[0] >= 1.0e-4
    P = sim.download_flow(1)
    probe = sim.probe_cells([k1])[0]
    assert all(probe[v] == sim.interior(1, P[v])[0, 0, 20] for v in range(8))
    io.write_history_file(tmp_path / "cone20-blk-1-cell-20.dat.0", sim, k1)
    lines = open(tmp_path / "cone20-blk-1-cell-20.dat.0").read().splitlines()
    assert lines[0].startswith("# 1:t 2:pos.x 3:pos.y") and len(lines) == n + 1
    assert len(lines[1].split()) == 1 + len(io.flow_variable_list(gm))
    sim.close()


def test_solution_directory_layout(tmp_path, oracle):
    """flow/tNNNN/<job>.flow.bBBBB.tNNNN.gz + config/<job>.times, the layout e4shared --post reads."""
    cfg, gm, blocks = cases.cone20(nx0=6, nx1=14, ny=16)
    sim = Simulation(cfg, gm, blocks, lib=oracle)
    io.write_solution_files(tmp_path, "cone20", sim, 0)
    sim.run(max_step=20, max_time=1.0)
    io.write_solution_files(tmp_path, "cone20", sim, 1)
    times = io.read_times(tmp_path, "cone20")
    assert times[0] == (0.0, sim.config.dt_init) and times[1] == (sim.time, sim.dt_global)
    for tindx in (0, 1):
        for b in (0, 1):
            f = io.read_flow(io.job_file(tmp_path, "cone20", "flow", b, tindx))
            assert f["sim_time"] == times[tindx][0] and f["data"]["rho"].shape == (1, 16, 6 if b == 0 else 14)
    rho = sim.interior(1, sim.download_flow(1)[0])
    assert np.array_equal(io.read_flow(io.job_file(tmp_path, "cone20", "flow", 1, 1))["data"]["rho"], rho)
    sim.close()


def test_prepared_job_round_trip(tmp_path, oracle):
    """write_job (what the preparation stage leaves: config JSON in the layout of output.lua / fluidblock.lua / bc.lua,
    grid and flow files, .times) and load_job / run_job: the job read back from disk runs exactly like the one
    built in Python, and writes its solutions and history files."""
    import shutil
    from gdtk_b200 import job as jobmod
    from gdtk_b200 import OutFlowBC_FixedP
    cfg, gm, blocks = cases.cone20(flux_calculator="adaptive", nx0=6, nx1=14, ny=16, max_step=40, dt_history=2.0e-5)
    blocks[1].bcList["north"] = OutFlowBC_FixedP(5955.0)
    sim = Simulation(cfg, gm, blocks, lib=oracle)
    shutil.copy(os.path.join(DATA, "ideal-air-gas-model.lua"), tmp_path / "ideal-air-gas-model.lua")
    jobmod.write_job(tmp_path, "cone20", cfg, gm, "ideal-air-gas-model.lua", blocks, sim, history_points=[(1, 7, 0, 0)])
    key = sim.set_history_point(1, 7, 0, 0)
    sim.run()
    U = sim.interior(1, sim.download_conserved(1)[0]).copy()

    cfg2, gm2, blocks2, hist, t0 = jobmod.load_job(tmp_path, "cone20")
    assert t0 == 0.0 and hist == [(1, 7, 0, 0)] and cfg2.flux_calculator == "adaptive" and cfg2.axisymmetric is True
    assert [type(blocks2[1].bcList[f]).__name__ for f in ("west", "east", "south", "north")] == \
        ["ExchangeBC_FullFace", "OutFlowBC_SimpleFlux", "WallBC_WithSlip", "OutFlowBC_FixedP"]
    assert type(blocks2[0].bcList["west"]).__name__ == "InFlowBC_Supersonic"
    s2 = jobmod.run_job(tmp_path, "cone20", lib=oracle)
    assert s2.step == sim.step == 40 and s2.dt_history == sim.dt_history
    assert np.array_equal(s2.interior(1, s2.download_conserved(1)[0]), U)
    assert np.array_equal(np.array(s2.history[key]), np.array(sim.history[key]))
    times = io.read_times(tmp_path, "cone20")
    assert sorted(times) == [0, 1] and times[1][0] == s2.time
    assert os.path.exists(tmp_path / "hist" / "cone20-blk-1-cell-7.dat.0")
    f = io.read_flow(io.job_file(tmp_path, "cone20", "flow", 1, 1))
    assert np.array_equal(f["data"]["rho"], s2.interior(1, s2.download_flow(1)[0]))
    sim.close(); s2.close()

    # a viscous job is refused
    import json
    path = tmp_path / "config" / "cone20.config"
    J = json.load(open(path)); J["viscous"] = True; json.dump(J, open(path, "w"))
    try:
        jobmod.load_job(tmp_path, "cone20")
        assert False, "viscous job was accepted"
    except ValueError:
        pass

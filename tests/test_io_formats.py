"""Eilmer's grid and flow file formats (gdtk_b200/io.py, job.py) and the reference's own sample output
(src/eilmer/sample-data of the reference: the cone20 grids and initial flow files its preparation stage wrote).
tests/golden/ref_sample_cone20.npz holds the arrays of those files (tests/golden/make_ref_sample_fixtures.py);
they pin the cell geometry and the relations of the ideal-gas state, and the job they describe (block 1 on its
area-orthogonality grid) is run through the oracle.  Where the reference tree itself is present (the build
container) the readers are also run on the original files."""
import os

import numpy as np
import pytest

from gdtk_b200 import Simulation, cases, io
from gdtk_b200.gas import FlowState, set_gas_model
from gdtk_b200.geometry import NG, geometry_2d
from gdtk_b200.sim import FluidBlock, InFlowBC_Supersonic, OutFlowBC_Simple, identify_block_connections

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
GAS = os.path.join(GOLD, "gas", "ideal-air-gas-model.json")
REF_SAMPLES = "/root/reference/src/eilmer/sample-data"
SAMPLE = dict(np.load(os.path.join(GOLD, "ref_sample_cone20.npz")))


def sample_flow(b):
    names = [str(n) for n in SAMPLE[f"b{b}_names"]]
    return {"sim_time": float(SAMPLE[f"b{b}_sim_time"]), "names": names, "label": "", "dimensions": 2,
            "data": {n: SAMPLE[f"b{b}_{n}"] for n in names}}


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="the reference tree is not on this machine")
def test_readers_on_the_reference_files():
    for b in (0, 1):
        g = io.read_grid(os.path.join(REF_SAMPLES, f"cone20.grid.b{b:04d}.t0000.gz"))
        assert g["dimensions"] == 2 and np.array_equal(g["X"][0], SAMPLE[f"b{b}_X"]) and np.array_equal(g["Y"][0], SAMPLE[f"b{b}_Y"])
        f = io.read_flow(os.path.join(REF_SAMPLES, f"cone20.flow.b{b:04d}.t0000.gz"))
        assert f["names"][:5] == ["pos.x", "pos.y", "pos.z", "volume", "rho"] and f["sim_time"] == 0.0
        assert all(np.array_equal(f["data"][n], SAMPLE[f"b{b}_{n}"]) for n in f["names"])


def test_sample_grids():
    X0, X1, Y1 = SAMPLE["b0_X"], SAMPLE["b1_X"], SAMPLE["b1_Y"]
    assert X0.shape == (41, 11) and X1.shape == (41, 31)
    assert X0[0, -1] == 0.2 and X1[0, 0] == 0.2                         # the blocks meet at x = 0.2
    assert abs(Y1[0, -1] - 0.29118) < 1e-12 and X1[0, -1] == 1.0         # end of the cone surface
    assert np.all(Y1[-1, :] == 1.0)


def test_cell_centres_and_axisymmetric_volumes_match_the_reference_output():
    """pos.x, pos.y and volume in the flow files were computed by the reference from the grid files
    (compute_primary_cell_geometric_data).  Both files carry 13 digits, and the reference worked from
    the unrounded vertices: differences of coordinates rounded at 5e-13 over cell sizes of 0.02 leave
    about 1e-10 of the cell volume."""
    for blk in (0, 1):
        geom = geometry_2d(SAMPLE[f"b{blk}_X"], SAMPLE[f"b{blk}_Y"], True)
        sl = (0, slice(NG, NG + geom.njc), slice(NG, NG + geom.nic))
        for mine, ref in ((geom.pos[0][sl], SAMPLE[f"b{blk}_pos.x"][0]), (geom.pos[1][sl], SAMPLE[f"b{blk}_pos.y"][0]),
                          (geom.vol[sl], SAMPLE[f"b{blk}_volume"][0])):
            assert np.max(np.abs(mine - ref) / np.abs(ref)) < 1.0e-10


def test_ideal_gas_relations_hold_in_the_reference_output():
    """The sample flow files were written with an earlier ideal-air file (R = 8.31451/0.028964, not the
    0.02896 of today's file), so only the relations between the columns are checked:
    a^2 = gamma p / rho, e = p / (rho (gamma - 1)), R = p / (rho T) constant."""
    gm = set_gas_model(GAS)
    for blk in (0, 1):
        p, T, rho = SAMPLE[f"b{blk}_p"], SAMPLE[f"b{blk}_T[0]"], SAMPLE[f"b{blk}_rho"]
        assert np.max(np.abs(np.sqrt(gm.gamma * p / rho) - SAMPLE[f"b{blk}_a"]) / SAMPLE[f"b{blk}_a"]) < 2e-12
        assert np.max(np.abs(p / (rho * (gm.gamma - 1.0)) - SAMPLE[f"b{blk}_e[0]"]) / SAMPLE[f"b{blk}_e[0]"]) < 2e-12
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


def test_cone20_on_the_reference_grid(oracle):
    """The job as the reference prepared it: its grids (block 1 is the area-orthogonality grid that
    cases.cone20 replaces by a Coons patch) and its initial flow, the default flux calculator.
    cone20-test.rb expects 833 +- 3 steps."""
    gm = set_gas_model(GAS)
    cfg, _, _ = cases.cone20(flux_calculator="adaptive_hanel_ausmdv")
    blocks = [FluidBlock((SAMPLE[f"b{b}_X"], SAMPLE[f"b{b}_Y"]), io.FlowFromFile(sample_flow(b)), id=b) for b in (0, 1)]
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
    shutil.copy(GAS, tmp_path / "ideal-air-gas-model.json")
    jobmod.write_job(tmp_path, "cone20", cfg, gm, "ideal-air-gas-model.json", blocks, sim, history_points=[(1, 7, 0, 0)])
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


def test_prepared_job_round_trip_thermally_perfect(tmp_path, oracle):
    """Five species: the flow files carry massf[i]-<name> columns and dt_chem, inflow FlowStates their mass fractions."""
    import shutil
    from gdtk_b200 import job as jobmod
    from gdtk_b200.grids import box_grid_3d, split_grid
    from gdtk_b200.sim import Config
    gas_file = os.path.join(GOLD, "gas", "therm-perf-5-species-air.json")
    gm = set_gas_model(gas_file)
    cfg = Config(dimensions=3, flux_calculator="hanel", max_step=3, max_time=1.0, dt_init=1.0e-7)
    inflow = FlowState(gm, p=95.84e3, T=3000.0, velx=3000.0, massf={"N2": 0.76, "O2": 0.23, "NO": 0.004, "N": 0.003, "O": 0.003})
    still = FlowState(gm, p=2.0e4, T=2000.0, massf={"N2": 0.767, "O2": 0.233})
    blocks = [FluidBlock(sub, inflow if ib == 0 else still, id=ib)
              for ib, jb, kb, sub in split_grid(box_grid_3d((0, 0, 0), (1.0, 0.5, 0.25), 8, 4, 2), 2, 1, 1)]
    blocks[0].bcList["west"] = InFlowBC_Supersonic(inflow)
    blocks[1].bcList["east"] = OutFlowBC_Simple()
    identify_block_connections(blocks, 3)
    sim = Simulation(cfg, gm, blocks, lib=oracle)
    shutil.copy(gas_file, tmp_path / "tpg.json")
    jobmod.write_job(tmp_path, "box", cfg, gm, "tpg.json", blocks, sim)
    sim.run()
    U = {b.id: sim.interior(b.id, sim.download_conserved(b.id)[5]).copy() for b in blocks}
    f = io.read_flow(io.job_file(tmp_path, "box", "flow", 0, 0))
    assert [n for n in f["names"] if n.startswith("massf[")] == [f"massf[{i}]-{s}" for i, s in enumerate(gm.species_names)]
    assert "dt_chem" in f["names"]
    s2 = jobmod.run_job(tmp_path, "box", lib=oracle)
    assert s2.step == 3 and s2.dt_history == sim.dt_history
    # (the temperature of a thermally perfect gas comes out of a Newton iteration that stops at 1e-6 K, started
    #  from the temperature in the file: a restart is close, not bit-identical -- in the reference too)
    for b in U:
        a = s2.interior(b, s2.download_conserved(b)[5])
        assert np.max(np.abs(a - U[b])) <= 1.0e-9 * np.max(np.abs(U[b]))
    sim.close(); s2.close()

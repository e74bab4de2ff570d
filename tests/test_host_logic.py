"""Host-side logic that sits either side of the accelerated path: block connections, the
full-face cell mapping, the time-step policy and the step-retry policy (run against the oracle
or a stub of the C ABI; no GPU)."""
import ctypes as C

import numpy as np
import pytest

from gdtk_b200 import _abi, Simulation, cases, FluidBlock, FlowState, Config, identify_block_connections
from gdtk_b200.sim import full_face_source, ExchangeBC_FullFace
from gdtk_b200.grids import box_grid_2d, box_grid_3d, split_grid, connect_block_array, quad_patch_grid
from gdtk_b200.distributed import distribute_blocks, octant_owner


def test_identify_block_connections_2d_all_senses():
    gm = cases.ideal_air()
    fs = FlowState(gm, p=1e5, T=300.0)
    # A: [0,1]x[0,1]; B sits east of A but is built rotated so that its NORTH face touches A's east
    xa, ya = box_grid_2d(0, 1, 0, 1, 4, 3)
    A = FluidBlock((xa, ya), fs, id=0)
    # rotated block: i runs along +y... use a patch with corners chosen so that north edge = x=1 line
    xb, yb = quad_patch_grid((2.0, 0.0), (2.0, 1.0), (1.0, 1.0), (1.0, 0.0), 3, 5)
    B = FluidBlock((xb, yb), fs, id=1)
    identify_block_connections([A, B], 2)
    assert isinstance(A.bcList["east"], ExchangeBC_FullFace)
    assert A.bcList["east"].otherBlock == 1 and A.bcList["east"].otherFace == _abi.NORTH
    assert B.bcList["north"].otherBlock == 0 and B.bcList["north"].otherFace == _abi.EAST


def test_full_face_source_matches_reference_table():
    # reference: full_face_copy.d:704-870, this=east other=west: i_src = 0+n, j_src = j
    assert full_face_source(2, _abi.EAST, (7, 5, 1), _abi.WEST, 3, 0, 1) == (1, 3, 0)
    # this=north other=north: j_src = njc-1-n, i_src = nic-i-1
    assert full_face_source(2, _abi.NORTH, (7, 5, 1), _abi.NORTH, 2, 0, 0) == (4, 4, 0)
    # this=west other=south: j_src = n, i_src = j
    assert full_face_source(2, _abi.WEST, (7, 5, 1), _abi.SOUTH, 2, 0, 1) == (2, 1, 0)
    # 3D aligned top <- bottom
    assert full_face_source(3, _abi.TOP, (4, 5, 6), _abi.BOTTOM, 1, 2, 1) == (1, 2, 1)
    with pytest.raises(ValueError):
        full_face_source(3, _abi.TOP, (4, 5, 6), _abi.EAST, 0, 0, 0)


def test_two_d_rotated_connection_runs_and_matches_unrotated(oracle):
    """The same physical domain meshed as [A | B] with B stored rotated by 90 degrees must give
    the same answer as with B stored aligned (exercises the reversed/transposed cell mapping)."""
    gm = cases.ideal_air()
    hi = FlowState(gm, p=1.0e5, T=348.4)
    lo = FlowState(gm, p=1.0e4, T=278.8)
    cfg = Config(dimensions=2, flux_calculator="ausmdv", dt_init=1e-6)
    xa, ya = box_grid_2d(0, 0.5, 0, 0.1, 20, 4)
    res = []
    for rotated in (False, True):
        A = FluidBlock((xa, ya), hi, id=0)
        if rotated:
            # B's i-direction runs along +y, j along -x: its NORTH... build via patch corners
            xb, yb = quad_patch_grid((1.0, 0.0), (1.0, 0.1), (0.5, 0.1), (0.5, 0.0), 4, 20)
        else:
            xb, yb = box_grid_2d(0.5, 1.0, 0, 0.1, 20, 4)
        B = FluidBlock((xb, yb), lo, id=1)
        identify_block_connections([A, B], 2)
        assert "east" in A.bcList
        sim = Simulation(cfg, gm, [A, B], lib=oracle)
        sim.run(max_step=40, max_time=1.0)
        rho = sim.interior(0, sim.download_flow(0)[0]).copy()
        res.append(rho)
        sim.close()
    assert np.allclose(res[0], res[1], rtol=1e-12, atol=0)


def test_block_distribution():
    cfg, gm, blocks = cases.box3d(n=16, nb=4)
    for w in (1, 2, 4, 8):
        own = octant_owner({v: next(b for b in blocks if b.id == k) for k, v in cfg.block_index.items()}, 4, w)
        counts = np.bincount(list(own.values()), minlength=w)
        assert counts.tolist() == [64 // w] * w
    own = distribute_blocks(blocks, 3, mode="load-balance")
    assert sorted(np.bincount(list(own.values())).tolist()) == [21, 21, 22]


class _StubLib:
    """Just enough of the C ABI to drive Simulation's dt / retry policy."""
    prefix = "stub_"

    def __init__(self, dt_allow_seq, fail_steps=()):
        self.seq, self.fails, self.nstep = list(dt_allow_seq), list(fail_steps), 0
        self.calls = []

    def check(self, rc, what):
        return rc

    def error(self):
        return "stub"

    def __getattr__(self, name):
        def f(*a):
            if name == "compute_dt":
                out = a[-1]
                out[0], out[1], out[2] = self.seq.pop(0), 0.4, 0.0
                return 0
            if name == "step":
                self.calls.append(a[2])
                if self.fails and self.fails[0]:
                    self.fails.pop(0)
                    return 1
                if self.fails:
                    self.fails.pop(0)
                return 0
            return 0
        return f


def test_time_step_policy_and_retry():
    cfg, gm, blocks = cases.sod(dims=2, ncells=8, nj=2)
    cfg.dt_init, cfg.dt_max, cfg.cfl_count = 1.0e-3, 2.5e-6, 2
    stub = _StubLib([4.0e-6, 1.0e-6, 8.0e-6, 8.0e-6], fail_steps=[0, 0, 1, 1, 0])
    sim = Simulation.__new__(Simulation)
    sim.config, sim.lib, sim.handle = cfg, stub, 0
    sim.time, sim.step, sim.dt_global, sim.dt_history, sim.cfl_max = 0.0, 0, cfg.dt_init, [], 0.0
    sim.determine_time_step_size()                 # step 0: min(dt_init, dt_allow)
    assert sim.dt_global == 4.0e-6
    sim.gasdynamic_step()
    sim.determine_time_step_size()                 # step 1: not a check step (cfl_count = 2)
    assert sim.dt_global == 4.0e-6
    sim.gasdynamic_step()
    sim.determine_time_step_size()                 # step 2: shrink immediately
    assert sim.dt_global == 1.0e-6
    sim.gasdynamic_step()                          # fails twice -> dt * 0.2 * 0.2, third attempt succeeds
    assert stub.calls[-3:] == [1.0e-6, 1.0e-6 * 0.2, 1.0e-6 * 0.2 * 0.2]
    assert sim.step == 3
    sim.step = 4
    sim.dt_global = 1.0e-6
    sim.determine_time_step_size()                 # grow by at most 1.5x, capped by dt_max
    assert sim.dt_global == 1.5e-6
    sim.step = 6
    sim.dt_global = 2.0e-6
    sim.determine_time_step_size()
    assert sim.dt_global == 2.5e-6                 # min(3e-6, dt_allow=8e-6, dt_max)


def test_retry_gives_up_after_max_attempts():
    cfg, gm, blocks = cases.sod(dims=2, ncells=8, nj=2)
    stub = _StubLib([], fail_steps=[1, 1, 1, 1])
    sim = Simulation.__new__(Simulation)
    sim.config, sim.lib, sim.handle = cfg, stub, 0
    sim.time, sim.step, sim.dt_global, sim.dt_history = 0.0, 0, 1.0e-6, []
    with pytest.raises(RuntimeError, match="after 3 attempts"):
        sim.gasdynamic_step()
    assert len(stub.calls) == 3


def test_config_options_added_for_the_wider_path():
    """Names and values are Eilmer's: flux calculators incl. "adaptive" (= adaptive_efm_ausmdv), the four
    thermo interpolators, and the boundary conditions with parameters."""
    from gdtk_b200 import _abi, cases
    from gdtk_b200.sim import Config, OutFlowBC_FixedP, OutFlowBC_FixedPT
    gm = cases.ideal_air()
    assert Config(flux_calculator="adaptive").to_struct(gm).flux_calculator == _abi.FLUX_CALCULATORS["adaptive_efm_ausmdv"] == 10
    assert Config(flux_calculator="efm").to_struct(gm).flux_calculator == 9
    for name, code in (("rhou", 0), ("pt", 1), ("rhop", 2), ("rhot", 3)):
        assert Config(flux_calculator="ausmdv", thermo_interpolator=name).to_struct(gm).thermo_interpolator == code
    try:
        Config(flux_calculator="ausmdv", thermo_interpolator="rhoe").to_struct(gm)
        assert False
    except ValueError:
        pass
    assert OutFlowBC_FixedP(2.0e4).params() == [2.0e4] and OutFlowBC_FixedP.kind == 5
    assert OutFlowBC_FixedPT(2.0e4, 300).params() == [2.0e4, 300.0] and OutFlowBC_FixedPT.kind == 6


def test_cfl_schedule_interpolation():
    """Schedule.interpolate_value (src/nm/schedule.d:39-56) times cfl_scale_factor (simcore_gasdynamic_step.d:77-78)."""
    from gdtk_b200.sim import Config
    c = Config(flux_calculator="ausmdv", cfl_value=0.4)
    assert c.cfl_at(0.0) == 0.4 and c.cfl_at(1.0) == 0.4
    c.cfl_schedule = [(0.0, 0.1), (1.0e-3, 0.5), (3.0e-3, 0.9)]
    assert c.cfl_at(-1.0) == 0.1 and c.cfl_at(0.0) == 0.1
    assert c.cfl_at(0.5e-3) == pytest.approx(0.3, rel=1e-14)
    assert c.cfl_at(2.0e-3) == pytest.approx(0.7, rel=1e-14)
    assert c.cfl_at(3.0e-3) == 0.9 and c.cfl_at(1.0) == 0.9
    c.cfl_scale_factor = 0.5
    assert c.cfl_at(1.0) == 0.45


def test_step_status_is_decided_collectively():
    """A rank whose own step succeeded while another rank's failed takes its step back (eb200_undo_step), scales
    dt and retries with everybody else; a fatal error elsewhere ends the run here too."""
    cfg, gm, blocks = cases.sod(dims=2, ncells=8, nj=2)

    class Lib(_StubLib):
        def __init__(self):
            super().__init__([], fail_steps=[0, 0, 0])
            self.undone = 0

        def __getattr__(self, name):
            if name == "undo_step":
                def f(*a):
                    self.undone += 1
                    return 0
                return f
            return super().__getattr__(name)

    lib = Lib()
    sim = Simulation.__new__(Simulation)
    sim.config, sim.lib, sim.handle = cfg, lib, 0
    sim.time, sim.step, sim.dt_global, sim.dt_history = 0.0, 0, 1.0e-6, []
    others = [1, 0]                       # the other rank fails the first attempt, then succeeds
    sim.reduce_step_status = lambda rc: max(rc, others.pop(0))
    sim.gasdynamic_step()
    assert lib.undone == 1 and lib.calls == [1.0e-6, 1.0e-6 * 0.2] and sim.step == 1
    sim.reduce_step_status = lambda rc: -2
    with pytest.raises(RuntimeError, match="another rank"):
        sim.gasdynamic_step()


def test_wall_without_ghost_cells_in_the_job_file():
    """bc.lua:783-806 writes WallBC_WithSlip1 as type wall_with_slip with ghost_cell_data_available = false and no
    effects; the job reader maps that (and only that) to the one-sided path."""
    from gdtk_b200 import job
    from gdtk_b200.sim import WallBC_WithSlip, WallBC_WithSlip1
    j = job._bc_to_json(WallBC_WithSlip1(), 1)
    assert j["type"] == "wall_with_slip" and j["ghost_cell_data_available"] is False and j["pre_recon_action"] == []
    assert isinstance(job._bc_from_json(None, "north", j), WallBC_WithSlip1)
    j2 = job._bc_to_json(WallBC_WithSlip(), 1)
    assert j2["ghost_cell_data_available"] is True
    assert type(job._bc_from_json(None, "north", j2)) is WallBC_WithSlip
    j["pre_recon_action"] = [{"type": "internal_copy_then_reflect"}]
    with pytest.raises(ValueError):
        job._bc_from_json(None, "north", j)

"""Block connections in 3D, every way round (the reference's examples/eilmer/3D/connection-test): two blocks
side by side whose logical axes are rotated through all 24 right-handed arrangements, so that any face of the
one can meet any face of the other in any of the four rotations.  The blocks stay where they are in space, the
initial flow is a function of position without symmetry across the common face, so after a few steps the flow
in physical space must be what the plainly aligned pair gives (to round-off: the order in which a cell sums its
face fluxes follows its logical axes)."""
import itertools
import math

import numpy as np
import pytest

from gdtk_b200 import Simulation, cases
from gdtk_b200.gas import FlowState
from gdtk_b200.grids import box_grid_3d
from gdtk_b200.sim import Config, ExchangeBC_FullFace, FluidBlock, identify_block_connections


def rotations():
    """The 24 right-handed relabellings of (i, j, k): new axis m runs along old axis perm[m] in direction sign[m]."""
    out = []
    for perm in itertools.permutations(range(3)):
        parity = 1 if perm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)) else -1
        for sign in itertools.product((1, -1), repeat=3):
            if parity * sign[0] * sign[1] * sign[2] == 1:
                out.append((perm, sign))
    return out


def relabel(grid, rot):
    perm, sign = rot
    out = []
    for a in grid:
        q = np.transpose(np.asarray(a).transpose(2, 1, 0), perm)        # [i, j, k] axes, relabelled
        for m in range(3):
            if sign[m] < 0:
                q = np.flip(q, axis=m)
        out.append(np.ascontiguousarray(q.transpose(2, 1, 0)))
    return tuple(out)


def make_case(gm, rot_a, rot_b):
    def init(x, y, z):
        s = math.sin(1.3 * x + 0.7 * y + 2.1 * z)
        c = math.cos(0.9 * x - 1.7 * y + 0.5 * z)
        return FlowState(gm, p=1.0e5 * (1.0 + 0.1 * s), T=300.0 * (1.0 + 0.05 * c), velx=100.0 * (1 + 0.2 * c),
                         vely=50.0 * s, velz=-30.0 * (s + c))
    ga = relabel(box_grid_3d((0.0, 0.0, 0.0), (1.0, 0.9, 0.8), 4, 3, 2), rot_a)
    gb = relabel(box_grid_3d((1.0, 0.0, 0.0), (2.0, 0.9, 0.8), 5, 3, 2), rot_b)
    cfg = Config(dimensions=3, flux_calculator="ausmdv", max_step=3, max_time=1.0, dt_init=1.0e-5, cfl_value=0.5)
    blocks = [FluidBlock(ga, init, id=0), FluidBlock(gb, init, id=1)]
    identify_block_connections(blocks, 3)
    return cfg, blocks


def physical_solution(sim, blocks):
    from gdtk_b200.geometry import NG
    out = {}
    for b in blocks:
        g = b.geom
        sl = (slice(g.kg, g.kg + g.nkc), slice(NG, NG + g.njc), slice(NG, NG + g.nic))
        P = [sim.interior(b.id, a) for a in sim.download_flow(b.id)]
        X, Y, Z = (g.pos[m][sl] for m in range(3))
        for idx in np.ndindex(X.shape):
            key = (round(float(X[idx]), 9), round(float(Y[idx]), 9), round(float(Z[idx]), 9))
            out[key] = np.array([p[idx] for p in P])
    return out


@pytest.fixture(scope="module")
def reference_solution(oracle):
    gm = cases.ideal_air()
    ident = ((0, 1, 2), (1, 1, 1))
    cfg, blocks = make_case(gm, ident, ident)
    assert blocks[0].bcList["east"].cell_map is None          # the aligned pair needs no map
    sim = Simulation(cfg, gm, blocks, lib=oracle)
    sim.run()
    sol = physical_solution(sim, blocks)
    sim.close()
    return sol


def compare(sol, ref):
    assert sol.keys() == ref.keys()
    scale = np.max(np.abs(np.array(list(ref.values()))), axis=0)
    scale[5:8] = scale[5:8].max()
    worst = max(float(np.max(np.abs(sol[k] - ref[k]) / scale)) for k in ref)
    assert worst < 1.0e-12, worst


# block A shows each of its six logical faces to block B once; block B goes through all 24 arrangements
ROT_A = [r for r in rotations() if r[1] == (1, 1, 1) or r[0] == (0, 1, 2)][:6]


def test_all_24_arrangements_of_one_block(oracle, reference_solution):
    gm = cases.ideal_air()
    seen = set()
    for rot_b in rotations():
        cfg, blocks = make_case(gm, ((0, 1, 2), (1, 1, 1)), rot_b)
        bc = blocks[0].bcList["east"]
        assert isinstance(bc, ExchangeBC_FullFace) and bc.otherBlock == 1
        seen.add((bc.otherFace, None if bc.cell_map is None else bc.cell_map[0, 0, 0].tobytes() + bc.cell_map[-1, 0, 0].tobytes()))
        sim = Simulation(cfg, gm, blocks, lib=oracle)
        sim.run()
        compare(physical_solution(sim, blocks), reference_solution)
        sim.close()
    assert len(seen) == 24 and {f for f, _ in seen} == set(range(6))


def test_every_face_of_the_first_block(oracle, reference_solution):
    gm = cases.ideal_air()
    all_rot = rotations()
    faces = set()
    for n, rot_a in enumerate(all_rot):
        rot_b = all_rot[(5 * n + 3) % 24]
        cfg, blocks = make_case(gm, rot_a, rot_b)
        (fa,) = [f for f, bc in blocks[0].bcList.items() if isinstance(bc, ExchangeBC_FullFace)]
        faces.add(fa)
        sim = Simulation(cfg, gm, blocks, lib=oracle)
        sim.run()
        compare(physical_solution(sim, blocks), reference_solution)
        sim.close()
    assert len(faces) == 6

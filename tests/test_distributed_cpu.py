"""world_size-2 run over gloo on CPU: the partitioned job (each rank owns half of the blocks,
ghost cells of connected blocks travel through the exchange callback) must reproduce the
single-process result bit for bit -- the mirror of the reference's cone20-mpi-test.rb, which
asserts the same numbers as the serial test.  Uses the CPU oracle as the implementation of the
C ABI, so this covers the host-side N>1 logic (block ownership, halo lists on both sides, the
callback, the dt all-reduce) without a GPU."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from conftest import build_oracle
    from gdtk_b200 import _abi, Simulation, cases
    from gdtk_b200.distributed import DistributedSimulation, distribute_blocks, octant_owner
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = _abi.load_library(build_oracle(), "orc_")
    ok = True
    def rotated_pair(**kw):
        # two 3D blocks whose common face is east of the one and bottom of the other, rotated: explicit cell maps
        from test_connections3d import make_case, rotations
        cfg, blocks = make_case(cases.ideal_air(), ((0, 1, 2), (1, 1, 1)), rotations()[17])
        assert blocks[0].bcList["east"].cell_map is not None
        cfg.max_step = 100
        return cfg, cases.ideal_air(), blocks

    for name, factory, kw, nsteps in (("box3d", cases.box3d, dict(n=16, nb=2), 12),
                                      ("rotated-pair", rotated_pair, dict(), 6),
                                      ("ffs", cases.ffs, dict(nx=60, ny=20), 25),
                                      ("cone20", cases.cone20, dict(), 30),
                                      ("cone20-adaptive", cases.cone20, dict(flux_calculator="adaptive_hanel_ausmdv"), 60),
                                      # walls without ghost cells and a static user-defined inflow profile, four blocks
                                      ("vortex", cases.vortex, dict(gfactor=2), 40)):
        cfg, gm, blocks = factory(**kw)
        if name == "box3d":
            owner = octant_owner({v: next(b for b in blocks if b.id == k) for k, v in cfg.block_index.items()}, 2, world)
        else:
            owner = distribute_blocks(blocks, world)
        sim = DistributedSimulation(cfg, gm, blocks, owner, lib=lib, device_buffers=False)
        sim.run(max_step=nsteps, max_time=1e30)
        mine = {b.id: [sim.interior(b.id, a).copy() for a in sim.download_conserved(b.id)] for b in sim.local_blocks}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        dts = sim.dt_history
        sim.close()
        if rank == 0:
            cfg, gm, blocks = factory(**kw)
            ref = Simulation(cfg, gm, blocks, lib=lib)
            ref.run(max_step=nsteps, max_time=1e30)
            nblk = 0
            for part in gathered:
                for bid, U in part.items():
                    nblk += 1
                    R = [ref.interior(bid, a) for a in ref.download_conserved(bid)]
                    if not all(np.array_equal(a, b) for a, b in zip(U, R)):
                        ok = False
                        print(f"MISMATCH {name} block {bid}")
            if nblk != len(blocks) or dts != ref.dt_history:
                ok = False
                print(f"MISMATCH {name}: blocks {nblk}/{len(blocks)} or dt history")
            ref.close()
            print(f"{name}: {world} ranks == 1 rank: {ok}", flush=True)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    dist.destroy_process_group()
    sys.exit(0 if flags[0] else 1)


def test_two_ranks_match_one_rank_gloo():
    port = 29600 + os.getpid() % 300
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__), "--worker"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "box3d: 2 ranks == 1 rank: True" in r.stdout
    assert "rotated-pair: 2 ranks == 1 rank: True" in r.stdout
    assert "cone20: 2 ranks == 1 rank: True" in r.stdout
    assert "cone20-adaptive: 2 ranks == 1 rank: True" in r.stdout
    assert "vortex: 2 ranks == 1 rank: True" in r.stdout


if __name__ == "__main__" and "--worker" in sys.argv:
    _worker()

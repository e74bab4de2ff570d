"""N-GPU result == 1-GPU result (mirror of the reference's cone20-mpi-test.rb asserting the same
numbers as the serial test).  Launched by tests/run_multi_gpu.sh under torchrun; each rank runs
the partitioned job, rank 0 also runs the whole job on its own GPU and compares bit for bit
(the exchange is a pure copy)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker():
    import torch
    import torch.distributed as dist
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gdtk_b200 import Simulation, cases
    from gdtk_b200.distributed import DistributedSimulation, octant_owner, distribute_blocks
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    transports = set()
    for name, factory, kw, nsteps in (("box3d", cases.box3d, dict(n=32, nb=2), 12),
                                      # odd sizes: a trailing tile / k-chunk one cell wide next to a remote face (the tile before
                                      # it reads remote ghost cells too), and odd padded widths (no TMA: the face-centred kernel)
                                      ("box3d-66", cases.box3d, dict(n=66, nb=2), 6),
                                      ("box3d-130", cases.box3d, dict(n=130, nb=2), 3),
                                      ("ffs", cases.ffs, dict(nx=120, ny=40), 30),
                                      ("cone20", cases.cone20, dict(), 40),
                                      ("cone20-adaptive", cases.cone20, dict(flux_calculator="adaptive_hanel_ausmdv"), 60),
                                      ("sod-adaptive", cases.sod, dict(dims=3, ncells=64, nj=4, nk=4, nblocks=4, flux_calculator="adaptive_hanel_ausmdv"), 40)):
        for strict in (True, False):
            cfg, gm, blocks = factory(**kw)
            cfg.strict_fp = strict
            if name.startswith("box3d"):
                owner = octant_owner({v: next(b for b in blocks if b.id == k) for k, v in cfg.block_index.items()}, 2, world)
            else:
                owner = distribute_blocks(blocks, world)
            # halo by direct NVLink stores (the default) and, for the first case, also through the exchange callback (NCCL)
            sim = DistributedSimulation(cfg, gm, blocks, owner, device=local, direct_halo=not (name == "box3d" and strict))
            transports.add(sim.halo_transport)
            sim.run(max_step=nsteps, max_time=1e30)
            mine = {b.id: [sim.interior(b.id, a).copy() for a in sim.download_conserved(b.id)] for b in sim.local_blocks}
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            dts = sim.dt_history
            sim.close()
            if rank == 0:
                cfg, gm, blocks = factory(**kw)
                cfg.strict_fp = strict
                ref = Simulation(cfg, gm, blocks, device=local)
                ref.run(max_step=nsteps, max_time=1e30)
                for part in gathered:
                    for bid, U in part.items():
                        R = [ref.interior(bid, a) for a in ref.download_conserved(bid)]
                        same = all(np.array_equal(a, b) for a, b in zip(U, R))
                        if not same:
                            ok = False
                            print(f"MISMATCH {name} strict={strict} block {bid}: max diff {max(np.abs(a - b).max() for a, b in zip(U, R)):.3e}")
                if dts != ref.dt_history:
                    ok = False
                    print(f"MISMATCH dt history {name} strict={strict}")
                ref.close()
                print(f"{name} strict={strict}: {world}-GPU == 1-GPU: {ok}", flush=True)
    if rank == 0:
        print("halo transports used:", sorted(transports), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def test_two_gpus_match_one_gpu():
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__), "--worker"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


if __name__ == "__main__" and "--worker" in sys.argv:
    _worker()

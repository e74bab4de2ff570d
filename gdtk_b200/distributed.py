"""One process per GPU: block-to-rank mapping and the halo-exchange callback.

Replaces the reference's MPI plumbing for this path:

* ``mpiDistributeBlocks{ntasks=, dist="load-balance"}``  src/eilmer/mpi.lua:8-130
  (-> ``distribute_blocks``; config/<job>.mpimap, src/eilmer/simcore.d:149-179)
* ``MPI_Irecv / MPI_Send / MPI_Wait`` per block face    full_face_copy.d:1681,1803-1842
  (-> one packed buffer per peer rank, moved by NCCL send/recv on the library's stream;
  on CPU, for the gloo tests, the same code moves host buffers)
* ``MPI_Allreduce`` of dt_allow / cfl_max                simcore_gasdynamic_step.d:105-107
  (-> one all_reduce(MIN) of the pair (dt_allow, -cfl_max))

torch.distributed is plumbing only; nothing numerical happens here.
"""
import ctypes as C

import numpy as np

from .sim import Simulation


def distribute_blocks(blocks, world_size, dims=3, mode="slab"):
    """Assign blocks to ranks.  'slab': contiguous runs of the block list (blocks made by
    cases.box3d are ordered i-major, so ranks own slabs/pencils/octants for 2/4/8 ranks when
    the block array is 4x4x4); 'load-balance': largest block first onto the least loaded rank
    (mpi.lua:60-130)."""
    if mode == "slab":
        n = len(blocks)
        return {b.id: min(world_size - 1, (idx * world_size) // n) for idx, b in enumerate(blocks)}
    loads = [0] * world_size
    owner = {}
    def ncells(b):
        g = b.grid
        if hasattr(g, "nic"):
            return g.nic * g.njc * g.nkc
        return int(np.prod([s - 1 for s in np.shape(g[0])]))
    for b in sorted(blocks, key=ncells, reverse=True):
        r = loads.index(min(loads))
        owner[b.id] = r
        loads[r] += ncells(b)
    return owner


def octant_owner(blocks_by_index, nb, world_size):
    """2x2x2-style ownership for a cubic nb^3 block array: split the i, j, k block ranges in
    two as many times as world_size has factors of two (8 ranks -> octants)."""
    splits = [1, 1, 1]
    w, d = world_size, 0
    while w > 1:
        splits[d % 3] *= 2
        w //= 2
        d += 1
    owner = {}
    for (ib, jb, kb), blk in blocks_by_index.items():
        r = 0
        for s, x in zip(splits, (ib, jb, kb)):
            r = r * s + (x * s) // nb
        owner[blk.id] = r
    return owner


class _DevArray:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {
            "shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}


def make_exchange(device_buffers):
    """Return the eb200_exchange_fn callback body (a Python callable).

    device_buffers=True: pointers are CUDA device memory, transport = NCCL on the given stream.
    device_buffers=False: pointers are host memory (oracle / gloo tests)."""
    import torch
    import torch.distributed as dist

    views = {}        # the library's halo buffers never move: wrap each (pointer, length) once

    def wrap(ptr, n):
        if n == 0:
            return None
        t = views.get((ptr, n))
        if t is None:
            if device_buffers:
                t = torch.as_tensor(_DevArray(ptr, n), device="cuda")
            else:
                buf = (C.c_double * n).from_address(ptr)
                t = torch.from_numpy(np.frombuffer(buf, dtype=np.float64))
            views[(ptr, n)] = t
        return t

    def exchange(user, npeers, peers, send, send_count, recv, recv_count, stream):
        try:
            ops = []
            for p in range(npeers):
                r = peers[p]
                st = wrap(send[p], send_count[p])
                rt = wrap(recv[p], recv_count[p])
                if rt is not None:
                    ops.append(dist.P2POp(dist.irecv, rt, r))
                if st is not None:
                    ops.append(dist.P2POp(dist.isend, st, r))
            if not ops:
                return 0
            if device_buffers:
                ext = torch.cuda.ExternalStream(int(stream))
                with torch.cuda.stream(ext):
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()          # stream-ordered for NCCL: the host is not blocked
            else:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            return 0
        except Exception as e:          # never unwind across the C ABI
            print(f"[eb200 exchange] {type(e).__name__}: {e}", flush=True)
            return 1

    return exchange


class DistributedSimulation(Simulation):
    """Simulation with blocks spread over the ranks of torch.distributed."""

    def __init__(self, config, gmodel, blocks, block_owner, lib=None, device_buffers=True, device=0, direct_halo=True):
        import torch.distributed as dist
        self._dist = dist
        self._device_buffers = device_buffers
        rank, world = dist.get_rank(), dist.get_world_size()
        super().__init__(config, gmodel, blocks, lib=lib, rank=rank, world_size=world,
                         block_owner=block_owner, device=device, exchange=make_exchange(device_buffers))

        self.halo_transport = "exchange callback"
        if device_buffers and direct_halo:
            self._setup_direct_halo()

    def _peer_ranks(self):
        from .sim import ExchangeBC_FullFace
        peers = set()
        for b in self.local_blocks:
            for bc in b.bcList.values():
                if isinstance(bc, ExchangeBC_FullFace) and self.block_owner[bc.otherBlock] != self.rank:
                    peers.add(self.block_owner[bc.otherBlock])
        return sorted(peers)

    def _setup_direct_halo(self):
        """eb200_p2p_export / eb200_p2p_import (include/eb200.h): every rank hands each halo peer a blob with the CUDA
        IPC handles of its arena; afterwards the library stores halo cells straight into the neighbours' ghost cells
        over NVLink and this module's exchange callback is no longer called.  All ranks decide together: if any
        import fails (ranks on different nodes, IPC not permitted), everybody stays on the callback."""
        dist, lib, h = self._dist, self.lib, self.handle
        peers = self._peer_ranks()
        size = lib.p2p_export(h, -1, None, 0) if peers else 0
        mine = {}
        ok = True
        for p in peers:
            buf = C.create_string_buffer(max(size, 1))
            if lib.p2p_export(h, p, buf, size) != size:
                ok = False
                break
            mine[p] = buf.raw
        everyone = [None] * self.world_size
        dist.all_gather_object(everyone, mine if ok else None)
        ok = ok and all(e is not None for e in everyone)
        ready = not peers
        if ok:
            for p in peers:
                blob = everyone[p].get(self.rank)
                rc = lib.p2p_import(h, p, blob, len(blob)) if blob is not None else -1
                if rc < 0:
                    ok = False
                    break
                ready = rc == 1
        import torch
        t = torch.tensor([1 if (ok and ready) else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) != 1:
            if ok and ready and peers:
                raise RuntimeError("direct halo exchange is set up on this rank but not on all ranks: "
                                   f"{lib.error()} -- rerun with direct_halo=False")
            return
        self.halo_transport = "direct NVLink stores (CUDA IPC)"

    def reduce_step_status(self, rc):
        import torch
        t = torch.tensor([2 if rc < 0 else rc], dtype=torch.int32)
        if self._device_buffers:
            t = t.cuda()
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        worst = int(t.cpu()[0])
        return -2 if worst >= 2 else worst

    def reduce_dt(self, dt_allow, cfl_max):
        import torch
        t = torch.tensor([dt_allow, -cfl_max], dtype=torch.float64)
        if self._device_buffers:
            t = t.cuda()
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MIN)
        t = t.cpu()
        return float(t[0]), -float(t[1])

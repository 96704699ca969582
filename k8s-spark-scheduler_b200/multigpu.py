"""One process per GPU on one box: what `bench.py --gpus N` runs between the ranks.

The placement path has NO exchange step -- pending applications are independent units against one snapshot (independent
mode) and the FIFO queues of different instance groups are independent -- so the only things that move between the
processes are the multi-GPU analogues of H2D and D2H:

  * the node snapshot: ONE broadcast of a flat buffer from the rank that owns the cluster state (`broadcast_snapshot`);
  * the placements: every rank copies ITS block straight into the scheduler's result buffer, a POSIX shared-memory segment
    that all worker processes of the box map and page-lock (`SharedResults`); on the GPU box the copy is a DMA over the
    rank's own PCIe link (gp_register_host + gp_pack_batch_wire), no gather through one GPU and no collective.

The same code runs over gloo on CPU in tests/test_multigpu_gloo.py (world size 2, the per-rank packer being the oracle).
A single host process that drives several GPUs uses gp_multi_* (csrc/gangpack_multi.cu) instead.
"""
from __future__ import annotations

from multiprocessing import shared_memory
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(q: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of the queue owned by `rank` (keeps queue order across ranks)."""
    return (q * rank) // world, (q * (rank + 1)) // world


def broadcast_snapshot(flat: torch.Tensor, src: int = 0, group=None) -> None:
    """One broadcast of the flat snapshot buffer [cpu | mem | gpu | orders] from the rank that owns the cluster state."""
    dist.broadcast(flat, src=src, group=group)


class SharedResults:
    """The scheduler's result buffer, shared by the worker processes of one box.

    Rank r owns bytes [base[r], base[r+1]): `driver` (int32 x q_r) followed by `executors` (node_dtype x total_r).
    Rank 0 -- the consumer -- sees every rank's block through `block(r, q_r, total_r)`."""

    def __init__(self, name: str, q: int, total_exec: int, node_dtype, device=None, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.node_dtype = np.dtype(node_dtype)
        self.q, self.total = int(q), int(total_exec)
        my_bytes = ((4 * self.q + self.node_dtype.itemsize * max(self.total, 1)) + 255) & ~255
        dev = device if device is not None else torch.device("cpu")
        sizes = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(sizes, torch.tensor([my_bytes, self.q, self.total], dtype=torch.int64, device=dev), group=group)
        self.sizes = np.stack([s.cpu().numpy() for s in sizes])
        self.base = np.concatenate([[0], np.cumsum(self.sizes[:, 0])]).astype(np.int64)
        if self.rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()      # stale segment of a killed run
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=int(self.base[-1]))
        dist.barrier(group=group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
            try:        # Python < 3.13 registers attached segments with its resource tracker, which would unlink rank 0's segment
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        dist.barrier(group=group)
        self.whole = np.frombuffer(self.shm.buf, dtype=np.uint8, count=int(self.base[-1]))
        self.mine = self.whole[int(self.base[self.rank]):int(self.base[self.rank + 1])]
        self.driver, self.executors = self._split(self.mine, self.q, self.total)
        self._group = group

    def _split(self, block, q, total):
        d = block[:4 * q].view(np.int32)
        e = block[4 * q:4 * q + self.node_dtype.itemsize * max(total, 1)].view(self.node_dtype)
        return d, e

    def block(self, r: int):
        """(driver_node, executor_nodes) of rank r as the consumer sees them."""
        b = self.whole[int(self.base[r]):int(self.base[r + 1])]
        return self._split(b, int(self.sizes[r, 1]), int(self.sizes[r, 2]))

    def close(self):
        self.whole = self.mine = self.driver = self.executors = None
        try:
            self.shm.close()
        except BufferError:
            pass
        dist.barrier(group=self._group)
        if self.rank == 0:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass

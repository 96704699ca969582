"""App-sharded gang placement over the GPUs of one box (independent mode).

One process per GPU (torch.distributed, NCCL over NVLink).  The path shards naturally: pending
applications are independent units against one snapshot, so

  1. rank 0 owns the node snapshot -> ONE broadcast of (avail cpu, mem, gpu, executor order, driver order);
  2. every rank packs its contiguous block of the queue on its own GPU (no collective in the data path);
  3. the emitted placements (driver node per app, ExecutorNodes) are all-gathered, padded to the largest
     shard so the collective is regular, and re-assembled in queue order.

FIFO mode shards by instance group instead (whole groups -> ranks; a group's queue is strictly
sequential): `assign_groups`.

The collectives work on whatever backend the process group has: NCCL on the GPU box, gloo in the CPU
tests (tests/test_multigpu_gloo.py), where the per-shard pack function is supplied by the test.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(q: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of the queue owned by `rank` (keeps queue order across ranks)."""
    return (q * rank) // world, (q * (rank + 1)) // world


def assign_groups(cost: List[int], world: int) -> List[int]:
    """Greedy longest-processing-time assignment of instance groups to ranks (FIFO mode):
    cost[g] ~ apps_g * nodes_g.  Returns owner rank per group."""
    owner = [0] * len(cost)
    load = [0] * world
    for g in sorted(range(len(cost)), key=lambda i: -cost[i]):
        r = min(range(world), key=lambda i: load[i])
        owner[g] = r
        load[r] += cost[g]
    return owner


def broadcast_snapshot(t: Dict[str, torch.Tensor], src: int = 0, group=None) -> None:
    """One broadcast per snapshot array from the rank that owns the cluster state."""
    for k in ("cpu", "mem", "gpu", "eorder", "dorder"):
        if k in t and t[k] is not None:
            dist.broadcast(t[k], src=src, group=group)


def gather_placements(driver_local: torch.Tensor, exec_local: torch.Tensor, n_exec_local: int, group=None):
    """All-gather (driver_node, executor_nodes) of every shard and re-assemble them in queue order.

    driver_local: int32 [q_local]; exec_local: int32 [>= n_exec_local].
    Returns (driver_all int32 [sum q], exec_all int32 [sum n_exec], exec_base int64 [world+1]) where
    shard r's executor offsets must be shifted by exec_base[r]."""
    world = dist.get_world_size(group)
    dev = driver_local.device
    sizes = torch.tensor([driver_local.numel(), int(n_exec_local)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    max_q, max_e = int(all_sizes[:, 0].max()), max(int(all_sizes[:, 1].max()), 1)
    dpad = torch.full((max_q,), -9, dtype=torch.int32, device=dev)
    dpad[: driver_local.numel()] = driver_local
    epad = torch.full((max_e,), -9, dtype=torch.int32, device=dev)
    epad[: int(n_exec_local)] = exec_local[: int(n_exec_local)]
    dg = torch.empty(max_q * world, dtype=torch.int32, device=dev)
    eg = torch.empty(max_e * world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(dg, dpad, group=group)
    dist.all_gather_into_tensor(eg, epad, group=group)
    dparts = [dg[r * max_q: r * max_q + int(all_sizes[r, 0])] for r in range(world)]
    eparts = [eg[r * max_e: r * max_e + int(all_sizes[r, 1])] for r in range(world)]
    exec_base = np.zeros(world + 1, dtype=np.int64)
    np.cumsum(all_sizes[:, 1], out=exec_base[1:])
    return torch.cat(dparts), torch.cat(eparts), exec_base


def sharded_pack(apps: Dict[str, np.ndarray], snapshot: Dict[str, torch.Tensor],
                 pack_shard: Callable[[Dict[str, np.ndarray], Dict[str, torch.Tensor]], Tuple[torch.Tensor, torch.Tensor, int]],
                 group=None):
    """Full multi-GPU round for an independent batch.

    apps: the WHOLE queue as host arrays (every rank sees the same queue, e.g. decoded from the same
    request); snapshot: device tensors, valid on rank 0 (others receive them).  pack_shard(local_apps,
    snapshot) -> (driver_local, exec_local, n_exec_local) runs the single-GPU hot path.
    Returns (driver_node [q], executor_nodes [sum count]) in queue order on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    q = len(apps["count"])
    lo, hi = shard_bounds(q, rank, world)
    broadcast_snapshot(snapshot, 0, group)
    local = {k: (np.ascontiguousarray(v[lo:hi]) if v is not None else None) for k, v in apps.items() if k != "off"}
    d_local, e_local, n_e = pack_shard(local, snapshot)
    d_all, e_all, _ = gather_placements(d_local, e_local, n_e, group)
    return d_all, e_all

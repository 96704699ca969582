"""Seeded synthetic clusters and pending-app queues.

Nodes: PRNG splitmix64 seed 0xB200; apps: seed 0x5CED.  Units: CPU millicores, memory bytes.
The node priority order is the host-side restatement of NodeSorter.PotentialNodes for a single zone
without label priorities (internal/sort/nodesorting.go:41-122): available memory ascending, then CPU
ascending, then name ascending => fullest nodes first.
"""
from __future__ import annotations

import numpy as np

Gi = 1 << 30
Mi = 1 << 20
NODE_SEED = 0xB200
APP_SEED = 0x5CED


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of splitmix64 started at `seed` (vectorised: state_i = seed + (i+1)*gamma)."""
    with np.errstate(over="ignore"):
        gamma = np.uint64(0x9E3779B97F4A7C15)
        z = np.uint64(seed) + gamma * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _pick(r: np.ndarray, choices) -> np.ndarray:
    c = np.asarray(choices, dtype=np.int64)
    return c[(r % np.uint64(len(c))).astype(np.int64)]


def make_nodes(n: int, seed: int = NODE_SEED, groups: int = 1, gpu_variant: bool = False, fill=(0.0, 0.9)) -> dict:
    """fill = (lo, hi): pre-existing usage fraction u ~ U[lo, hi) (U[0, 0.9); the deep-scan workloads use
    a nearly full cluster, e.g. (0.90, 1.0))."""
    r = splitmix64(seed, 3 * n).reshape(3, n)
    t = (r[0] % np.uint64(4)).astype(np.int64)                       # instance type
    alloc_cpu = np.array([16, 32, 64, 96], dtype=np.int64)[t] * 1000
    alloc_mem = np.array([64, 128, 256, 384], dtype=np.int64)[t] * Gi
    u = fill[0] + (r[1] >> np.uint64(11)).astype(np.float64) / float(1 << 53) * (fill[1] - fill[0])   # usage fraction U[lo,hi)
    used_cpu = (np.floor(alloc_cpu * u / 250.0)).astype(np.int64) * 250      # quantised to 250m
    used_mem = (np.floor(alloc_mem * u / (256.0 * Mi))).astype(np.int64) * (256 * Mi)
    alloc_gpu = _pick(r[2], [0, 8]) if gpu_variant else np.zeros(n, dtype=np.int64)
    used_gpu = np.zeros(n, dtype=np.int64)
    group = (np.arange(n, dtype=np.int64) % groups).astype(np.int32)
    return {
        "n": n,
        "alloc_cpu": alloc_cpu, "alloc_mem": alloc_mem, "alloc_gpu": alloc_gpu,
        "avail_cpu": alloc_cpu - used_cpu, "avail_mem": alloc_mem - used_mem, "avail_gpu": alloc_gpu - used_gpu,
        "group": group, "groups": groups,
    }


def node_names(n: int) -> list:
    return ["node-%06d" % i for i in range(n)]


def priority_order(avail_cpu, avail_mem, subset=None) -> np.ndarray:
    """Node indices in scheduling priority order (single zone, no label priority).
    Names are node-%06d so name order == index order."""
    idx = np.arange(len(avail_cpu), dtype=np.int64) if subset is None else np.asarray(subset, dtype=np.int64)
    key = np.lexsort((idx, np.asarray(avail_cpu)[idx], np.asarray(avail_mem)[idx]))
    return idx[key].astype(np.int32)


def group_orders(nodes: dict):
    """Per-instance-group executor / driver orders as CSR (off[G+1], order[N])."""
    g = nodes["groups"]
    offs = [0]
    parts = []
    for k in range(g):
        members = np.nonzero(nodes["group"] == k)[0]
        parts.append(priority_order(nodes["avail_cpu"], nodes["avail_mem"], members))
        offs.append(offs[-1] + len(members))
    order = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    return np.asarray(offs, dtype=np.int32), order


def make_apps(q: int, seed: int = APP_SEED, groups: int = 1, da_sweep: bool = False,
              young_frac: float = 0.0, gpu_variant: bool = False, readme_app: bool = False, deep: bool = False) -> dict:
    r = splitmix64(seed, 8 * q).reshape(8, q)
    if readme_app:  # README.md:37-41
        drv_cpu = np.full(q, 1000, np.int64); drv_mem = np.full(q, 1 * Gi, np.int64)
        exe_cpu = np.full(q, 2000, np.int64); exe_mem = np.full(q, 4 * Gi, np.int64)
        count = np.full(q, 8, np.int32)
    else:
        drv_cpu = _pick(r[0], [1, 2]) * 1000
        drv_mem = _pick(r[1], [1, 2, 4]) * Gi
        exe_cpu = _pick(r[2], [1, 2, 4]) * 1000
        exe_mem = _pick(r[3], [2, 4, 8, 16]) * Gi
        count = (r[4] % np.uint64(32)).astype(np.int32) + 1              # U{1..32}
    if deep:        # deep-scan queue for a nearly full cluster (make_nodes(fill=(0.95, 1.0))): gangs of 4..128 executors that
        # walk ~8 000 nodes of the priority order before they find room, and one third of the applications ask for
        # 8-core executors that fit nowhere (no node has 8 free cores) -> full-table scan, no fit
        exe_cpu = _pick(r[2], [1, 2, 8]) * 1000
        count = count * 4
    max_count = count.copy()
    if da_sweep:  # dynamic allocation: only MIN is packed (EXT/resource.go:242,325)
        count = _pick(r[4], [0, 1, 2, 4, 8, 16]).astype(np.int32)
        max_count = (count + _pick(r[5], [0, 4, 16])).astype(np.int32)
    drv_gpu = np.zeros(q, np.int64)
    exe_gpu = _pick(r[6], [0, 0, 0, 1]) if gpu_variant else np.zeros(q, np.int64)
    young = ((r[7] >> np.uint64(11)).astype(np.float64) / float(1 << 53) < young_frac).astype(np.uint8)
    group = (np.arange(q, dtype=np.int64) % groups).astype(np.int32)
    return {
        "q": q,
        "drv_cpu": drv_cpu, "drv_mem": drv_mem, "drv_gpu": drv_gpu,
        "exe_cpu": exe_cpu, "exe_mem": exe_mem, "exe_gpu": exe_gpu,
        "count": count, "max_count": max_count, "young": young, "group": group, "groups": groups,
    }


def exec_offsets(count) -> np.ndarray:
    """exclusive prefix sum of executor counts: app i owns executor_nodes[off[i]:off[i+1]]."""
    off = np.zeros(len(count) + 1, dtype=np.int64)
    np.cumsum(np.asarray(count, dtype=np.int64), out=off[1:])
    return off

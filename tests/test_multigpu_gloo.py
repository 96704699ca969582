"""The N>1 path on CPU: world_size-2 `gloo` process group drives the same sharding / broadcast /
all-gather / re-assembly code (k8s-spark-scheduler_b200/multigpu.py) that bench.py runs over NCCL.
The per-shard pack function is the CPU oracle here (the product has no CPU pack path); the result must
equal a single-process run over the whole queue."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, algo, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import k8s_spark_scheduler_b200.multigpu as mg
        import k8s_spark_scheduler_b200.synth as synth
        from oracle import oracle as orc
        from helpers import res_aos
        nodes = synth.make_nodes(300)
        apps = synth.make_apps(257)          # not divisible by 2: uneven shards, uneven executor totals
        order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
        # only rank 0 knows the snapshot
        snap = {k: torch.from_numpy(np.ascontiguousarray(v)).clone() for k, v in
                (("cpu", nodes["avail_cpu"]), ("mem", nodes["avail_mem"]), ("gpu", nodes["avail_gpu"]),
                 ("eorder", order), ("dorder", order))}
        if rank != 0:
            for v in snap.values():
                v.zero_()

        def pack_shard(local, s):
            drv = res_aos(local["drv_cpu"], local["drv_mem"], local["drv_gpu"])
            exe = res_aos(local["exe_cpu"], local["exe_mem"], local["exe_gpu"])
            _, dn, en, off, _ = orc.closed_batch(algo, 0, s["cpu"].numpy(), s["mem"].numpy(), s["gpu"].numpy(),
                                                 s["dorder"].numpy(), s["eorder"].numpy(), drv, exe, local["count"])
            return torch.from_numpy(dn), torch.from_numpy(en), int(off[-1])

        a = {k: apps[k] for k in ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count")}
        d_all, e_all = mg.sharded_pack(a, snap, pack_shard)
        # single-process truth over the whole queue
        drv = res_aos(a["drv_cpu"], a["drv_mem"], a["drv_gpu"]); exe = res_aos(a["exe_cpu"], a["exe_mem"], a["exe_gpu"])
        _, wd, we, woff, _ = orc.closed_batch(algo, 0, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"],
                                              order, order, drv, exe, a["count"])
        ok = bool(np.array_equal(d_all.numpy(), wd) and np.array_equal(e_all.numpy(), we[: int(woff[-1])]))
        ok = ok and bool(torch.equal(snap["cpu"], torch.from_numpy(nodes["avail_cpu"])))   # broadcast reached this rank
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("algo", [0, 1])
def test_sharded_pack_world2_gloo(oracle, algo):
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, algo, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_shard_bounds_and_group_assignment():
    import k8s_spark_scheduler_b200.multigpu as mg
    for q in (0, 1, 7, 100000):
        for world in (1, 2, 4, 8):
            b = [mg.shard_bounds(q, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == q
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    owner = mg.assign_groups([10, 9, 8, 1, 1, 1], 2)
    loads = [sum(c for c, o in zip([10, 9, 8, 1, 1, 1], owner) if o == r) for r in range(2)]
    assert sum(loads) == 30 and max(loads) <= 20   # LPT: makespan <= 4/3 * optimum (15)

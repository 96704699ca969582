"""The N>1 path on CPU: a world_size-2 `gloo` process group drives the same code `bench.py --gpus N` runs between the ranks
(k8s-spark-scheduler_b200/multigpu.py): ONE broadcast of the flat snapshot buffer from the rank that owns the cluster
state, every rank packs its contiguous block of the queue, and every rank writes ITS placements straight into the
scheduler's shared-memory result buffer (no gather through one rank, no collective on the results).  The per-rank packer
is the CPU oracle here (the product has no CPU pack path); rank 0 -- the consumer -- must find exactly the placements of
a single-process run over the whole queue."""
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
        n, q = 300, 257                       # q not divisible by 2: uneven blocks, uneven executor totals
        nodes = synth.make_nodes(n)
        apps = synth.make_apps(q)
        order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
        # the flat snapshot buffer of bench.py: [cpu | mem | gpu | order]; only rank 0 knows it
        flat = torch.zeros(3 * n + (n + 1) // 2, dtype=torch.int64)
        if rank == 0:
            flat[0:n] = torch.from_numpy(nodes["avail_cpu"]); flat[n:2 * n] = torch.from_numpy(nodes["avail_mem"])
            flat[2 * n:3 * n] = torch.from_numpy(nodes["avail_gpu"])
            flat[3 * n:].view(torch.int32)[:n] = torch.from_numpy(order)
        mg.broadcast_snapshot(flat, src=0)
        cpu, mem, gpu = flat[0:n].numpy(), flat[n:2 * n].numpy(), flat[2 * n:3 * n].numpy()
        eorder = flat[3 * n:].view(torch.int32)[:n].numpy()
        lo, hi = mg.shard_bounds(q, rank, world)
        keys = ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count")
        a = {k: np.ascontiguousarray(apps[k][lo:hi]) for k in keys}
        drv = res_aos(a["drv_cpu"], a["drv_mem"], a["drv_gpu"]); exe = res_aos(a["exe_cpu"], a["exe_mem"], a["exe_gpu"])
        _, dn, en, off, _ = orc.closed_batch(algo, 0, cpu, mem, gpu, eorder, eorder, drv, exe, a["count"])
        total = int(off[-1])
        res = mg.SharedResults(f"gangpack_test_{port}", hi - lo, total, np.uint16)
        res.driver[:] = dn
        res.executors[:total] = en[:total].astype(np.uint16)
        dist.barrier()
        ok = True
        if rank == 0:                          # the consumer: every rank's block against the single-process truth
            A = {k: apps[k] for k in keys}
            _, wd, we, woff, _ = orc.closed_batch(algo, 0, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order,
                                                  res_aos(A["drv_cpu"], A["drv_mem"], A["drv_gpu"]), res_aos(A["exe_cpu"], A["exe_mem"], A["exe_gpu"]), A["count"])
            got_d = np.concatenate([res.block(r)[0] for r in range(world)])
            got_e = np.concatenate([res.block(r)[1][: int(res.sizes[r, 2])] for r in range(world)])
            ok = bool(np.array_equal(got_d, wd) and np.array_equal(got_e.astype(np.int32), we[: int(woff[-1])]))
        ok = ok and bool(np.array_equal(cpu, nodes["avail_cpu"]))         # the broadcast reached this rank
        ret[rank] = ok
        res.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("algo", [0, 1])
def test_world_size_2_gloo(algo):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, algo, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_bounds_cover_the_queue():
    sys.path.insert(0, ROOT)
    import k8s_spark_scheduler_b200.multigpu as mg
    for q in (0, 1, 7, 100000, 1000003):
        for world in (1, 2, 3, 8):
            b = [mg.shard_bounds(q, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == q and all(b[i][1] == b[i + 1][0] for i in range(world - 1))

#!/usr/bin/env python
"""fitEarlierDrivers with single-az-tightly-pack: ONE launch (gp_pack_fifo_zones) against the per-driver call sequence it
replaces (gp_pack_batch_zones for one application + gp_reserve_placements(subtract), exact accounting so that both run the
same arithmetic).  3 000 nodes in 3 zones, a queue of 2 000 drivers; host buffers in, host results out, wall clock.
Prints one JSON line; the two paths must agree bit for bit."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import k8s_spark_scheduler_b200 as g  # noqa: E402
from k8s_spark_scheduler_b200 import synth  # noqa: E402


def main():
    n, Z, q, q_loop = 3000, 3, 2000, 300
    nodes = synth.make_nodes(n)
    cpu, mem, gpu = nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"]
    order = synth.priority_order(cpu, mem)
    zone_of = np.arange(n) % Z
    eo = np.concatenate([order[zone_of[order] == z] for z in range(Z)]).astype(np.int32)
    eoff = np.concatenate([[0], np.cumsum([(zone_of == z).sum() for z in range(Z)])]).astype(np.int32)
    apps = synth.make_apps(q, seed=11)
    keys = ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count")
    a = {k: apps[k] for k in keys}
    a["young"] = np.ones(q, np.uint8)
    p = g.GangPacker()
    sc, sm, sg = cpu + 4000, mem + (8 << 30), gpu

    def fresh():
        p.set_snapshot(cpu, mem, gpu, eo, eo, eoff, eoff)
        p.set_schedulable(sc, sm, sg)

    ts = []
    for _ in range(5):
        fresh()
        t0 = time.perf_counter()
        zone, gd, ge, goff, _ = p.pack_fifo_zones(a, 0, 2)
        ts.append(time.perf_counter() - t0)
    one_launch_ms = float(np.min(ts[1:])) * 1e3
    # the per-driver sequence on the first q_loop drivers
    fresh()
    t0 = time.perf_counter()
    ld = np.full(q_loop, -9, np.int32)
    lex = []
    for i in range(q_loop):
        one = {k: a[k][i:i + 1] for k in keys}
        z1, d1, e1, o1, _ = p.pack_batch_zones(one, 0)
        ld[i] = d1[0]
        lex.append(e1.copy() if d1[0] >= 0 else np.full(int(a["count"][i]), -9, np.int32))
        if d1[0] >= 0:
            p.reserve_placements(one, (d1, e1, o1), subtract=True)
    loop_ms = (time.perf_counter() - t0) * 1e3
    same = bool(np.array_equal(ld, gd[:q_loop]))
    for i in range(q_loop):
        if gd[i] >= 0:
            same = same and bool(np.array_equal(lex[i], ge[goff[i]:goff[i + 1]]))
    print(json.dumps({"workload": "single-az-tightly-pack under FIFO, %d nodes in %d zones, queue of %d drivers (exact accounting)" % (n, Z, q),
                      "one_launch_ms": one_launch_ms, "one_launch_us_per_driver": one_launch_ms * 1e3 / q,
                      "placed": int((gd >= 0).sum()), "per_driver_calls_us_per_driver": loop_ms * 1e3 / q_loop,
                      "per_driver_sample": q_loop, "paths_agree": same}))
    p.close()


if __name__ == "__main__":
    main()

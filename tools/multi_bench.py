#!/usr/bin/env python
"""BASELINE configs[3] / configs[4] through the SINGLE-PROCESS multi-GPU handle (gp_multi_*, include/gangpack.h):

  python tools/multi_bench.py --config 3 --devices 4    # DA sweep, 10k nodes x 50k apps, FIFO on, 16 instance groups -> 4 GPUs
  python tools/multi_bench.py --config 4 --devices 8    # 50k nodes x 1M apps, tightly-pack, app-sharded over 8 GPUs

One host process (what a cgo scheduler is), pinned host buffers in, host results out; every device copies its own block
of the placements over its own PCIe link.  Wall clock around gp_multi_set_snapshot + gp_multi_pack_batch; the results are
verified against a single-context run of the same batch (bit-exact) outside the timed region.  Prints one JSON line."""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, choices=[3, 4], required=True)
    ap.add_argument("--devices", type=int, default=0, help="number of GPUs (0 = all visible)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    import torch
    nd = args.devices or torch.cuda.device_count()
    devices = list(range(nd))
    if args.config == 3:
        nodes = synth.make_nodes(10000, groups=16)
        apps = synth.make_apps(50000, groups=16, da_sweep=True)
        algo, mode, desc = 0, 1, "DA min/max sweep, 10k nodes x 50k apps, FIFO (reference accounting), 16 instance groups"
        wire = dict(quantity_bits=64, node_bits=32, offsets=True)
        src = {k: apps[k] for k in ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count", "group", "young")}
    else:
        nodes = synth.make_nodes(50000)
        apps = synth.make_apps(1000000)
        algo, mode, desc = 0, 0, "50k nodes x 1M apps, tightly-pack, independent, contiguous block of the queue per GPU"
        wire = dict(quantity_bits=32, mem_shift=20, node_bits=16, offsets=False)
        src = g.native.compact_apps({k: apps[k] for k in ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu")}, 20)
        src["count"] = apps["count"]
    eoff, eorder = synth.group_orders(nodes)
    q = len(apps["count"])
    off = synth.exec_offsets(apps["count"])
    total = int(off[-1])
    m = g.MultiGangPacker(devices)
    pin = {}
    for k, v in src.items():
        if k in ("drv_gpu", "exe_gpu") and not np.asarray(v).any():
            continue
        pin[k] = m.pinned(len(v), np.asarray(v).dtype); pin[k][:] = v
    if wire["offsets"]:
        pin["off"] = m.pinned(q + 1, np.int64); pin["off"][:] = off
    pn = {k: m.pinned(len(v), v.dtype) for k, v in (("cpu", nodes["avail_cpu"]), ("mem", nodes["avail_mem"]), ("gpu", nodes["avail_gpu"]),
                                                      ("eorder", eorder), ("eoff", eoff))}
    for k, v in (("cpu", nodes["avail_cpu"]), ("mem", nodes["avail_mem"]), ("gpu", nodes["avail_gpu"]), ("eorder", eorder), ("eoff", eoff)):
        pn[k][:] = v
    od = m.pinned(q, np.int32)
    oe = m.pinned(max(total, 1), np.uint16 if wire["node_bits"] == 16 else np.int32)

    def step():
        m.set_snapshot(pn["cpu"], pn["mem"], pn["gpu"], pn["eorder"], pn["eorder"], pn["eoff"], pn["eoff"])
        m.pack_batch(pin, algo, mode, out=(od, oe), wire=wire)
        return int(od[0])

    for _ in range(args.warmup):
        step()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    ms = float(np.mean(ts)) * 1e3
    line = {"metric": "gang_placements_per_sec", "value": q / (ms * 1e-3), "unit": "decisions/s", "n_gpus": nd, "ms_per_step": ms,
            "ms_min": float(np.min(ts)) * 1e3, "steps": args.steps, "warmup": args.warmup, "path": "gp_multi_set_snapshot + gp_multi_pack_batch, one host process, host buffers in / out",
            "config": {"workload": desc, "nodes": int(nodes["n"]), "apps_total": q, "instance_groups": int(nodes["groups"]), "wire": wire},
            "h2d_bytes_per_step": int(sum(v.nbytes for v in pin.values())) + nd * int(sum(v.nbytes for v in pn.values())),
            "d2h_bytes_per_step": int(od.nbytes + oe.itemsize * total)}
    if mode != 0:
        line["group_owner"] = m.group_owner().tolist()
    if not args.no_verify:
        # bit-exact against ONE context running the whole batch
        s = g.GangPacker(device=0)
        s.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], eorder, eorder, eoff, eoff)
        t0 = time.perf_counter()
        wd, we, _ = s.pack_batch(src, algo, mode, wire=wire)
        line["single_gpu_ms_unpinned"] = (time.perf_counter() - t0) * 1e3
        fits = wd >= 0
        emask = np.repeat(fits, apps["count"])
        mism = int((np.asarray(od) != wd).sum()) + int((np.asarray(oe[:total])[emask] != np.asarray(we)[emask]).sum())
        line["parity_checked"] = q
        line["mismatches"] = mism
        if mode != 0:
            a_, b_ = m.get_snapshot(), s.get_snapshot()
            line["final_snapshot_equal"] = bool(all(np.array_equal(x, y) for x, y in zip(a_, b_)))
        s.close()
    m.close()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Where does the end-to-end time of one batch go?  (diagnostic, not part of the bench contract)"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import k8s_spark_scheduler_b200 as g
from bench import WORKLOADS, make_workload

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "tightly-100k"]
nodes, a, eoff, eorder = make_workload(w, 0)
p = g.GangPacker(0)
q = len(a["count"]); total = int(a["off"][-1])
pin = {k: p.pinned(len(a[k]), a[k].dtype) for k in a}
for k in a: pin[k][:] = a[k]
if w["groups"] == 1: pin.pop("group")
if w["mode"] == 0: pin.pop("young")
od = p.pinned(q, np.int32); oe = p.pinned(max(total, 1), np.int32)

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e6

pn = {}
for k, v in (("cpu", nodes["avail_cpu"]), ("mem", nodes["avail_mem"]), ("gpu", nodes["avail_gpu"]), ("eorder", eorder), ("eoff", eoff)):
    pn[k] = p.pinned(len(v), v.dtype); pn[k][:] = v
snap_pageable = lambda: p.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], eorder, eorder, eoff, eoff)
snap = lambda: p.set_snapshot(pn["cpu"], pn["mem"], pn["gpu"], pn["eorder"], pn["eorder"], pn["eoff"], pn["eoff"])
print("set_snapshot pageable %8.1f us" % t(snap_pageable))
print("pack_one            %8.1f us" % t(lambda: p.pack_one(0, (1000, 1 << 30, 0), (2000, 4 << 30, 0), 8)))
pack = lambda: p.pack_batch(pin, w["algo"], w["mode"], out=(od, oe))
print("potential_nodes (device sort, 10k nodes) %8.1f us" % t(lambda: p.potential_nodes(pn["cpu"], pn["mem"])))
import k8s_spark_scheduler_b200.synth as _synth
for _nn in (50000, 500000):
    _n50 = _synth.make_nodes(_nn)
    _pc = p.pinned(_nn, np.int64); _pc[:] = _n50["avail_cpu"]
    _pm = p.pinned(_nn, np.int64); _pm[:] = _n50["avail_mem"]
    print("potential_nodes (device sort, %d nodes, pinned inputs) %8.1f us" % (_nn, t(lambda: p.potential_nodes(_pc, _pm), 5)))
    print("potential_nodes (device sort, %d nodes, pageable inputs) %8.1f us" % (_nn, t(lambda: p.potential_nodes(_n50["avail_cpu"], _n50["avail_mem"]), 5)))
_t0 = time.perf_counter(); _synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"]); print("numpy lexsort 10k nodes %8.1f us" % ((time.perf_counter() - _t0) * 1e6))
_R = 200000
_rng = np.random.default_rng(1)
_rnode = _rng.integers(0, len(nodes["avail_cpu"]), _R).astype(np.int32)
_res = [(_rng.integers(0, 4, _R) * 500).astype(np.int64), (_rng.integers(0, 8, _R) << 29).astype(np.int64), np.zeros(_R, np.int64)]
_alloc = [nodes["alloc_cpu"], nodes["alloc_mem"], nodes["alloc_gpu"]]
print("build_availability (10k nodes, 200k reservations) %8.1f us" % t(lambda: p.build_availability(_alloc, None, _rnode, _res)))
print("prepare_cluster (availability + sort + layout, chained) %8.1f us" % t(lambda: p.prepare_cluster(_alloc, None, _rnode, _res)))
print("set_snapshot        %8.1f us" % t(snap))
print("pack_batch (pinned) %8.1f us" % t(pack), p.stats())
# compact wire format: int32 millicores / MiB, no offsets, uint16 node indices
if w["mode"] == 0 and w["algo"] != 2:
    c32 = g.native.compact_apps({k: a[k] for k in ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu")}, 20)
    cols = [k for k in ("drv_cpu", "drv_mem", "exe_cpu", "exe_mem", "drv_gpu", "exe_gpu") if a[k].any()]
    pc = p.pinned_columns(q, cols, dtype=np.int32)
    for k in cols: pc[k][:] = c32[k]
    pc["count"] = pin["count"]
    oe16 = p.pinned(max(total, 1), np.uint16)
    wire = dict(quantity_bits=32, mem_shift=20, node_bits=16, offsets=False)
    print("pack_batch_wire (int32 in, no offsets, uint16 out) %8.1f us" % t(lambda: p.pack_batch(pc, w["algo"], w["mode"], out=(od, oe16), wire=wire)), p.stats())
    for chunk in (8192, 16384, 65536):
        os.environ["GANGPACK_CHUNK_APPS"] = str(chunk)
        p2 = g.GangPacker(0)
        p2.set_snapshot(pn["cpu"], pn["mem"], pn["gpu"], pn["eorder"], pn["eorder"], pn["eoff"], pn["eoff"])
        print("   chunk_apps %6d: %8.1f us" % (chunk, t(lambda: p2.pack_batch(pc, w["algo"], w["mode"], out=(od, oe16), wire=wire))))
        p2.close()
    os.environ.pop("GANGPACK_CHUNK_APPS", None)
pageable = {k: np.array(v) for k, v in pin.items()}
print("pack_batch (pageable in, pinned out) %8.1f us" % t(lambda: p.pack_batch(pageable, w["algo"], w["mode"], out=(od, oe))))
# raw PCIe
h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for nbytes in (1 << 20, 8 << 20, 64 << 20):
    def h2d(): d[:nbytes].copy_(h[:nbytes], non_blocking=True); torch.cuda.synchronize()
    def d2h(): h[:nbytes].copy_(d[:nbytes], non_blocking=True); torch.cuda.synchronize()
    a1, a2 = t(h2d), t(d2h)
    print("PCIe %3d MiB  H2D %7.1f us (%5.1f GB/s)   D2H %7.1f us (%5.1f GB/s)" % (nbytes >> 20, a1, nbytes / a1 / 1e3, a2, nbytes / a2 / 1e3))
def both():
    s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
    with torch.cuda.stream(s1): d[:32 << 20].copy_(h[:32 << 20], non_blocking=True)
    with torch.cuda.stream(s2): h[32 << 20:].copy_(d[32 << 20:], non_blocking=True)
    torch.cuda.synchronize()
print("PCIe duplex 32+32 MiB %7.1f us" % t(both))
for sub in (1000, 10000, 25000, 50000):
    ps = {k: v[:sub] if k != "off" else v[:sub + 1] for k, v in pin.items()}
    print("pack_batch %6d apps %8.1f us" % (sub, t(lambda: p.pack_batch(ps, w["algo"], w["mode"], out=(od, oe)))), p.stats()["pack_kernel_ns"])
p.close()

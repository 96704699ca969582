#!/usr/bin/env python
"""Sweep the host-path knobs of gp_pack_batch (diagnostic)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time, numpy as np, torch
sys.path.insert(0, %r)
import k8s_spark_scheduler_b200 as g
from bench import WORKLOADS, make_workload
w = WORKLOADS["tightly-100k"]; nodes, a, eoff, eorder = make_workload(w, 0)
p = g.GangPacker(0); q = len(a["count"]); total = int(a["off"][-1])
pin = {k: p.pinned(len(a[k]), a[k].dtype) for k in a}
for k in a: pin[k][:] = a[k]
pin.pop("group"); pin.pop("young")
od = p.pinned(q, np.int32); oe = p.pinned(total, np.int32)
p.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], eorder, eorder, eoff, eoff)
for _ in range(5): p.pack_batch(pin, 0, 0, out=(od, oe))
ts = []
for _ in range(int(os.environ.get("REPS", "30"))):
    t0 = time.perf_counter(); p.pack_batch(pin, 0, 0, out=(od, oe)); ts.append(time.perf_counter() - t0)
print("%%-40s %%7.1f us" %% (os.environ.get("TAG"), np.median(ts) * 1e6), p.stats()["pack_kernel_ns"], p.stats()["prep_kernel_ns"])
''' % ROOT
for zc_in in (0, 1000000):
    for chunk in (12288, 24576, 33334, 50000, 100000):
        env = dict(os.environ, GANGPACK_ZC_IN_MAX=str(zc_in), GANGPACK_CHUNK_APPS=str(chunk), TAG=f"zc_in_max={zc_in} chunk={chunk}")
        subprocess.run([sys.executable, "-c", code], env=env)
env = dict(os.environ, GANGPACK_TRACE="1", TAG="trace default", REPS="3")
subprocess.run([sys.executable, "-c", code], env=env)
env = dict(os.environ, GANGPACK_TRACE="1", GANGPACK_ZC_IN_MAX="1000000", TAG="trace zc", REPS="3")
subprocess.run([sys.executable, "-c", code], env=env)

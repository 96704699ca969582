#!/usr/bin/env python
"""gp_potential_nodes (node priority order on the device) at 10 k / 50 k / 500 k nodes: run under
`ncu --metrics gpu__time_duration.sum` for the kernel-only times (the call times incl. the copies are in e2e_breakdown)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import k8s_spark_scheduler_b200 as g  # noqa: E402
from k8s_spark_scheduler_b200 import synth  # noqa: E402

p = g.GangPacker()
for n in (10000, 50000, 500000):
    nodes = synth.make_nodes(n)
    p.potential_nodes(nodes["avail_cpu"], nodes["avail_mem"])
    t0 = time.perf_counter()
    d, e = p.potential_nodes(nodes["avail_cpu"], nodes["avail_mem"])
    print("potential_nodes %d nodes: %.1f us per call (pageable inputs), %d in the order" % (n, (time.perf_counter() - t0) * 1e6, len(e)))
p.close()

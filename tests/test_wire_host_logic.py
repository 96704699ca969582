"""Host-side rules of the compact wire format (what a shim does while marshalling, native.compact_apps): the 32-bit layout is
chosen only when it is EXACT -- memory a whole multiple of 2^mem_shift bytes, every quantity in [0, 2^31) -- so that the
library can rebuild the int64 quantities bit for bit (gp_apps_wire, include/gangpack.h).  No GPU needed."""
import numpy as np

import k8s_spark_scheduler_b200 as g
from k8s_spark_scheduler_b200 import synth


def _apps(n=64, seed=5):
    a = synth.make_apps(n, seed=seed)
    return {k: a[k] for k in ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count")}


def test_compact_layout_is_exact_and_reversible():
    a = _apps()
    c = g.native.compact_apps(a, mem_shift=20)
    assert c is not None
    for k in ("drv_cpu", "drv_gpu", "exe_cpu", "exe_gpu"):
        assert c[k].dtype == np.int32 and np.array_equal(c[k].astype(np.int64), a[k])
    for k in ("drv_mem", "exe_mem"):
        assert c[k].dtype == np.int32 and np.array_equal(c[k].astype(np.int64) << 20, a[k])
    assert c["count"] is a["count"]                       # untouched columns are passed through


def test_compact_layout_refused_when_inexact():
    a = _apps()
    b = dict(a); b["exe_mem"] = a["exe_mem"].copy(); b["exe_mem"][3] += 1            # not a multiple of 1 MiB
    assert g.native.compact_apps(b, 20) is None
    b = dict(a); b["drv_cpu"] = a["drv_cpu"].copy(); b["drv_cpu"][0] = 1 << 31        # does not fit int32
    assert g.native.compact_apps(b, 20) is None
    b = dict(a); b["drv_mem"] = a["drv_mem"].copy(); b["drv_mem"][1] = (1 << 31) << 20   # shifted value does not fit
    assert g.native.compact_apps(b, 20) is None
    b = dict(a); b["exe_cpu"] = a["exe_cpu"].copy(); b["exe_cpu"][2] = -1             # negative: left to the int64 path's validation
    assert g.native.compact_apps(b, 20) is None
    # a finer shift keeps byte-granular requests representable as long as they stay below 2^31 units
    b = dict(a); b["exe_mem"] = np.full_like(a["exe_mem"], 1536); b["drv_mem"] = np.full_like(a["drv_mem"], 512)
    c = g.native.compact_apps(b, 9)
    assert c is not None and (c["exe_mem"] == 3).all() and (c["drv_mem"] == 1).all()


def test_missing_gpu_columns_stay_missing():
    a = _apps()
    del a["drv_gpu"], a["exe_gpu"]
    c = g.native.compact_apps(a, 20)
    assert c is not None and "drv_gpu" not in c and "exe_gpu" not in c

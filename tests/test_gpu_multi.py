"""gp_multi_*: one host process driving several devices (both rows).  The GPU test box may have a single GPU:
the handle then holds several contexts on device 0, which exercises the same partitioning / worker threads / scatter
logic; with more GPUs visible every one of them is used as well.  Results must equal the single-context run bit for bit
(which the other tests compare with the oracle) and, for a FIFO sample, the literal oracle directly."""
import numpy as np
import pytest

from helpers import assert_same_results, literal_fifo

pytestmark = pytest.mark.gpu
KEYS = ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count")


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


@pytest.fixture(scope="module")
def single(gangpack):
    p = gangpack.GangPacker()
    yield p
    p.close()


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_independent(gangpack, single, devices):
    import k8s_spark_scheduler_b200.synth as synth
    nodes = synth.make_nodes(6000, groups=3)
    eoff, eorder = synth.group_orders(nodes)
    apps = synth.make_apps(50001, groups=3)
    a = {k: apps[k] for k in KEYS + ("group",)}
    single.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], eorder, eorder, eoff, eoff)
    m = gangpack.MultiGangPacker(devices)
    try:
        m.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], eorder, eorder, eoff, eoff)
        c32 = gangpack.native.compact_apps(a, 20)
        for algo in (0, 1, 2):
            want = single.pack_batch(a, algo, 0)
            got = m.pack_batch(a, algo, 0)
            assert_same_results(got, want, f"multi int64 algo {algo}")
            if algo != 2:
                got = m.pack_batch(c32, algo, 0, wire=dict(quantity_bits=32, mem_shift=20, node_bits=16, offsets=False))
                assert_same_results((got[0], got[1].astype(np.int32), got[2]), want, f"multi compact algo {algo}")
        # pinned result buffers: every device copies its block straight into them
        total = int(want[2][-1])
        od = m.pinned(50001, np.int32); oe = m.pinned(max(total, 1), np.int32)
        od[:] = -7; oe[:] = -7
        want = single.pack_batch(a, 0, 0)
        got = m.pack_batch(a, 0, 0, out=(od, oe))
        assert_same_results(got, want, "multi pinned out")
        # fewer applications than devices
        tiny = {k: v[:2] for k, v in a.items()}
        assert_same_results(m.pack_batch(tiny, 0, 0), single.pack_batch(tiny, 0, 0), "tiny batch")
    finally:
        m.close()


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("mode", [1, 2])
def test_multi_fifo_groups(gangpack, oracle, single, devices, mode):
    """BASELINE configs[3] shape, scaled down: DA sweep queue, 16 instance groups spread over the devices."""
    import k8s_spark_scheduler_b200.synth as synth
    G = 16
    nodes = synth.make_nodes(4000, groups=G)
    eoff, eorder = synth.group_orders(nodes)
    apps = synth.make_apps(12000, groups=G, da_sweep=True, young_frac=0.05)
    a = {k: apps[k] for k in KEYS + ("group", "young")}
    a["count"] = (a["count"] * 3).astype(np.int32)              # hungry enough to block some groups
    m = gangpack.MultiGangPacker(devices)
    try:
        for algo in (0, 1):
            single.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], eorder, eorder, eoff, eoff)
            want = single.pack_batch(a, algo, mode)
            wsnap = single.get_snapshot()
            m.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], eorder, eorder, eoff, eoff)
            got = m.pack_batch(a, algo, mode)
            assert_same_results(got, want, f"multi fifo algo {algo} mode {mode}")
            gsnap = m.get_snapshot()
            for x, y in zip(gsnap, wsnap):
                assert np.array_equal(x, y)
            owner = m.group_owner()
            assert len(set(owner.tolist())) == min(len(devices), G)       # every device got work
        # one group against the literal restatement
        g = 5
        sel = np.nonzero(apps["group"] == g)[0]
        order = eorder[eoff[g]:eoff[g + 1]]
        sub = {k: np.asarray(v)[sel] for k, v in a.items()}
        (ld, le, loff), _ = literal_fifo(oracle, 1, mode, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order, sub, sub["young"])
        assert np.array_equal(got[0][sel], ld)
        for j, i in enumerate(sel):
            if ld[j] >= 0:
                assert np.array_equal(got[1][got[2][i]:got[2][i + 1]], le[loff[j]:loff[j + 1]])
    finally:
        m.close()

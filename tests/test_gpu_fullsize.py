"""Full-size parity of the CUDA path against the oracle (VERDICT r01 "weak #1"): every one of the 100 000 applications
of the headline workload (tightly-pack and distribute-evenly, BASELINE configs[1]/[2] shape), the 125 000-application
per-GPU share of configs[4] on 50 000 nodes, and >= 5 000-application slices through the LITERAL restatement
(string-keyed maps, loop for loop: pack_tightly.go:45-61, distribute_evenly.go:49-70, binpack.go:60-87)."""
import numpy as np
import pytest

from helpers import assert_same_results, literal_batch, literal_fifo, res_aos

pytestmark = pytest.mark.gpu

KEYS = ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count")


@pytest.fixture(scope="module")
def packer(gangpack):
    p = gangpack.GangPacker()
    yield p
    p.close()


def _closed(oracle, algo, nodes, order, a, threads=16):
    drv = res_aos(a["drv_cpu"], a["drv_mem"], a["drv_gpu"])
    exe = res_aos(a["exe_cpu"], a["exe_mem"], a["exe_gpu"])
    _, dn, en, off, _ = oracle.closed_batch(algo, 0, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order,
                                            drv, exe, a["count"], None, n_threads=threads)
    return dn, en, off


@pytest.mark.parametrize("algo", [0, 1])
def test_headline_all_100k_apps(oracle, packer, algo):
    import k8s_spark_scheduler_b200.synth as synth
    nodes = synth.make_nodes(10000)
    apps = synth.make_apps(100000)
    order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
    a = {k: apps[k] for k in KEYS}
    packer.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order)
    got = packer.pack_batch(a, algo, 0)
    assert_same_results(got, _closed(oracle, algo, nodes, order, a), f"headline 100k closed algo {algo}")
    # the literal restatement on a 6 000-application slice taken from the middle of the queue
    lo, hi = 47000, 53000
    sl = {k: v[lo:hi] for k, v in a.items()}
    want = literal_batch(oracle, algo, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order, sl, n_threads=16)
    off = got[2]
    sub = (got[0][lo:hi], got[1][off[lo]:off[hi]], off[lo:hi + 1] - off[lo])
    assert_same_results(sub, want, f"headline literal slice algo {algo}")


def test_config4_share_125k_apps_50k_nodes(oracle, packer):
    import k8s_spark_scheduler_b200.synth as synth
    nodes = synth.make_nodes(50000)
    apps = synth.make_apps(125000)
    order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
    a = {k: apps[k] for k in KEYS}
    packer.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order)
    got = packer.pack_batch(a, 0, 0)
    assert_same_results(got, _closed(oracle, 0, nodes, order, a), "50k x 125k closed")
    lo, hi = 100000, 105000
    sl = {k: v[lo:hi] for k, v in a.items()}
    want = literal_batch(oracle, 0, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order, sl, n_threads=16)
    off = got[2]
    assert_same_results((got[0][lo:hi], got[1][off[lo]:off[hi]], off[lo:hi + 1] - off[lo]), want, "50k x 125k literal slice")


@pytest.mark.parametrize("algo", [0, 1])
def test_deep_workload_all_apps(oracle, packer, algo):
    """The deep-scan workload of bench.py (`tightly-100k-deep`: cluster ~97 % full, large gangs, many applications that do
    not fit at all): every decision against the closed form, a slice against the literal restatement."""
    import k8s_spark_scheduler_b200.synth as synth
    nodes = synth.make_nodes(10000, fill=(0.95, 1.0))
    apps = synth.make_apps(40000, deep=True)
    order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
    a = {k: apps[k] for k in KEYS}
    packer.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order)
    got = packer.pack_batch(a, algo, 0)
    want = _closed(oracle, algo, nodes, order, a)
    assert_same_results(got, want, f"deep closed algo {algo}")
    nofit = float((want[0] < 0).mean())
    assert 0.05 < nofit < 0.95, nofit          # the workload really mixes fitting and non-fitting gangs
    # the literal restatement pays the reference's O(driver candidates x executor nodes) for every gang that fits nowhere
    # (~0.1 s per application and core here): a 240-application slice keeps the suite inside a few minutes
    L = 240
    sl = {k: v[:L] for k, v in a.items()}
    lit = literal_batch(oracle, algo, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order, sl, n_threads=32)
    off = got[2]
    assert (lit[0] < 0).any() and (lit[0] >= 0).any()
    assert_same_results((got[0][:L], got[1][:off[L]], off[:L + 1]), lit, f"deep literal slice algo {algo}")


def test_fifo_then_independent_pack_sees_charged_snapshot(oracle, packer):
    """the driver's own pack follows fitEarlierDrivers on the SAME metadata (resource.go:255 then :321).
    A FIFO batch mutates the device snapshot; a following gp_pack_one / independent batch (all three packers, i.e. also
    the compact 32-bit view) must see the charged availability."""
    import k8s_spark_scheduler_b200.synth as synth
    nodes = synth.make_nodes(3000)
    order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
    queue = synth.make_apps(400, seed=11)
    q = {k: queue[k] for k in KEYS}
    young = np.ones(400, np.uint8)
    for mode in (1, 2):
        packer.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order)
        packer.pack_batch({**q, "young": young}, 0, mode)
        (_, _, _), final = literal_fifo(oracle, 0, mode, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order, q, young)
        fc, fm, fg = packer.get_snapshot()
        assert np.array_equal(fc, final[0]) and np.array_equal(fm, final[1])
        assert not np.array_equal(fc, nodes["avail_cpu"])            # the queue really consumed something
        mine = synth.make_apps(300, seed=12)
        m = {k: mine[k] for k in KEYS}
        for algo, oalgo in ((0, 0), (1, 1), (2, 4)):
            drv = res_aos(m["drv_cpu"], m["drv_mem"], m["drv_gpu"]); exe = res_aos(m["exe_cpu"], m["exe_mem"], m["exe_gpu"])
            _, wd, we, woff, _ = oracle.closed_batch(oalgo, 0, final[0], final[1], final[2], order, order, drv, exe, m["count"], None, n_threads=8)
            got = packer.pack_batch(m, algo, 0)
            assert_same_results(got, (wd, we, woff), f"after fifo mode {mode} algo {algo}")
        ok, d, ex = packer.pack_one(0, (int(m["drv_cpu"][0]), int(m["drv_mem"][0]), 0), (int(m["exe_cpu"][0]), int(m["exe_mem"][0]), 0), int(m["count"][0]))
        _, wd0, we0, woff0, _ = oracle.closed_batch(0, 0, final[0], final[1], final[2], order, order, drv[:1], exe[:1], m["count"][:1], None)
        assert ok == (wd0[0] >= 0)
        if ok:
            assert d == wd0[0] and np.array_equal(ex, we0[:woff0[1]])

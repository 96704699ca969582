"""The compact wire formats (gp_pack_batch_wire: int32 quantities, device-derived offsets, uint16 node indices) give the
same bits as the int64 layout, and the two decision paths of independent tightly-pack / distribute-evenly -- per-shape
capacity tables (default) and the node-order scan (GANGPACK_TABLES=0, shape overflow, counts above the table clamp,
multi-round distribute-evenly) -- agree with the oracle and with each other."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import assert_same_results, literal_batch, random_apps, random_cluster, res_aos

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count")


@pytest.fixture(scope="module")
def packer(gangpack):
    p = gangpack.GangPacker()
    yield p
    p.close()


def _closed(oracle, algo, cpu, mem, gpu, drv_idx, exec_idx, a, threads=8):
    drv = res_aos(a["drv_cpu"], a["drv_mem"], a["drv_gpu"]); exe = res_aos(a["exe_cpu"], a["exe_mem"], a["exe_gpu"])
    _, dn, en, off, _ = oracle.closed_batch(algo, 0, cpu, mem, gpu, drv_idx, exec_idx, drv, exe, a["count"], None, n_threads=threads)
    return dn, en, off


@pytest.mark.parametrize("q", [1, 700, 70000])
def test_wire_formats_bit_identical(gangpack, oracle, packer, q):
    """int64 layout vs (int32 millicores / MiB, NULL offsets, uint16 node indices): pageable and pinned, the small-batch
    route (inputs gathered by one kernel, results written into mapped memory) and the pipelined DMA route."""
    import k8s_spark_scheduler_b200.synth as synth
    nodes = synth.make_nodes(3000)
    order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
    packer.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order)
    apps = synth.make_apps(q, gpu_variant=True)
    a = {k: apps[k] for k in KEYS}
    compact = gangpack.native.compact_apps(a, mem_shift=20)
    assert compact is not None
    for algo in (0, 1):
        ref = packer.pack_batch(a, algo, 0)
        if q <= 700:
            assert_same_results(ref, _closed(oracle, algo, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order, a), f"int64 q={q}")
        for wire in (dict(quantity_bits=32, mem_shift=20, node_bits=16, offsets=False),
                     dict(quantity_bits=32, mem_shift=20, node_bits=32, offsets=True),
                     dict(quantity_bits=64, node_bits=16, offsets=False)):
            src = compact if wire["quantity_bits"] == 32 else a
            got = packer.pack_batch(src, algo, 0, wire=wire)
            assert_same_results((got[0], got[1].astype(np.int32), got[2]), ref, f"pageable {wire} q={q} algo {algo}")
            # pinned buffers (what the shim uses)
            qdt = np.int32 if wire["quantity_bits"] == 32 else np.int64
            pin = packer.pinned_columns(q, [k for k in KEYS if k != "count"], dtype=qdt)
            for k in pin:
                pin[k][:] = src[k]
            pin["count"] = packer.pinned(q, np.int32); pin["count"][:] = a["count"]
            total = int(ref[2][-1])
            od = packer.pinned(q, np.int32); od[:] = -7
            oe = packer.pinned(max(total, 1), np.uint16 if wire["node_bits"] == 16 else np.int32); oe[:] = 7
            got = packer.pack_batch(pin, algo, 0, out=(od, oe), wire=wire)
            goff = got[2] if got[2] is not None else ref[2]      # device-derived offsets are not returned on the caller-buffer path
            assert_same_results((got[0], got[1][:total].astype(np.int32), goff), ref, f"pinned {wire} q={q} algo {algo}")


def test_wire_format_rejections(gangpack, packer):
    Gi = 1 << 30
    idx = np.arange(2, dtype=np.int32)
    packer.set_snapshot([8000, 8000], [8 * Gi, 8 * Gi], [0, 0], idx, idx)
    base = {"drv_cpu": [1000], "drv_mem": [Gi], "drv_gpu": [0], "exe_cpu": [1000], "exe_mem": [Gi], "exe_gpu": [0], "count": [1]}
    from k8s_spark_scheduler_b200 import GangpackError
    with pytest.raises(GangpackError):           # 16-bit node indices: independent tightly/evenly only
        packer.pack_batch(base, 2, 0, wire=dict(quantity_bits=64, node_bits=16))
    with pytest.raises(GangpackError):
        packer.pack_batch({**base, "young": [0]}, 0, 1, wire=dict(quantity_bits=64, node_bits=16))
    with pytest.raises(GangpackError):           # NULL offsets: not for the FIFO modes
        packer.pack_batch({**base, "young": [0]}, 0, 1, wire=dict(quantity_bits=64, node_bits=32, offsets=False))
    with pytest.raises(GangpackError):
        packer.pack_batch(base, 0, 0, wire=dict(quantity_bits=48))
    # a negative int32 quantity is a validation error, not a silently shifted value
    neg = {k: np.asarray(v, np.int32) for k, v in base.items()}
    neg["exe_mem"] = np.array([-1], np.int32)
    with pytest.raises(GangpackError) as e:
        packer.pack_batch(neg, 0, 0, wire=dict(quantity_bits=32, mem_shift=20))
    assert e.value.status == 1
    # 32-bit inputs work in the FIFO modes and for minimal-fragmentation too (int32 results)
    c = gangpack.native.compact_apps(base, 20)
    for algo, mode in ((0, 1), (1, 2), (2, 0)):
        packer.set_snapshot([8000, 8000], [8 * Gi, 8 * Gi], [0, 0], idx, idx)
        got = packer.pack_batch({**c, "young": [0]} if mode else c, algo, mode, wire=dict(quantity_bits=32, mem_shift=20, node_bits=32))
        packer.set_snapshot([8000, 8000], [8 * Gi, 8 * Gi], [0, 0], idx, idx)
        want = packer.pack_batch({**base, "young": [0]} if mode else base, algo, mode)
        assert_same_results(got, want, f"32-bit inputs algo {algo} mode {mode}")


def test_many_shapes_clamp_and_rounds(oracle, packer):
    """Inputs the tables cannot answer alone: > 64 distinct executor shapes in one batch (hash overflow -> scan path for
    the rest), an executor count above the table clamp of a 50k-node group, distribute-evenly placements that need
    several rounds; all mixed with table-path applications in the same launch."""
    import k8s_spark_scheduler_b200.synth as synth
    rng = np.random.default_rng(99)
    nodes = synth.make_nodes(50000)
    order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
    packer.set_snapshot(nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order)
    q = 3000
    apps = synth.make_apps(q)
    a = {k: apps[k].copy() for k in KEYS}
    a["exe_mem"] = a["exe_mem"] + rng.integers(0, 300, q) * (1 << 20)          # ~300 x 3 distinct shapes
    a["count"][5] = 90000                                                       # > clamp (2^32 / 50k = 85 899)
    a["exe_cpu"][5] = 250; a["exe_mem"][5] = 1 << 28
    for algo in (0, 1):
        got = packer.pack_batch(a, algo, 0)
        assert_same_results(got, _closed(oracle, algo, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"], order, order, a), f"many shapes algo {algo}")
        st = packer.stats()
        assert 0 < st["scan_path_apps"] < q, st            # both paths ran in this launch
        assert got[0][5] >= 0
    # several rounds: a small, nearly full cluster
    cpu, mem, gpu = random_cluster(rng, 60, tight=True)
    idx = np.arange(60, dtype=np.int32)
    packer.set_snapshot(cpu, mem, gpu, idx, idx)
    b = random_apps(rng, 400)
    b["exe_cpu"][:] = 500; b["exe_mem"][:] = 1 << 29
    b["count"] = rng.integers(20, 90, 400).astype(np.int32)
    got = packer.pack_batch(b, 1, 0)
    want = _closed(oracle, 1, cpu, mem, gpu, idx, idx, b)
    assert_same_results(got, want, "multi-round evenly")
    fits = want[0] >= 0
    multi = [i for i in np.nonzero(fits)[0] if len(np.unique(want[1][want[2][i]:want[2][i + 1]])) < b["count"][i]]
    assert len(multi) > 10                                   # the case really occurs
    lit = literal_batch(oracle, 1, cpu, mem, gpu, idx, idx, b)
    assert_same_results(got, lit, "multi-round evenly, literal")


def test_tables_multi_group_and_driver_only_nodes(oracle, packer):
    """Table path with instance groups, driver candidates that are not executor candidates, the gpu dimension, negative
    availability, zero-request dimensions -- few shapes so that (nearly) every application takes the tables."""
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = int(rng.integers(40, 400))
        cpu, mem, gpu = random_cluster(rng, n, tight=bool(trial % 2), gpus=bool(trial % 3 == 0), negative=bool(trial % 2))
        perm = rng.permutation(n)
        exec_idx = perm[rng.random(n) < 0.8].astype(np.int32)
        drv_idx = rng.permutation(n)[: max(1, int(n * 0.6))].astype(np.int32)
        q = 600
        b = random_apps(rng, q, gpus=bool(trial % 3 == 0), zero_dims=bool(trial == 4))
        shapes = rng.integers(0, 5, q)                          # five executor shapes
        b["exe_cpu"] = np.array([250, 500, 1000, 2000, 0 if trial == 4 else 750], np.int64)[shapes]
        b["exe_mem"] = np.array([1 << 28, 3 << 28, (1 << 30) + 7, 1 << 31, 5 << 27], np.int64)[shapes]
        packer.set_snapshot(cpu, mem, gpu, exec_idx, drv_idx)
        for algo in (0, 1):
            got = packer.pack_batch(b, algo, 0)
            assert_same_results(got, _closed(oracle, algo, cpu, mem, gpu, drv_idx, exec_idx, b), f"trial {trial} algo {algo}")
            if trial < 2:
                assert_same_results(got, literal_batch(oracle, algo, cpu, mem, gpu, drv_idx, exec_idx, b), f"literal trial {trial} algo {algo}")
        st = packer.stats()
        assert st["scan_path_apps"] < q


def test_scan_path_forced():
    """GANGPACK_TABLES=0: every independent decision takes the node-order scan (the kernel reported as 'scan' in bench.py):
    the parity tests of the independent mode must pass on that path as well."""
    env = dict(os.environ, GANGPACK_TABLES="0")
    tests = ["tests/test_gpu_parity.py::test_random_independent", "tests/test_gpu_parity.py::test_awkward_divisors",
             "tests/test_gpu_parity.py::test_synthetic_bench_workload_sample", "tests/test_gpu_parity.py::test_golden_pack_cases",
             "tests/test_gpu_fullsize.py::test_headline_all_100k_apps", "tests/test_gpu_fullsize.py::test_deep_workload_all_apps",
             "tests/test_gpu_wire_and_tables.py::test_wire_formats_bit_identical"]
    p = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x"] + tests, env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]


@pytest.mark.parametrize("n", [4095, 4097, 50000, 500000])
def test_node_order_large(packer, n):
    """PotentialNodes (internal/sort/nodesorting.go:83-122) at sizes that span one tile, a tile boundary, BASELINE
    configs[4] and 10x that: the device order (tile sort in shared memory + binary-search ranks) must equal a host
    lexicographic sort on (zone priority, memory, cpu, name) with heavy ties; label-rank re-sort on the candidate lists."""
    rng = np.random.default_rng(n)
    Z = 3
    zone = rng.integers(0, Z, n).astype(np.int32)
    cpu = (rng.integers(0, 40, n) * 500).astype(np.int64) - (rng.integers(0, 50, n) == 0) * 3000
    mem = (rng.integers(0, 64, n) << 30).astype(np.int64)
    rank = rng.permutation(n).astype(np.int32)                      # rank of the node name
    cand = (rng.random(n) < 0.7).astype(np.uint8)
    unsched = (rng.random(n) < 0.05).astype(np.uint8)
    ready = (rng.random(n) < 0.97).astype(np.uint8)
    lrd = rng.integers(-1, 4, n).astype(np.int32)
    d, e = packer.potential_nodes(cpu, mem, zone, Z, rank, cand, unsched, ready, lrd, None)
    # host truth
    tot_m = np.array([mem[zone == z].sum() for z in range(Z)]); tot_c = np.array([cpu[zone == z].sum() for z in range(Z)])
    zorder = sorted(range(Z), key=lambda z: (tot_m[z], tot_c[z], z))
    zprio = np.empty(Z, np.int64); zprio[zorder] = np.arange(Z)
    order = np.lexsort((rank, cpu, mem, zprio[zone]))
    drv = order[cand[order] != 0]
    key = np.where(lrd[drv] < 0, 1 << 30, lrd[drv])
    drv = drv[np.argsort(key, kind="stable")]
    exe = order[(unsched[order] == 0) & (ready[order] != 0)]
    assert np.array_equal(d, drv.astype(np.int32))
    assert np.array_equal(e, exe.astype(np.int32))


def test_undefined_ties_are_reported(packer):
    """equal (zone priority, memory, cpu) with different gpu is 'not less' both ways in the reference's
    comparator (nodesorting.go:83-93) -> sort.Slice may order the pair either way.  The device orders by name and says so."""
    cpu = np.array([4000, 4000, 4000, 8000], np.int64); mem = np.array([1 << 33] * 3 + [1 << 34], np.int64)
    gpu = np.array([0, 1, 0, 0], np.int64)
    d, e = packer.potential_nodes(cpu, mem, avail_gpu=gpu)
    assert list(d) == [0, 1, 2, 3] and packer.undefined_ties == 2          # (n0,n1) and (n1,n2) differ in gpu only
    d, e = packer.potential_nodes(cpu, mem, avail_gpu=np.zeros(4, np.int64))
    assert packer.undefined_ties == 0
    d, e = packer.potential_nodes(cpu, mem)
    assert packer.undefined_ties == 0                                        # no gpu column: nothing to detect

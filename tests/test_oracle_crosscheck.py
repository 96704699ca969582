"""Three independent statements of the algorithm must agree on randomized inputs:
pure-Python literal (oracle/pyref.py) == literal C (string maps) == closed-form C (int64 SoA).
Inputs include negative availability, zero-request dimensions, GPU dimension, driver candidates that
are not executor candidates, orders naming nodes missing from the metadata, count == 0."""
import numpy as np
import pytest

from helpers import random_apps, random_cluster, res_aos
from oracle import pyref

ALGOS = [(0, "tightly-pack"), (1, "distribute-evenly"), (4, "minimal-fragmentation")]


def _orders(rng, n, names):
    perm = rng.permutation(n)
    exec_idx = perm[rng.random(n) < 0.85]
    drv_idx = rng.permutation(n)[: max(1, int(n * rng.uniform(0.3, 1.0)))]
    return exec_idx.astype(np.int32), drv_idx.astype(np.int32)


@pytest.mark.parametrize("seed", range(8))
def test_pack_three_way(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    for trial in range(6):
        n = int(rng.integers(1, 40))
        cpu, mem, gpu = random_cluster(rng, n, tight=bool(trial % 2), gpus=bool(seed % 2), negative=bool(seed % 3 == 0))
        names = ["node-%03d" % i for i in range(n)]
        exec_idx, drv_idx = _orders(rng, n, names)
        apps = random_apps(rng, 24, gpus=bool(seed % 2), zero_dims=bool(seed % 4 == 1), big_counts=bool(seed % 4 == 2))
        # the literal oracles also see names that are NOT in the metadata
        exec_names = [names[i] for i in exec_idx] + ["ghost-e"]
        drv_names = ["ghost-d"] + [names[i] for i in drv_idx]
        cl = oracle.Cluster(names, cpu, mem, gpu)
        drv = res_aos(apps["drv_cpu"], apps["drv_mem"], apps["drv_gpu"])
        exe = res_aos(apps["exe_cpu"], apps["exe_mem"], apps["exe_gpu"])
        meta = {names[i]: (int(cpu[i]), int(mem[i]), int(gpu[i])) for i in range(n)}
        for algo_id, algo in ALGOS:
            ld, le, off = cl.binpack_batch(algo_id, drv, exe, apps["count"], drv_names, exec_names, n_threads=2)
            _, cd, ce, coff, _ = oracle.closed_batch(algo_id, 0, cpu, mem, gpu, drv_idx, exec_idx, drv, exe, apps["count"])
            assert np.array_equal(off, coff)
            assert np.array_equal(ld, cd), (seed, trial, algo)
            for q in range(len(apps["count"])):
                d, ex, ok = pyref.spark_bin_pack(tuple(int(x) for x in drv[q]), tuple(int(x) for x in exe[q]),
                                                 int(apps["count"][q]), drv_names, exec_names, dict(meta),
                                                 pyref.DISTRIBUTORS[algo])
                assert ok == (ld[q] >= 0), (seed, trial, algo, q)
                if ok:
                    assert names[ld[q]] == d
                    assert [names[i] for i in le[off[q]:off[q + 1]]] == ex, (seed, trial, algo, q)
                    assert np.array_equal(le[off[q]:off[q + 1]], ce[off[q]:off[q + 1]]), (seed, trial, algo, q)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_fifo_three_way(oracle, seed, mode):
    rng = np.random.default_rng(2000 + seed)
    mode_id = {"reference": 1, "exact": 2}[mode]
    for trial in range(4):
        n = int(rng.integers(2, 30))
        cpu, mem, gpu = random_cluster(rng, n, gpus=bool(seed % 2))
        names = ["node-%03d" % i for i in range(n)]
        exec_idx, drv_idx = _orders(rng, n, names)
        q = 30
        apps = random_apps(rng, q, gpus=bool(seed % 2), zero_dims=bool(seed % 3 == 1))
        young = (rng.random(q) < 0.3).astype(np.uint8)
        drv = res_aos(apps["drv_cpu"], apps["drv_mem"], apps["drv_gpu"])
        exe = res_aos(apps["exe_cpu"], apps["exe_mem"], apps["exe_gpu"])
        exec_names = [names[i] for i in exec_idx]
        drv_names = [names[i] for i in drv_idx]
        for algo_id, algo in ALGOS:
            cl = oracle.Cluster(names, cpu, mem, gpu)
            lb, ld, le, off = cl.fifo(algo_id, mode_id, drv, exe, apps["count"], young, drv_names, exec_names)
            cb, cd, ce, coff, cfinal = oracle.closed_batch(algo_id, mode_id, cpu, mem, gpu, drv_idx, exec_idx, drv, exe,
                                                           apps["count"], young)
            assert lb == cb and np.array_equal(ld, cd), (seed, trial, algo)
            lfinal = cl.available()
            for a, b in zip(lfinal, cfinal):
                assert np.array_equal(a, b), (seed, trial, algo)
            meta = {names[i]: (int(cpu[i]), int(mem[i]), int(gpu[i])) for i in range(n)}
            py_apps = [{"drv": tuple(int(x) for x in drv[i]), "exe": tuple(int(x) for x in exe[i]),
                        "count": int(apps["count"][i]), "young": bool(young[i])} for i in range(q)]
            pb, pres = pyref.fit_earlier_drivers(py_apps, drv_names, exec_names, meta, algo, mode)
            assert pb == lb
            for i, (d, ex) in enumerate(pres):
                if d == "unevaluated":
                    assert ld[i] == -2
                elif d is None:
                    assert ld[i] == -1
                else:
                    assert names[ld[i]] == d
                    assert [names[j] for j in le[off[i]:off[i + 1]]] == ex
                    assert np.array_equal(le[off[i]:off[i + 1]], ce[off[i]:off[i + 1]])
            for i in range(n):
                assert meta[names[i]] == (lfinal[0][i], lfinal[1][i], lfinal[2][i])


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_fifo_with_zone_aware_packers_two_way(oracle, seed, mode):
    """fitEarlierDrivers (resource.go:224-262) with the zone-aware packers as BinpackFunc: the literal C loop (orc_fifo over
    single_az.go / az_aware_pack_tightly.go / single_az_minimal_fragmentation.go) against the pure-Python statement of the
    same loop -- placements, blocking and the availability left behind.  This is the checker of gp_pack_fifo_zones."""
    rng = np.random.default_rng(2600 + seed)
    mode_id = {"reference": 1, "exact": 2}[mode]
    fns = ((2, pyref.single_az_tightly_pack), (3, pyref.az_aware_tightly_pack), (5, pyref.single_az_minimal_fragmentation))
    for trial in range(3):
        n = int(rng.integers(6, 30))
        names = ["node-%03d" % i for i in range(n)]
        sched_cpu = rng.integers(2, 17, n) * 1000; sched_mem = rng.integers(2, 33, n) * (1 << 30); sched_gpu = rng.integers(0, 3, n) * (seed % 2)
        cpu = (sched_cpu * rng.uniform(0.3, 1, n)).astype(np.int64) // 250 * 250
        mem = (sched_mem * rng.uniform(0.3, 1, n)).astype(np.int64) // (1 << 28) * (1 << 28)
        gpu = np.minimum(sched_gpu, rng.integers(0, 3, n))
        zones = {nm: "z%d" % rng.integers(0, 3) for nm in names}
        exec_idx, drv_idx = _orders(rng, n, names)
        exec_names = [names[i] for i in exec_idx]; drv_names = [names[i] for i in drv_idx]
        q = 24
        apps = random_apps(rng, q, gpus=bool(seed % 2))
        apps["count"] = np.minimum(apps["count"], 6).astype(np.int32)
        young = (rng.random(q) < 0.8).astype(np.uint8)
        drv = res_aos(apps["drv_cpu"], apps["drv_mem"], apps["drv_gpu"]); exe = res_aos(apps["exe_cpu"], apps["exe_mem"], apps["exe_gpu"])
        sched = {names[i]: (int(sched_cpu[i]), int(sched_mem[i]), int(sched_gpu[i])) for i in range(n)}
        for algo_id, fn in fns:
            cl = oracle.Cluster(names, cpu, mem, gpu, sched=(sched_cpu, sched_mem, sched_gpu), zone=[zones[nm] for nm in names])
            lb, ld, le, off = cl.fifo(algo_id, mode_id, drv, exe, apps["count"], young, drv_names, exec_names, with_efficiencies=True)
            meta = {names[i]: (int(cpu[i]), int(mem[i]), int(gpu[i])) for i in range(n)}
            blocked = -1
            for i in range(q):
                if blocked >= 0:
                    assert ld[i] == -2, (seed, trial, algo_id, i)
                    continue
                d_req = tuple(int(x) for x in drv[i]); e_req = tuple(int(x) for x in exe[i])
                d, ex, ok = fn(d_req, e_req, int(apps["count"][i]), drv_names, exec_names, meta, sched, zones)
                if not ok:
                    assert ld[i] == -1, (seed, trial, algo_id, i)
                    if not young[i]:
                        blocked = i
                    continue
                assert names[ld[i]] == d and [names[j] for j in le[off[i]:off[i + 1]]] == ex, (seed, trial, algo_id, i)
                if mode == "reference":
                    for nm, u in pyref.spark_resource_usage(d_req, e_req, d, ex).items():
                        if nm in meta:
                            meta[nm] = pyref.sub(meta[nm], u)
                else:
                    meta[d] = pyref.sub(meta[d], d_req)
                    for nm in ex:
                        meta[nm] = pyref.sub(meta[nm], e_req)
            assert lb == blocked, (seed, trial, algo_id)
            final = cl.available()
            for i in range(n):
                assert meta[names[i]] == (final[0][i], final[1][i], final[2][i]), (seed, trial, algo_id, names[i])


def test_synthetic_workload_oracles_agree(oracle):
    """The bench workload shape at reduced size: literal (threaded) == closed form."""
    import k8s_spark_scheduler_b200.synth as synth
    nodes = synth.make_nodes(600)
    apps = synth.make_apps(300)
    order = synth.priority_order(nodes["avail_cpu"], nodes["avail_mem"])
    names = synth.node_names(600)
    cl = oracle.Cluster(names, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"])
    onames = [names[i] for i in order]
    drv = res_aos(apps["drv_cpu"], apps["drv_mem"], apps["drv_gpu"])
    exe = res_aos(apps["exe_cpu"], apps["exe_mem"], apps["exe_gpu"])
    for algo in (0, 1):
        ld, le, off = cl.binpack_batch(algo, drv, exe, apps["count"], onames, onames, n_threads=4)
        _, cd, ce, coff, _ = oracle.closed_batch(algo, 0, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"],
                                                 order, order, drv, exe, apps["count"], n_threads=4)
        assert np.array_equal(ld, cd) and np.array_equal(le, ce)
        assert (ld >= 0).mean() > 0.5


def _doc_cluster():
    """the example of LIB/binpack/minimal_fragmentation.go:43-44: capacities a1 b1 c3 d5 e5 f17 for a 1-cpu executor"""
    names = ["a", "b", "c", "d", "e", "f"]
    caps = [1, 1, 3, 5, 5, 17]
    cpu = np.array([1000 * c for c in caps], dtype=np.int64)
    mem = np.full(6, 1 << 40, dtype=np.int64)
    gpu = np.zeros(6, dtype=np.int64)
    return names, cpu, mem, gpu


# (executorCount, expected ExecutorNodes) -- the doc comment of minimalFragmentation (minimal_fragmentation.go:43-58).
# Its last example (count 19 -> [f x17, a, b]) contradicts the code below it: after f is consumed the remainder 2
# goes to the first node with capacity >= 2, which is c (internalMinimalFragmentation :107-114) -> the code wins.
MINFRAG_DOC = [
    (11, ["d"] * 5 + ["e"] * 5 + ["a"]),
    (6, ["d"] * 5 + ["a"]),
    (15, ["d"] * 5 + ["e"] * 5 + ["c"] * 3 + ["a", "b"]),
    (17, ["f"] * 17),
    (19, ["f"] * 17 + ["c", "c"]),
]


@pytest.mark.parametrize("count,expected", MINFRAG_DOC)
def test_minimal_fragmentation_doc_examples(oracle, count, expected):
    names, cpu, mem, gpu = _doc_cluster()
    exe = (1000, 1, 0)
    meta = {n: (int(cpu[i]), int(mem[i]), 0) for i, n in enumerate(names)}
    nodes, ok = pyref.minimal_fragmentation(exe, count, names, meta, {})
    assert ok and nodes == expected
    # through SparkBinPack with a driver that only fits on a node outside the executor order
    names2 = names + ["drv"]
    cl = oracle.Cluster(names2, np.append(cpu, 500), np.append(mem, 1 << 40), np.append(gpu, 0))
    has, d, ex, _ = cl.binpack(4, (500, 1, 0), exe, count, ["drv"], names)
    assert has and d == "drv" and ex == expected
    drv = res_aos([500], [1], [0]); exa = res_aos([1000], [1], [0])
    _, cd, ce, _, _ = oracle.closed_batch(4, 0, np.append(cpu, 500), np.append(mem, 1 << 40), np.append(gpu, 0),
                                          [6], list(range(6)), drv, exa, [count])
    assert cd[0] == 6 and [names2[i] for i in ce[:count]] == expected


@pytest.mark.parametrize("seed", range(10))
def test_minimal_fragmentation_three_way_dense(oracle, seed):
    """small capacities with many ties, counts around the interesting thresholds, unlimited capacities"""
    rng = np.random.default_rng(7000 + seed)
    for trial in range(40):
        n = int(rng.integers(1, 24))
        caps = rng.integers(0, 9 if trial % 2 else 30, size=n)
        cpu = (caps * 1000 + rng.integers(0, 1000, size=n)).astype(np.int64)
        mem = np.full(n, 1 << 40, dtype=np.int64)
        gpu = np.zeros(n, dtype=np.int64)
        if seed % 5 == 4:
            cpu[rng.integers(0, n)] = -5
        names = ["n%02d" % i for i in range(n)]
        exec_idx = rng.permutation(n)[: max(1, int(n * rng.uniform(0.5, 1.0)))].astype(np.int32)
        drv_idx = rng.permutation(n)[: max(1, int(n * rng.uniform(0.2, 1.0)))].astype(np.int32)
        q = 12
        count = rng.integers(0, max(2, int(caps.sum()) + 3), size=q).astype(np.int32)
        exe_cpu = np.full(q, 1000, dtype=np.int64)
        exe_mem = np.ones(q, dtype=np.int64)
        if seed % 5 == 3:
            exe_cpu[::3] = 0; exe_mem[::3] = 0          # every dimension zero: capacity math.MaxInt, target wraps
        drv_cpu = rng.integers(0, 4000, size=q).astype(np.int64)
        drv = res_aos(drv_cpu, np.ones(q, dtype=np.int64), np.zeros(q, dtype=np.int64))
        exe = res_aos(exe_cpu, exe_mem, np.zeros(q, dtype=np.int64))
        cl = oracle.Cluster(names, cpu, mem, gpu)
        exec_names = [names[i] for i in exec_idx]; drv_names = [names[i] for i in drv_idx]
        ld, le, off = cl.binpack_batch(4, drv, exe, count, drv_names, exec_names)
        _, cd, ce, coff, _ = oracle.closed_batch(4, 0, cpu, mem, gpu, drv_idx, exec_idx, drv, exe, count)
        assert np.array_equal(ld, cd), (seed, trial)
        meta = {names[i]: (int(cpu[i]), int(mem[i]), 0) for i in range(n)}
        for a in range(q):
            d, ex, ok = pyref.spark_bin_pack(tuple(int(x) for x in drv[a]), tuple(int(x) for x in exe[a]), int(count[a]),
                                             drv_names, exec_names, dict(meta), pyref.minimal_fragmentation)
            assert ok == (ld[a] >= 0), (seed, trial, a)
            if ok:
                assert names[ld[a]] == d
                assert [names[i] for i in le[off[a]:off[a + 1]]] == ex, (seed, trial, a)
                assert np.array_equal(le[off[a]:off[a + 1]], ce[off[a]:off[a + 1]]), (seed, trial, a, count[a])


@pytest.mark.parametrize("algo_id", [0, 1, 4])
def test_placements_respect_capacity_and_shape(oracle, algo_id):
    """size-independent properties of every packer on mid-size clusters (closed-form oracle, the GPU suite checks the same on
    the device): count executors, every node's load fits its availability with the driver on it, tightly-pack is node-major
    in priority order, distribute-evenly is round-major without repeats inside a round, minimal-fragmentation fills
    every node but the last one it uses to capacity"""
    rng = np.random.default_rng(31337 + algo_id)
    for trial in range(12):
        n = int(rng.integers(50, 400))
        cpu, mem, gpu = random_cluster(rng, n, tight=bool(trial % 2))
        order = rng.permutation(n).astype(np.int32)
        apps = random_apps(rng, 64, big_counts=True)
        drv = res_aos(apps["drv_cpu"], apps["drv_mem"], apps["drv_gpu"])
        exe = res_aos(apps["exe_cpu"], apps["exe_mem"], apps["exe_gpu"])
        _, dn, en, off, _ = oracle.closed_batch(algo_id, 0, cpu, mem, gpu, order, order, drv, exe, apps["count"])
        pos = np.empty(n, np.int64); pos[order] = np.arange(n)
        for a in range(len(dn)):
            if dn[a] < 0:
                continue
            ex = en[off[a]:off[a + 1]]
            assert len(ex) == apps["count"][a]
            load = np.bincount(ex, minlength=n)
            for d, (node_av, e_req, d_req) in enumerate(((cpu, exe[a][0], drv[a][0]), (mem, exe[a][1], drv[a][1]))):
                used = load * e_req
                used[dn[a]] += d_req
                assert (used[load > 0] <= node_av[load > 0]).all() and used[dn[a]] <= node_av[dn[a]], (trial, a, d)
            if len(ex) == 0:
                continue
            if algo_id == 0:
                assert (np.diff(pos[ex]) >= 0).all(), (trial, a)
            elif algo_id == 1:
                # round-major: inside a round the priority positions increase and no node appears twice
                rounds = np.split(ex, np.nonzero(np.diff(pos[ex]) < 0)[0] + 1)
                assert all(len(set(r)) == len(r) for r in rounds), (trial, a)
            else:
                # consumed nodes appear as runs; every run but the last is a full node (no room for one more executor)
                runs = [(ex[i], int(load[ex[i]])) for i in range(len(ex)) if i == 0 or ex[i] != ex[i - 1]]
                assert len({r[0] for r in runs}) == len(runs), (trial, a)
                for node, cnt in runs[:-1]:
                    extra_cpu = (cnt + 1) * exe[a][0] + (drv[a][0] if node == dn[a] else 0)
                    extra_mem = (cnt + 1) * exe[a][1] + (drv[a][1] if node == dn[a] else 0)
                    assert extra_cpu > cpu[node] or extra_mem > mem[node], (trial, a, node)

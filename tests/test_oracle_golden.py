"""The CPU oracles against every golden vector: what the reference's own tests pin for this path
(fit/no-fit, min-executor semantics, node priority orders -- cited per case in the fixture) and the
hand-derived vectors of SURVEY App. A.5."""
import numpy as np
import pytest

from helpers import ALGO_ID, MODE_ID, case_arrays, order_indices, res_aos


def test_fixture_is_current(golden):
    """tests/gen_golden.py (pyref, hand-derived expectations asserted inside) reproduces the fixture."""
    import json, os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "tests", "gen_golden.py")).read()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "tests", "golden"))
        # run the generator against a scratch ROOT that sees the real oracle package
        code = src.replace('ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', f'ROOT = {td!r}')
        code = code.replace("sys.path.insert(0, ROOT)", f"sys.path.insert(0, {root!r})")
        subprocess.check_call([sys.executable, "-c", code], stdout=subprocess.DEVNULL)
        regenerated = json.load(open(os.path.join(td, "tests", "golden", "hotpath_vectors.json")))
    assert regenerated == golden


@pytest.mark.parametrize("algo", ["tightly-pack", "distribute-evenly"])
def test_literal_oracle_pack_cases(golden, oracle, algo):
    for case in golden["pack_cases"]:
        names, cpu, mem, gpu = case_arrays(case["nodes"])
        cl = oracle.Cluster(names, cpu, mem, gpu)
        app = case["app"]
        ok, d, ex, _ = cl.binpack(ALGO_ID[algo], app["drv"], app["exe"], app["count"],
                                  case["driver_order"], case["exec_order"])
        exp = case["expect"][algo]
        assert ok == exp["fit"], case["id"]
        if ok:
            assert d == exp["driver"], case["id"]
            assert ex == exp["executors"], case["id"]
        # SparkBinPackFunction must not mutate the metadata (binpack.go:60-87)
        assert np.array_equal(cl.available()[0], cpu) and np.array_equal(cl.available()[1], mem)


@pytest.mark.parametrize("algo", ["tightly-pack", "distribute-evenly"])
def test_closed_oracle_pack_cases(golden, oracle, algo):
    for case in golden["pack_cases"]:
        names, cpu, mem, gpu = case_arrays(case["nodes"])
        app = case["app"]
        drv = res_aos([app["drv"][0]], [app["drv"][1]], [app["drv"][2]])
        exe = res_aos([app["exe"][0]], [app["exe"][1]], [app["exe"][2]])
        _, dn, en, off, _ = oracle.closed_batch(
            ALGO_ID[algo], 0, cpu, mem, gpu, order_indices(case["driver_order"], names),
            order_indices(case["exec_order"], names), drv, exe, [app["count"]])
        exp = case["expect"][algo]
        assert (dn[0] >= 0) == exp["fit"], case["id"]
        if exp["fit"]:
            assert names[dn[0]] == exp["driver"], case["id"]
            assert [names[i] for i in en[:app["count"]]] == exp["executors"], case["id"]


def _fifo_inputs(case):
    names, cpu, mem, gpu = case_arrays(case["nodes"])
    apps = case["apps"]
    drv = res_aos([a["drv"][0] for a in apps], [a["drv"][1] for a in apps], [a["drv"][2] for a in apps])
    exe = res_aos([a["exe"][0] for a in apps], [a["exe"][1] for a in apps], [a["exe"][2] for a in apps])
    count = np.array([a["count"] for a in apps], np.int32)
    young = np.array([1 if a.get("young") else 0 for a in apps], np.uint8)
    return names, cpu, mem, gpu, drv, exe, count, young


def _check_fifo(case, names, blocked, dn, en, off, final):
    exp = case["expect"]
    assert blocked == exp["blocked"], case["id"]
    for i, r in enumerate(exp["results"]):
        if r["driver"] == "unevaluated":
            assert dn[i] == -2, case["id"]
        elif r["driver"] is None:
            assert dn[i] == -1, case["id"]
        else:
            assert names[dn[i]] == r["driver"], case["id"]
            assert [names[j] for j in en[off[i]:off[i + 1]]] == r["executors"], case["id"]
    for j, f in enumerate(exp["final_available"]):
        assert (final[0][j], final[1][j], final[2][j]) == (f["cpu"], f["mem"], f["gpu"]), case["id"]


def test_literal_oracle_fifo_cases(golden, oracle):
    for case in golden["fifo_cases"]:
        names, cpu, mem, gpu, drv, exe, count, young = _fifo_inputs(case)
        cl = oracle.Cluster(names, cpu, mem, gpu)
        blocked, dn, en, off = cl.fifo(ALGO_ID[case["algo"]], MODE_ID[case["mode"]], drv, exe, count, young, names, names)
        _check_fifo(case, names, blocked, dn, en, off, cl.available())


def test_closed_oracle_fifo_cases(golden, oracle):
    for case in golden["fifo_cases"]:
        names, cpu, mem, gpu, drv, exe, count, young = _fifo_inputs(case)
        idx = np.arange(len(names), dtype=np.int32)
        blocked, dn, en, off, final = oracle.closed_batch(ALGO_ID[case["algo"]], MODE_ID[case["mode"]], cpu, mem, gpu,
                                                          idx, idx, drv, exe, count, young)
        _check_fifo(case, names, blocked, dn, en, off, final)


def test_literal_oracle_zone_fifo_cases(golden, oracle):
    """fitEarlierDrivers with the zone-aware packers (the checker of gp_pack_fifo_zones): the literal C loop reproduces the
    committed fixture (generated by the pure-Python statement, tests/gen_golden.py)."""
    packer_id = {"single-az-tightly-pack": 2, "az-aware-tightly-pack": 3, "single-az-minimal-fragmentation": 5}
    assert len(golden["zone_fifo_cases"]) == 6
    for case in golden["zone_fifo_cases"]:
        names, cpu, mem, gpu, drv, exe, count, young = _fifo_inputs(case)
        sched = {s["name"]: (s["cpu"], s["mem"], s["gpu"]) for s in case["schedulable"]}
        sc = tuple(np.array([sched[n][k] for n in names], np.int64) for k in range(3))
        cl = oracle.Cluster(names, cpu, mem, gpu, sched=sc, zone=[case["zones"].get(n, "default") for n in names])
        blocked, dn, en, off = cl.fifo(packer_id[case["packer"]], MODE_ID[case["mode"]], drv, exe, count, young, names, names,
                                       with_efficiencies=True)
        _check_fifo(case, names, blocked, dn, en, off, cl.available())


def test_node_priority_order_goldens(golden, oracle):
    """internal/sort/nodesorting_test.go:98-182 through PotentialNodes (all nodes are candidates)."""
    for case in golden["sort_cases"]:
        names, cpu, mem, gpu = case_arrays(case["nodes"])
        zone = [case["zones"].get(n, "default") for n in names]
        cl = oracle.Cluster(names, cpu, mem, gpu, zone=zone)
        d, e = cl.potential_nodes(case["candidates"])
        assert d == case["expect_priority_order"], case["id"]
        assert e == case["expect_priority_order"], case["id"]
        assert d == case["expect_driver"] and e == case["expect_executor"]


def test_label_priority_goldens(golden, oracle):
    """internal/sort/nodesorting_test.go:195-252: stable re-sort by configured label rank.  The input
    order is reproduced by giving the nodes memory in that order (priority = memory ascending)."""
    for case in golden["label_cases"]:
        names = case["input"]
        n = len(names)
        cl = oracle.Cluster(names, np.ones(n, np.int64), np.arange(1, n + 1, dtype=np.int64), np.zeros(n, np.int64))
        rank = [case["rank"].get(nm, -1) for nm in names]
        d, e = cl.potential_nodes(names, driver_label_rank=rank, exec_label_rank=rank)
        assert d == case["expect"], case["id"]
        assert e == case["expect"], case["id"]


def test_potential_nodes_filters(oracle):
    """PotentialNodes: drivers = priority order ∩ candidates; executors = schedulable ∧ ready
    (internal/sort/nodesorting.go:51-58)."""
    names = ["a", "b", "c", "d"]
    cl = oracle.Cluster(names, [4, 3, 2, 1], [4, 3, 2, 1], None, unschedulable=[0, 1, 0, 0], ready=[1, 1, 0, 1])
    d, e = cl.potential_nodes(["a", "c", "zzz"])
    assert d == ["c", "a"]
    assert e == ["d", "a"]


def test_packing_efficiency_by_product(oracle):
    """LIB/binpack/efficiency.go:66-156 on V1 (schedulable == initial available): after tightly-pack
    n0 carries driver + 3 executors = 7 of 8 cores, 13 of 16 GiB."""
    Gi = 1 << 30
    names = ["n0", "n1", "n2", "n3"]
    cpu = np.full(4, 8000, np.int64); mem = np.full(4, 16 * Gi, np.int64); gpu = np.zeros(4, np.int64)
    cl = oracle.Cluster(names, cpu, mem, gpu, sched=(cpu, mem, gpu))
    ok, d, ex, eff = cl.binpack(0, (1000, Gi, 0), (2000, 4 * Gi, 0), 8, names, names, with_efficiencies=True)
    assert ok and d == "n0"
    # per node cpu: 7/8, 8/8, 2/8, 0 ; mem: 13/16, 16/16, 4/16, 0
    assert eff[0] == pytest.approx((7 / 8 + 1 + 2 / 8 + 0) / 4, rel=0, abs=1e-15)
    assert eff[1] == pytest.approx((13 / 16 + 1 + 4 / 16 + 0) / 4, rel=0, abs=1e-15)
    assert eff[2] == 1.0          # no node has GPUs -> 1 (efficiency.go:141-145)
    assert eff[3] == pytest.approx((7 / 8 + 1 + 2 / 8 + 0) / 4, rel=0, abs=1e-15)


ZONE_ALGO = {"single-az-tightly-pack": 2, "az-aware-tightly-pack": 3, "single-az-minimal-fragmentation": 5}


def test_minimal_fragmentation_goldens(golden, oracle):
    """the worked examples in the doc comment of minimalFragmentation (LIB/binpack/minimal_fragmentation.go:43-58) --
    the one place the reference states exact ExecutorNodes for this packer -- on the literal and the closed-form oracle"""
    from helpers import order_indices, res_aos
    assert sum(c["pinned"]["executors"] == "reference-doc-comment" for c in golden["minfrag_cases"]) == 4
    for case in golden["minfrag_cases"]:
        names, cpu, mem, gpu = case_arrays(case["nodes"])
        app, exp = case["app"], case["expect"]
        cl = oracle.Cluster(names, cpu, mem, gpu)
        ok, d, ex, _ = cl.binpack(4, app["drv"], app["exe"], app["count"], case["driver_order"], case["exec_order"])
        assert ok == exp["fit"], case["id"]
        _, cd, ce, _, _ = oracle.closed_batch(4, 0, cpu, mem, gpu, order_indices(case["driver_order"], names),
                                              order_indices(case["exec_order"], names), res_aos(*[[v] for v in app["drv"]]),
                                              res_aos(*[[v] for v in app["exe"]]), [app["count"]])
        assert (cd[0] >= 0) == exp["fit"], case["id"]
        if ok:
            assert d == exp["driver"] and ex == exp["executors"], case["id"]
            assert names[cd[0]] == exp["driver"] and [names[i] for i in ce[:app["count"]]] == exp["executors"], case["id"]


def test_literal_oracle_zone_cases(golden, oracle):
    """single-az / az-aware tightly-pack: the harness shapes the reference's own tests run
    through `single-az-tightly-pack` (fit / no-fit pinned) and derived multi-zone cases incl. the
    chooseBestResult efficiency comparison and its Max > 0 quirk."""
    for case in golden["zone_cases"]:
        names, cpu, mem, gpu = case_arrays(case["nodes"])
        _, sc, sm, sg = case_arrays(case["schedulable"])
        zone = [case["zones"].get(n, "default") for n in names]
        cl = oracle.Cluster(names, cpu, mem, gpu, sched=(sc, sm, sg), zone=zone)
        app = case["app"]
        for algo, aid in ZONE_ALGO.items():
            ok, d, ex, _ = cl.binpack(aid, app["drv"], app["exe"], app["count"], names, names, with_efficiencies=True)
            exp = case["expect"][algo]
            assert ok == exp["fit"], (case["id"], algo)
            if ok:
                assert d == exp["driver"] and ex == exp["executors"], (case["id"], algo)


def test_zone_packers_random_three_way(oracle):
    """literal C == pure Python on random multi-zone clusters (efficiency-driven choice included)."""
    from oracle import pyref
    rng = np.random.default_rng(99)
    for trial in range(40):
        n = int(rng.integers(2, 24))
        names = ["n%02d" % i for i in range(n)]
        sched_cpu = rng.integers(1, 17, n) * 1000; sched_mem = rng.integers(1, 33, n) * (1 << 30); sched_gpu = rng.integers(0, 3, n)
        cpu = (sched_cpu * rng.uniform(0, 1, n)).astype(np.int64) // 250 * 250
        mem = (sched_mem * rng.uniform(0, 1, n)).astype(np.int64) // (1 << 28) * (1 << 28)
        gpu = np.minimum(sched_gpu, rng.integers(0, 3, n))
        zones = {nm: "z%d" % rng.integers(0, 3) for nm in names}
        order = list(rng.permutation(names))
        cl = oracle.Cluster(names, cpu, mem, gpu, sched=(sched_cpu, sched_mem, sched_gpu), zone=[zones[nm] for nm in names])
        meta = {names[i]: (int(cpu[i]), int(mem[i]), int(gpu[i])) for i in range(n)}
        sched = {names[i]: (int(sched_cpu[i]), int(sched_mem[i]), int(sched_gpu[i])) for i in range(n)}
        for _ in range(6):
            drv = (int(rng.integers(0, 3)) * 500, int(rng.integers(0, 3)) << 29, int(rng.integers(0, 2)))
            exe = (int(rng.integers(1, 5)) * 500, int(rng.integers(1, 9)) << 28, int(rng.integers(0, 2)))
            k = int(rng.integers(0, 10))
            for fn, aid in ((pyref.single_az_tightly_pack, 2), (pyref.az_aware_tightly_pack, 3), (pyref.single_az_minimal_fragmentation, 5)):
                d, ex, ok = fn(drv, exe, k, order, order, dict(meta), sched, zones)
                lok, ld, lex, _ = cl.binpack(aid, drv, exe, k, order, order, with_efficiencies=True)
                assert ok == lok, (trial, aid)
                if ok:
                    assert d == ld and ex == lex, (trial, aid)


def test_snapshot_build_oracles_agree(oracle):
    """f2: usage from hard + soft reservations and overhead -> available / schedulable (literal C == Python)."""
    from oracle import pyref
    rng = np.random.default_rng(7)
    n = 50
    names = ["n%02d" % i for i in range(n)]
    alloc = [rng.integers(1, 64, n) * 1000, rng.integers(1, 256, n) * (1 << 30), rng.integers(0, 9, n)]
    over = [rng.integers(0, 3, n) * 250, rng.integers(0, 4, n) * (1 << 28), np.zeros(n, np.int64)]
    R = 400
    rn = [names[i] if i < n else "gone-%d" % i for i in rng.integers(0, n + 5, R)]     # some reservations on nodes that left
    res = [rng.integers(0, 8, R) * 500, rng.integers(0, 16, R) * (1 << 29), rng.integers(0, 2, R)]
    av, sc = oracle.node_scheduling_metadata(names, alloc, over, rn, res)
    pa, ps = pyref.node_scheduling_metadata({names[i]: tuple(int(x[i]) for x in alloc) for i in range(n)},
                                            {names[i]: tuple(int(x[i]) for x in over) for i in range(n)},
                                            [(rn[r], tuple(int(x[r]) for x in res)) for r in range(R)])
    for i, nm in enumerate(names):
        assert (av[0][i], av[1][i], av[2][i]) == pa[nm]
        assert (sc[0][i], sc[1][i], sc[2][i]) == ps[nm]
    assert (av[0] < 0).any()      # over-committed nodes exist: availability may be negative


def test_reschedule_goldens(golden, oracle):
    """rescheduleExecutor's node choice: two vectors pinned by the reference's own tests
    (TestMinimalFragmentation / TestMinimalFragmentationEdgeCase) plus derived ones, on the literal C oracle"""
    assert sum(c["pinned"] == "reference-test" for c in golden["resched_cases"]) == 2
    for case in golden["resched_cases"]:
        names, cpu, mem, gpu = case_arrays(case["nodes"])
        cl = oracle.Cluster(names, cpu, mem, gpu)
        got = cl.reschedule_executor(case["min_frag"], case["exe"], case["exec_order"],
                                     {k: tuple(v) for k, v in case["overhead"].items()}, case["hosting"])
        assert got == case["expect"], case["id"]


def test_reschedule_random_two_way(oracle):
    """literal C == pure Python on random clusters: ties in capacity, hosting sets, overhead larger than what is left"""
    from oracle import pyref
    rng = np.random.default_rng(4242)
    for trial in range(300):
        n = int(rng.integers(1, 30))
        names = ["n%02d" % i for i in range(n)]
        cpu = rng.integers(-1, 9, n) * 1000; mem = rng.integers(0, 9, n) * (1 << 30); gpu = rng.integers(0, 3, n)
        cl = oracle.Cluster(names, cpu, mem, gpu)
        meta = {names[i]: (int(cpu[i]), int(mem[i]), int(gpu[i])) for i in range(n)}
        order = [names[i] for i in rng.permutation(n)[: int(rng.integers(1, n + 1))]] + ["ghost"]
        exe = (int(rng.integers(0, 4)) * 1000, int(rng.integers(0, 3)) << 30, int(rng.integers(0, 2)))
        over = {names[i]: (int(rng.integers(0, 3)) * 500, int(rng.integers(0, 2)) << 29, 0) for i in range(n) if rng.random() < 0.3}
        hosting = [names[i] for i in range(n) if rng.random() < 0.25]
        assert cl.reschedule_executor(False, exe, order) == pyref.reschedule_first_fit(exe, order, meta), trial
        assert cl.reschedule_executor(True, exe, order, over, hosting) == \
            pyref.reschedule_minimal_fragmentation(exe, order, meta, over, set(hosting)), trial


def test_input_side_goldens(golden):
    """the reference's sparkpods_test.go pins the annotation -> tuple parsing (3 cases) and the FIFO
    queue membership/order (4 cases); the Python restatement reproduces the committed vectors"""
    from oracle import pyref
    assert sum(c["pinned"] == "reference-test" for c in golden["annotation_cases"]) == 3
    assert sum(c["pinned"] == "reference-test" for c in golden["queue_cases"]) == 4
    for case in golden["annotation_cases"]:
        err, got = pyref.spark_resources(case["annotations"])
        assert err == case["error"] and got == case["expect"], case["id"]
    for case in golden["queue_cases"]:
        assert [p["uid"] for p in pyref.filter_to_earliest_and_sort(case["driver"], case["pods"])] == case["expect"], case["id"]

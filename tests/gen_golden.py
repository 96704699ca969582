#!/usr/bin/env python
"""Generates tests/golden/hotpath_vectors.json.

Go cannot run in this environment and the reference is not importable, so the golden vectors are:
  (a) what the reference's OWN tests pin for this path (fit / no-fit, min-executor semantics, node
      priority orders) -- each case cites the reference test file:line; and
  (b) hand-derived vectors whose full outputs are re-derived here by the
      independent pure-Python literal restatement oracle/pyref.py; generation FAILS if pyref
      disagrees with the hand-derived expectation embedded below.
Fields marked "pinned": "reference-test" are asserted by a reference test; "derived" are not
(parity unpinned for those outputs -- authority is the cited source lines).

Run:  python tests/gen_golden.py      (rewrites the JSON in place; committed with the fixture)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

Gi = 1 << 30
ALGOS = ("tightly-pack", "distribute-evenly")


def nodes(*specs):
    return [{"name": n, "cpu": c, "mem": m, "gpu": g} for (n, c, m, g) in specs]


def pack_case(cid, source, node_list, app, driver_order=None, exec_order=None, expect_hand=None, pinned_fit=None):
    names = [n["name"] for n in node_list]
    driver_order = driver_order if driver_order is not None else names
    exec_order = exec_order if exec_order is not None else names
    meta = {n["name"]: (n["cpu"], n["mem"], n["gpu"]) for n in node_list}
    expect = {}
    for algo in ALGOS:
        d, ex, ok = pyref.spark_bin_pack(tuple(app["drv"]), tuple(app["exe"]), app["count"],
                                         driver_order, exec_order, dict(meta), pyref.DISTRIBUTORS[algo])
        expect[algo] = {"fit": ok, "driver": d, "executors": ex}
        if expect_hand and algo in expect_hand:
            hd, hex_ = expect_hand[algo]
            assert ok and d == hd and ex == hex_, (cid, algo, (d, ex), (hd, hex_))
        if pinned_fit is not None:
            assert ok == pinned_fit, (cid, algo, ok, pinned_fit)
    return {"id": cid, "source": source, "nodes": node_list, "driver_order": driver_order,
            "exec_order": exec_order, "app": app, "expect": expect,
            "pinned": {"fit": "reference-test" if pinned_fit is not None else "derived",
                       "driver": "derived", "executors": "derived"}}


def main():
    app_readme = {"drv": [1000, 1 * Gi, 0], "exe": [2000, 4 * Gi, 0], "count": 8}  # README.md:37-41
    cases = []
    # ---- SURVEY App. A.5 hand-derived vectors -------------------------------------------------
    v1_nodes = nodes(*[(f"n{i}", 8000, 16 * Gi, 0) for i in range(4)])
    cases.append(pack_case(
        "V1", "BASELINE.json configs[0]; README.md:37-41 app on 4 free nodes (hand-derived)",
        v1_nodes, app_readme,
        expect_hand={"tightly-pack": ("n0", ["n0"] * 3 + ["n1"] * 4 + ["n2"]),
                     "distribute-evenly": ("n0", ["n0", "n1", "n2", "n3"] * 2)}))
    cases.append(pack_case(
        "V3", "hand-derived: driver skips n0 (cpu), executors skip n0/n1",
        nodes(("n0", 500, 16 * Gi, 0), ("n1", 8000, 2 * Gi, 0), ("n2", 8000, 16 * Gi, 0), ("n3", 4000, 8 * Gi, 0)),
        {**app_readme, "count": 5},
        expect_hand={"tightly-pack": ("n1", ["n2"] * 4 + ["n3"]),
                     "distribute-evenly": ("n1", ["n2", "n3", "n2", "n3", "n2"])}))
    cases.append(pack_case(
        "V4", "hand-derived: driver shares n0 with one executor",
        nodes(("n0", 3000, 5 * Gi, 0), ("n1", 4000, 8 * Gi, 0), ("n2", 2000, 4 * Gi, 0)),
        {**app_readme, "count": 4},
        expect_hand={"tightly-pack": ("n0", ["n0", "n1", "n1", "n2"]),
                     "distribute-evenly": ("n0", ["n0", "n1", "n2", "n1"])}))
    cases.append(pack_case(
        "V6", "hand-derived: first driver candidate rejected by the executor fit (binpack.go:74-83)",
        nodes(("n0", 2000, 4 * Gi, 0), ("n1", 3000, 5 * Gi, 0)),
        {**app_readme, "count": 2},
        expect_hand={"tightly-pack": ("n1", ["n0", "n1"]),
                     "distribute-evenly": ("n1", ["n0", "n1"])}))
    # ---- pinned by the reference's own tests ---------------------------------------------------
    harness_nodes = nodes(("node1", 8000, 8 * Gi, 1), ("node2", 8000, 8 * Gi, 1))  # extender_test_utils.go:239-271
    static_app = {"drv": [1000, 1, 1], "exe": [1000, 1, 0]}  # StaticAllocationSparkPods "1","1","1","1" + driver gpu 1
    cases.append(pack_case(
        "T0", "internal/extender/resource_test.go:27-51 TestScheduler: 2-executor app fits 2 nodes",
        harness_nodes, {**static_app, "count": 2}, pinned_fit=True))
    cases.append(pack_case(
        "T1a", "internal/extender/unschedulablepods_test.go:24-43: 2-executor app does not exceed capacity",
        harness_nodes, {**static_app, "count": 2}, pinned_fit=True))
    cases.append(pack_case(
        "T1b", "internal/extender/unschedulablepods_test.go:45-53: 100-executor app exceeds capacity",
        harness_nodes, {**static_app, "count": 100}, pinned_fit=False))
    cases.append(pack_case(
        "T2", "internal/extender/unschedulablepods_test.go:55-80 TestSchedulerFailsToScheduleWhenNotEnoughNvidiaGPUs",
        harness_nodes, {"drv": [1000, 1, 1], "exe": [1000, 1, 1], "count": 2}, pinned_fit=False))
    cases.append(pack_case(
        "T3", "internal/extender/resource_test.go:172-196 dynamic allocation min=1/max=3: only MIN is gang-packed "
              "(EXT/resource.go:242,325) -> exactly one executor slot",
        harness_nodes, {**static_app, "count": 1}, pinned_fit=True))
    # ---- edge cases the semantics imply (derived) ------------------------------------------------
    cases.append(pack_case(
        "E0", "count==0 succeeds on the first driver-feasible node (pack_tightly.go:42-44, distribute_evenly.go:46-48)",
        nodes(("n0", 500, 1 * Gi, 0), ("n1", 4000, 8 * Gi, 0)), {**app_readme, "count": 0}))
    cases.append(pack_case(
        "E1", "negative availability: over-committed node never hosts, even for zero-request dims (resources.go:239)",
        nodes(("n0", 8000, 16 * Gi, -1), ("n1", -500, 64 * Gi, 0), ("n2", 8000, 16 * Gi, 0)),
        {**app_readme, "count": 3}))
    cases.append(pack_case(
        "E2", "driver node not in executor order; executor order names a node missing from metadata",
        nodes(("n0", 1000, 1 * Gi, 0), ("n1", 8000, 16 * Gi, 0), ("n2", 8000, 16 * Gi, 0)),
        {**app_readme, "count": 4}, driver_order=["ghost", "n0", "n1"], exec_order=["ghost", "n1", "n2"]))
    cases.append(pack_case(
        "E3", "no driver fits anywhere", nodes(("n0", 500, 16 * Gi, 0), ("n1", 900, 16 * Gi, 0)),
        {**app_readme, "count": 1}))
    cases.append(pack_case(
        "E4", "zero-resource executors: unlimited capacity on the first admissible node",
        nodes(("n0", 1000, 1 * Gi, 0), ("n1", 8000, 16 * Gi, 0)),
        {"drv": [1000, 1 * Gi, 0], "exe": [0, 0, 0], "count": 5}))
    cases.append(pack_case(
        "E5", "distribute-evenly needs 3 rounds with uneven capacities",
        nodes(("n0", 7000, 64 * Gi, 0), ("n1", 2000, 64 * Gi, 0), ("n2", 4500, 64 * Gi, 0)),
        {**app_readme, "count": 6}))
    cases.append(pack_case(
        "E6", "empty executor order with count>0 never fits; empty driver order never fits",
        nodes(("n0", 8000, 16 * Gi, 0)), {**app_readme, "count": 1}, exec_order=[]))

    # ---- single-AZ / AZ-aware tightly-pack ----------------------------------------
    zone_cases = []

    def zone_case(cid, source, node_list, zones, app, pinned_fit=None, sched=None):
        names = [n["name"] for n in node_list]
        meta = {n["name"]: (n["cpu"], n["mem"], n["gpu"]) for n in node_list}
        sch = sched if sched is not None else dict(meta)
        expect = {}
        for algo, fn in (("single-az-tightly-pack", pyref.single_az_tightly_pack), ("az-aware-tightly-pack", pyref.az_aware_tightly_pack),
                         ("single-az-minimal-fragmentation", pyref.single_az_minimal_fragmentation)):
            d, ex, ok = fn(tuple(app["drv"]), tuple(app["exe"]), app["count"], names, names, dict(meta), sch, zones)
            expect[algo] = {"fit": ok, "driver": d, "executors": ex}
            if pinned_fit is not None:
                assert ok == pinned_fit, (cid, algo)
        return {"id": cid, "source": source, "nodes": node_list, "zones": zones,
                "schedulable": [{"name": n, "cpu": sch[n][0], "mem": sch[n][1], "gpu": sch[n][2]} for n in names],
                "app": app, "expect": expect,
                "pinned": {"fit": "reference-test" if pinned_fit is not None else "derived", "driver": "derived", "executors": "derived"}}

    hz = {"node1": "zone1", "node2": "zone1"}
    zone_cases.append(zone_case("Z-T0", "internal/extender/resource_test.go:27-51 TestScheduler through single-az-tightly-pack (the harness packer)",
                                harness_nodes, hz, {**static_app, "count": 2}, pinned_fit=True))
    zone_cases.append(zone_case("Z-T1b", "internal/extender/unschedulablepods_test.go:45-53: 100 executors exceed capacity",
                                harness_nodes, hz, {**static_app, "count": 100}, pinned_fit=False))
    zone_cases.append(zone_case("Z-T2", "internal/extender/unschedulablepods_test.go:55-80: not enough GPUs",
                                harness_nodes, hz, {"drv": [1000, 1, 1], "exe": [1000, 1, 1], "count": 2}, pinned_fit=False))
    three_zone = nodes(("a1", 8000, 16 * Gi, 0), ("a2", 2000, 4 * Gi, 0), ("b1", 16000, 64 * Gi, 0), ("b2", 4000, 8 * Gi, 0),
                       ("c1", 3000, 6 * Gi, 0))
    tz = {"a1": "za", "a2": "za", "b1": "zb", "b2": "zb", "c1": "zc"}
    tsched = {"a1": (8000, 16 * Gi, 0), "a2": (8000, 16 * Gi, 0), "b1": (16000, 64 * Gi, 0), "b2": (8000, 16 * Gi, 0), "c1": (4000, 8 * Gi, 0)}
    zone_cases.append(zone_case("Z1", "derived: both za and zb fit; chooseBestResult takes the higher average Max efficiency (single_az.go:75-97)",
                                three_zone, tz, {**app_readme, "count": 3}, sched=tsched))
    zone_cases.append(zone_case("Z2", "derived: no single zone fits 9 executors; az-aware falls back to plain tightly-pack (az_aware_pack_tightly.go:33-37)",
                                three_zone, tz, {**app_readme, "count": 9}, sched=tsched))
    zone_cases.append(zone_case("Z3", "derived: zero-resource app: every zone 'fits' with efficiency... chooseBestResult needs Max > 0",
                                nodes(("a1", 0, 0, 0), ("b1", 0, 0, 0)), {"a1": "za", "b1": "zb"},
                                {"drv": [0, 0, 0], "exe": [0, 0, 0], "count": 2}, sched={"a1": (0, 0, 0), "b1": (0, 0, 0)}))

    # TestMinimalFragmentation / TestMinimalFragmentationEdgeCase (internal/extender/resource_test.go:73-165) schedule their
    # drivers through single-az-minimal-fragmentation: a dynamic-allocation app (min 1 executor) on the two harness nodes fits
    zone_cases.append(zone_case("Z-MF", "internal/extender/resource_test.go:73-124 TestMinimalFragmentation: driver of DynamicAllocationSparkPods(app, 1, 2)",
                                harness_nodes, hz, {**static_app, "count": 1}, pinned_fit=True))

    # ---- minimal fragmentation: LIB/binpack/minimal_fragmentation.go ---------------------------
    # the doc comment of minimalFragmentation (:43-58): capacities a1 b1 c3 d5 e5 f17 for one executor shape
    mf_nodes = nodes(("a", 1000, 64 * Gi, 0), ("b", 1000, 64 * Gi, 0), ("c", 3000, 64 * Gi, 0), ("d", 5000, 64 * Gi, 0),
                     ("e", 5000, 64 * Gi, 0), ("f", 17000, 64 * Gi, 0), ("drv", 500, 64 * Gi, 0))
    mf_exec_order = ["a", "b", "c", "d", "e", "f"]
    minfrag_cases = []

    def minfrag_case(cid, source, count, doc_expect=None, pinned="reference-doc-comment"):
        app = {"drv": [500, 1, 0], "exe": [1000, 1, 0], "count": count}
        meta = {n["name"]: (n["cpu"], n["mem"], n["gpu"]) for n in mf_nodes}
        d, ex, ok = pyref.spark_bin_pack(tuple(app["drv"]), tuple(app["exe"]), count, ["drv"], mf_exec_order, dict(meta),
                                         pyref.minimal_fragmentation)
        if doc_expect is not None:
            assert ok and d == "drv" and ex == doc_expect, (cid, ex, doc_expect)
        return {"id": cid, "source": source, "nodes": mf_nodes, "driver_order": ["drv"], "exec_order": mf_exec_order, "app": app,
                "expect": {"fit": ok, "driver": d, "executors": ex}, "pinned": {"executors": pinned}}

    minfrag_cases.append(minfrag_case("MF-11", "minimal_fragmentation.go:45-46 executorCount = 11", 11, ["d"] * 5 + ["e"] * 5 + ["a"]))
    minfrag_cases.append(minfrag_case("MF-6", "minimal_fragmentation.go:48-49 executorCount = 6", 6, ["d"] * 5 + ["a"]))
    minfrag_cases.append(minfrag_case("MF-15", "minimal_fragmentation.go:51-52 executorCount = 15", 15,
                                      ["d"] * 5 + ["e"] * 5 + ["c"] * 3 + ["a", "b"]))
    minfrag_cases.append(minfrag_case("MF-17", "minimal_fragmentation.go:54-55 executorCount = 17", 17, ["f"] * 17))
    minfrag_cases.append(minfrag_case(
        "MF-19", "minimal_fragmentation.go:57-58 says [f x17, a, b] for executorCount = 19, but the code below it (:105-114) puts "
        "the remainder 2 on the first node whose capacity is >= 2, i.e. c -- the code is the authority", 19,
        ["f"] * 17 + ["c", "c"], pinned="derived-from-code"))
    minfrag_cases.append(minfrag_case("MF-32", "derived: every node is consumed", 32,
                                      ["f"] * 17 + ["d"] * 5 + ["e"] * 5 + ["c"] * 3 + ["a", "b"], pinned="derived"))
    minfrag_cases.append(minfrag_case("MF-33", "derived: one more than the cluster holds", 33, pinned="derived"))

    # ---- executor reschedule: the node choice pinned by the reference's own tests ----------------
    resched_cases = []

    def resched_case(cid, source, node_list, exe, order, min_frag, hosting=(), overhead=None, pinned=None):
        meta = {n["name"]: (n["cpu"], n["mem"], n["gpu"]) for n in node_list}
        if min_frag:
            got = pyref.reschedule_minimal_fragmentation(tuple(exe), order, meta, overhead or {}, set(hosting))
        else:
            got = pyref.reschedule_first_fit(tuple(exe), order, meta)
        if pinned is not None:
            assert got == pinned, (cid, got, pinned)
        return {"id": cid, "source": source, "nodes": node_list, "exe": list(exe), "exec_order": order, "min_frag": bool(min_frag),
                "hosting": list(hosting), "overhead": {k: list(v) for k, v in (overhead or {}).items()},
                "expect": got, "pinned": "reference-test" if pinned is not None else "derived"}

    # TestMinimalFragmentation (resource_test.go:73-124): static-app (driver + 2 executors, 1 cpu / 1 B each) sits on node1,
    # dyn-app's driver and exec-0 sit on node2; exec-1 is offered [node1, node2] and must go to node2 although node1 sorts first
    resched_cases.append(resched_case(
        "R-MF1", "internal/extender/resource_test.go:73-124 TestMinimalFragmentation: 'attracted to the node already hosting the first executor'",
        nodes(("node1", 5000, 8 * Gi - 3, 0), ("node2", 6000, 8 * Gi - 2, 0)), (1000, 1, 0), ["node1", "node2"], True,
        hosting=["node2"], pinned="node2"))
    # TestMinimalFragmentationEdgeCase (resource_test.go:126-165): node1 hosts a 1-cpu / 4-B driver, node2 a 4-cpu / 1-B driver;
    # the 3-cpu executor has capacity 2 on node1 and 1 on node2 -> node2 ("has the smallest capacity"), no node hosts the app yet
    resched_cases.append(resched_case(
        "R-MF2", "internal/extender/resource_test.go:126-165 TestMinimalFragmentationEdgeCase: 'scheduled on node2 as it has the smallest capacity'",
        nodes(("node1", 7000, 8 * Gi - 4, 0), ("node2", 4000, 8 * Gi - 1, 0)), (3000, 1, 0), ["node1", "node2"], True, pinned="node2"))
    resched_cases.append(resched_case(
        "R-MF3", "derived: the overhead map is taken off again inside GetNodeCapacities (resource.go:682) -> node1 drops to capacity 0",
        nodes(("node1", 3000, 8 * Gi, 0), ("node2", 9000, 8 * Gi, 0)), (3000, 1, 0), ["node1", "node2"], True,
        overhead={"node1": (500, 0, 0)}))
    resched_cases.append(resched_case(
        "R-FF1", "derived: first fit over the executor order (resource.go:657-662)",
        nodes(("node1", 500, 8 * Gi, 0), ("node2", 1000, 8 * Gi, 0), ("node3", 8000, 8 * Gi, 0)), (1000, 1, 0),
        ["node1", "node2", "node3"], False))
    resched_cases.append(resched_case(
        "R-FF2", "derived: nothing fits -> 'not enough capacity to reschedule the executor' (resource.go:672)",
        nodes(("node1", 500, 8 * Gi, 0)), (1000, 1, 0), ["node1"], False))

    # ---- inputs of the path: annotations -> tuple, queue membership/order -----------------------
    Mi = 1 << 20
    static_ann = {"spark-driver-cpu": "1", "spark-driver-mem": "2432Mi", "spark-driver-nvidia.com/gpu": "1", "spark-executor-cpu": "2",
                  "spark-executor-mem": "6758Mi", "spark-executor-nvidia.com/gpu": "1", "spark-executor-count": "2"}
    dyn_ann = {k: v for k, v in static_ann.items() if k != "spark-executor-count"}
    dyn_ann.update({"spark-dynamic-allocation-enabled": "true", "spark-dynamic-allocation-min-executor-count": "2",
                    "spark-dynamic-allocation-max-executor-count": "5"})
    nogpu_ann = {k: v for k, v in static_ann.items() if "gpu" not in k}
    annotation_cases = []

    def annotation_case(cid, source, ann, pinned=None):
        err, got = pyref.spark_resources(ann)
        if pinned is not None:
            assert err is None and got == pinned, (cid, err, got)
        return {"id": cid, "source": source, "annotations": ann, "error": err, "expect": got,
                "pinned": "reference-test" if pinned is not None else "derived"}

    annotation_cases.append(annotation_case("A-static", "internal/extender/sparkpods_test.go:46-66 static allocation", static_ann,
                                            {"drv": [1000, 2432 * Mi, 1], "exe": [2000, 6758 * Mi, 1], "min": 2, "max": 2, "exact": True}))
    annotation_cases.append(annotation_case("A-dynamic", "internal/extender/sparkpods_test.go:67-89 dynamic allocation", dyn_ann,
                                            {"drv": [1000, 2432 * Mi, 1], "exe": [2000, 6758 * Mi, 1], "min": 2, "max": 5, "exact": True}))
    annotation_cases.append(annotation_case("A-nogpu", "internal/extender/sparkpods_test.go:90-108 no gpu annotations", nogpu_ann,
                                            {"drv": [1000, 2432 * Mi, 0], "exe": [2000, 6758 * Mi, 0], "min": 2, "max": 2, "exact": True}))
    annotation_cases.append(annotation_case("A-err-count", "sparkpods.go:90-91", {k: v for k, v in nogpu_ann.items() if k != "spark-executor-count"}))
    annotation_cases.append(annotation_case("A-err-missing", "sparkpods.go:98", {k: v for k, v in nogpu_ann.items() if k != "spark-executor-mem"}))
    annotation_cases.append(annotation_case("A-err-parse", "sparkpods.go:101-104", {**nogpu_ann, "spark-executor-mem": "4 GiB"}))
    annotation_cases.append(annotation_case("A-submilli", "derived: a valid Quantity finer than a millicore is flagged, not rounded",
                                            {**nogpu_ann, "spark-executor-cpu": "0.0005"}))

    def pod(created, uid, **kw):
        return {"uid": uid, "created": created, "node": "", "scheduler": "", "group": "instance-group-foobar", "deleting": False, **kw}

    queue_cases = []

    def queue_case(cid, source, driver, pods, pinned=None):
        got = [p["uid"] for p in pyref.filter_to_earliest_and_sort(driver, pods)]
        if pinned is not None:
            assert got == pinned, (cid, got, pinned)
        return {"id": cid, "source": source, "driver": driver, "pods": pods, "expect": got,
                "pinned": "reference-test" if pinned is not None else "derived"}

    me = pod(100, "1")
    queue_cases.append(queue_case("Q1", "internal/extender/sparkpods_test.go:182-189 selects earliest unassigned", me,
                                  [pod(101, "3"), pod(150, "2"), pod(100, "1")], []))
    queue_cases.append(queue_case("Q2", "sparkpods_test.go:190-194 selects if earliest and not in cache", me, [pod(101, "2")], []))
    queue_cases.append(queue_case("Q3", "sparkpods_test.go:195-202 does not select when not earliest", me,
                                  [pod(101, "3"), pod(99, "2"), pod(100, "1")], ["2"]))
    queue_cases.append(queue_case("Q4", "sparkpods_test.go:203-209 does not select when not earliest and not in cache", me,
                                  [pod(99, "3"), pod(101, "2")], ["3"]))
    queue_cases.append(queue_case("Q5", "derived: the other filters of sparkpods.go:59-64 and the ordering of :70-72", me,
                                  [pod(90, "b"), pod(50, "s", node="n1"), pod(60, "o", group="another-group"), pod(70, "d", deleting=True),
                                   pod(80, "f", scheduler="default-scheduler"), pod(10, "a")]))

    # ---- FIFO loop (fitEarlierDrivers) ----------------------------------------------------------
    fifo_cases = []

    def fifo_case(cid, source, node_list, apps, algo, mode, expect_hand=None):
        names = [n["name"] for n in node_list]
        meta = {n["name"]: (n["cpu"], n["mem"], n["gpu"]) for n in node_list}
        py_apps = [{"drv": tuple(a["drv"]), "exe": tuple(a["exe"]), "count": a["count"], "young": a.get("young", False)}
                   for a in apps]
        blocked, res = pyref.fit_earlier_drivers(py_apps, names, names, meta, algo, mode)
        out = [{"driver": d, "executors": ex} for d, ex in res]
        final = [{"name": n, "cpu": meta[n][0], "mem": meta[n][1], "gpu": meta[n][2]} for n in names]
        if expect_hand:
            expect_hand(blocked, out, final)
        return {"id": cid, "source": source, "nodes": node_list, "apps": apps, "algo": algo, "mode": mode,
                "expect": {"blocked": blocked, "results": out, "final_available": final},
                "pinned": "derived"}

    def v5_ref(blocked, out, final):
        assert blocked == -1
        assert out[0] == {"driver": "n0", "executors": ["n0"] * 3 + ["n1"] * 4 + ["n2"]}
        assert out[1] == {"driver": "n0", "executors": ["n0"] * 2 + ["n1"] * 3 + ["n2"] * 3}
        assert out[2] == {"driver": "n0", "executors": ["n0"] + ["n1"] * 2 + ["n2"] * 2 + ["n3"] * 3}
        assert [(f["cpu"], f["mem"]) for f in final] == [(2000, 4 * Gi)] * 3 + [(6000, 12 * Gi)]

    def v5_exact(blocked, out, final):
        assert blocked == 1 and out[0]["driver"] == "n0" and out[1]["driver"] is None and out[2]["driver"] == "unevaluated"

    three = [dict(app_readme) for _ in range(3)]
    fifo_cases.append(fifo_case("V5-reference", "SURVEY App. A.5 V5: sparkResourceUsage overwrite quirk (EXT/sparkpods.go:139-146)",
                                v1_nodes, three, "tightly-pack", "reference", v5_ref))
    fifo_cases.append(fifo_case("V5-exact", "SURVEY App. A.5 V5 with exact accounting: second app blocks the queue",
                                v1_nodes, three, "tightly-pack", "exact", v5_exact))
    young = [dict(app_readme), {**app_readme, "count": 40, "young": True}, dict(app_readme),
             {**app_readme, "count": 40}, dict(app_readme)]
    fifo_cases.append(fifo_case("F1-young-skip", "EXT/resource.go:244-253: non-fitting young driver is skipped, old one blocks",
                                v1_nodes, young, "tightly-pack", "reference"))
    fifo_cases.append(fifo_case("F2-evenly", "FIFO with distribute-evenly, reference accounting",
                                v1_nodes, three, "distribute-evenly", "reference"))
    fifo_cases.append(fifo_case("F3-evenly-exact", "FIFO with distribute-evenly, exact accounting",
                                v1_nodes, three, "distribute-evenly", "exact"))

    # ---- FIFO loop with the zone-aware packers (fitEarlierDrivers calling single_az.go per queued driver) -------------
    zone_fifo_cases = []
    ZONE_FN = {"single-az-tightly-pack": pyref.single_az_tightly_pack, "az-aware-tightly-pack": pyref.az_aware_tightly_pack,
               "single-az-minimal-fragmentation": pyref.single_az_minimal_fragmentation}

    def zone_fifo_case(cid, source, node_list, zones, sched, apps, packer, mode):
        names = [n["name"] for n in node_list]
        meta = {n["name"]: (n["cpu"], n["mem"], n["gpu"]) for n in node_list}
        fn = ZONE_FN[packer]
        out, blocked = [], -1
        for i, a in enumerate(apps):
            if blocked >= 0:
                out.append({"driver": "unevaluated", "executors": []})
                continue
            d, ex, ok = fn(tuple(a["drv"]), tuple(a["exe"]), a["count"], names, names, meta, sched, zones)
            if not ok:
                out.append({"driver": None, "executors": []})
                if not a.get("young", False):
                    blocked = i
                continue
            out.append({"driver": d, "executors": ex})
            if mode == "reference":
                for n, u in pyref.spark_resource_usage(tuple(a["drv"]), tuple(a["exe"]), d, ex).items():
                    if n in meta:
                        meta[n] = pyref.sub(meta[n], u)
            else:
                meta[d] = pyref.sub(meta[d], tuple(a["drv"]))
                for n in ex:
                    meta[n] = pyref.sub(meta[n], tuple(a["exe"]))
        final = [{"name": n, "cpu": meta[n][0], "mem": meta[n][1], "gpu": meta[n][2]} for n in names]
        return {"id": cid, "source": source, "nodes": node_list, "zones": zones,
                "schedulable": [{"name": n, "cpu": sched[n][0], "mem": sched[n][1], "gpu": sched[n][2]} for n in names],
                "apps": apps, "packer": packer, "mode": mode,
                "expect": {"blocked": blocked, "results": out, "final_available": final}, "pinned": "derived"}

    zqueue = [{**app_readme, "count": 3}, {**app_readme, "count": 2}, {**app_readme, "count": 12, "young": True},
              {**app_readme, "count": 3}, {**app_readme, "count": 12}, {**app_readme, "count": 1}]
    for packer in ZONE_FN:
        for mode in ("reference", "exact"):
            zone_fifo_cases.append(zone_fifo_case(
                "ZF-%s-%s" % (packer, mode),
                "derived: EXT/resource.go:224-262 with %s as BinpackFunc (LIB/binpack/single_az.go:23-97): the zone chosen for one "
                "driver changes what the next one sees; a young driver that fits nowhere is skipped, an old one blocks the queue" % packer,
                three_zone, tz, tsched, zqueue, packer, mode))

    # ---- node priority order (internal/sort) ------------------------------------------------------
    sort_cases = []

    def sort_case(cid, source, node_list, zones, expect, candidates=None, **kw):
        meta = {n["name"]: (n["cpu"], n["mem"], n["gpu"]) for n in node_list}
        names = [n["name"] for n in node_list]
        got = pyref.node_names_in_priority_order(meta, zones)
        assert got == expect, (cid, got, expect)
        d, e = pyref.potential_nodes(meta, zones, candidates if candidates is not None else names, **kw)
        return {"id": cid, "source": source, "nodes": node_list, "zones": zones,
                "candidates": candidates if candidates is not None else names,
                "expect_priority_order": expect, "expect_driver": d, "expect_executor": e,
                "pinned": {"priority_order": "reference-test"}}

    sort_cases.append(sort_case(
        "S1", "internal/sort/nodesorting_test.go:98-141 TestAZAwareNodeSorting",
        nodes(("zone1Node1", 1, 1, 0), ("zone1Node2", 1, 2, 0), ("zone1Node3", 2, 1, 0), ("zone2Node1", 1, 1, 0)),
        {"zone1Node1": "zone1", "zone1Node2": "zone1", "zone1Node3": "zone1", "zone2Node1": "zone2"},
        ["zone2Node1", "zone1Node1", "zone1Node3", "zone1Node2"]))
    sort_cases.append(sort_case(
        "S2", "internal/sort/nodesorting_test.go:143-182 TestAZAwareNodeSortingWorksIfZoneLabelIsMissing",
        nodes(("node1", 2, 1, 0), ("node2", 2, 2, 0), ("node3", 1, 1, 0)), {},
        ["node3", "node1", "node2"]))
    # label priority (nodesorting_test.go:195-252): input order + label ranks -> expected order
    label_cases = [
        {"id": "L1", "source": "internal/sort/nodesorting_test.go:203-216 sorts when extra label values",
         "input": ["node1", "node3", "node2"], "rank": {"node2": 1, "node3": 0}, "expect": ["node3", "node2", "node1"]},
        {"id": "L2", "source": "internal/sort/nodesorting_test.go:217-230 extra nodes with no labels set",
         "input": ["node2", "node3", "node1"], "rank": {"node2": 1, "node3": 0}, "expect": ["node3", "node2", "node1"]},
        {"id": "L3", "source": "internal/sort/nodesorting_test.go:231-244 all nodes have ranked values",
         "input": ["node1", "node2", "node3"], "rank": {"node1": 1, "node2": 2, "node3": 0}, "expect": ["node3", "node1", "node2"]},
    ]

    out = {"_comment": "generated by tests/gen_golden.py -- do not edit by hand",
           "units": {"cpu": "millicores", "mem": "bytes", "gpu": "units"},
           "pack_cases": cases, "zone_cases": zone_cases, "minfrag_cases": minfrag_cases, "resched_cases": resched_cases, "annotation_cases": annotation_cases, "queue_cases": queue_cases, "fifo_cases": fifo_cases, "zone_fifo_cases": zone_fifo_cases, "sort_cases": sort_cases, "label_cases": label_cases}
    path = os.path.join(ROOT, "tests", "golden", "hotpath_vectors.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=False)
        f.write("\n")
    print("wrote", path, len(cases), "pack,", len(fifo_cases), "fifo,", len(sort_cases), "sort cases")


if __name__ == "__main__":
    main()

"""The reference's own end-to-end scenarios, driven through the device entry points for every step that is in scope
(SURVEY 8a/8f): reservations -> availability (gp_build_availability) -> node priority order (gp_potential_nodes) ->
single-az-tightly-pack for the driver + min executors (gp_pack_batch_zones) -> reservation table (gp_reserve_placements)
-> node choice for an executor without a reservation (gp_reschedule_executors).  Pod/CRD bookkeeping (which pod binds to
which reservation slot) is the reference's control plane and is modelled by a few lines of Python here.

Pinned by /root/reference/internal/extender/resource_test.go:
  TestDynamicAllocationScheduling :172-372  -- the node of every soft reservation (`expectedPodToNodeSoftReservationsMap`)
                                              and that only MIN executors are gang-reserved (`expectedReservations`)
  TestMinimalFragmentation        :73-120   -- the extra executor lands on the node already hosting the application
  TestMinimalFragmentationEdgeCase:122-165  -- ... on the node with the smallest capacity for THIS executor size
Harness shapes: extendertest.NewNode = 8 cpu / 8 GiB / 1 gpu (extender_test_utils.go:239-271); pods = 1 cpu / "1" byte of
memory, the DRIVER asks for 1 gpu (:281-337) -- so a node can host one driver only.  Every node's ZoneLabel is "default"
in the harness, i.e. the single-AZ packers see one zone."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Gi = 1 << 30


class MiniExtender:
    """State the reference keeps in its ResourceReservation / soft-reservation caches."""

    def __init__(self, packer, node_names, algo):
        self.p, self.names, self.algo = packer, list(node_names), algo
        n = len(self.names)
        self.alloc = [np.full(n, 8000, np.int64), np.full(n, 8 * Gi, np.int64), np.ones(n, np.int64)]
        self.res = []           # (node, cpu, mem, gpu, app, slot) hard reservations
        self.soft = []          # (node, cpu, mem, gpu, app, pod)
        self.pods = {}          # (app, executor pod) -> node it runs on
        self.bound = {}         # (app, reservation slot) -> pod
        self.order = sorted(range(n), key=lambda i: self.names[i])
        self.rank = np.empty(n, np.int32); self.rank[self.order] = np.arange(n)

    def _availability(self):
        rows = self.res + self.soft
        rn = np.array([r[0] for r in rows] or [0], np.int32)[: len(rows)]
        rs = [np.array([r[k] for r in rows], np.int64) for k in (1, 2, 3)]
        return self.p.build_availability(self.alloc, None, rn, rs)

    def schedule_driver(self, app, drv, exe, min_count, candidates):
        """selectDriverNode, internal/extender/resource.go:272-370 (FIFO queue empty in these scenarios)."""
        (ac, am, ag), (sc, sm, sg) = self._availability()
        cand = np.array([nm in candidates for nm in self.names], np.uint8)
        d_order, e_order = self.p.potential_nodes(ac, am, name_rank=self.rank, is_driver_candidate=cand, avail_gpu=ag)
        assert self.p.undefined_ties == 0
        self.p.set_snapshot(ac, am, ag, e_order, d_order)
        self.p.set_schedulable(sc, sm, sg)
        a = {"drv_cpu": [drv[0]], "drv_mem": [drv[1]], "drv_gpu": [drv[2]], "exe_cpu": [exe[0]], "exe_mem": [exe[1]], "exe_gpu": [exe[2]], "count": [min_count]}
        zone, dn, en, off, _ = self.p.pack_batch_zones(a, self.algo)
        assert zone[0] == 0, "the harness has one zone and capacity for the gang"
        rows = self.p.reserve_placements(a, (dn, en, off), subtract=False)
        for r in range(len(rows["app"])):
            self.res.append((int(rows["node"][r]), int(rows["cpu"][r]), int(rows["mem"][r]), int(rows["gpu"][r]), app, int(rows["slot"][r])))
        return self.names[dn[0]], [self.names[i] for i in en]

    def schedule_extra_executor(self, app, pod, exe, candidates):
        """rescheduleExecutor, resource.go:594-673 (overhead is empty in the harness): first fit over the executor order, or
        rescheduleExecutorWithMinimalFragmentation (:675-705)."""
        (ac, am, ag), _ = self._availability()
        idx = [i for i, nm in enumerate(self.names) if nm in candidates]            # availableNodes = getNodes(nodeNames)
        sub_rank = np.argsort(np.argsort([self.names[i] for i in idx])).astype(np.int32)
        _, e_local = self.p.potential_nodes(ac[idx], am[idx], name_rank=sub_rank, avail_gpu=ag[idx])
        e_order = np.array([idx[j] for j in e_local], np.int32)
        self.p.set_snapshot(ac, am, ag, e_order, e_order)
        hosting = None
        if self.algo == 2:      # getNodesWithExecutorsBelongingToSameApp (:683): nodes of the application's executor PODS
            hosting = [sorted({n for (a_, _), n in self.pods.items() if a_ == app})]
        node = self.p.reschedule_executors(([exe[0]], [exe[1]], [exe[2]]), min_frag=self.algo == 2, hosting=hosting)[0]
        assert node >= 0
        self.soft.append((int(node), exe[0], exe[1], exe[2], app, pod))
        self.pods[(app, pod)] = int(node)
        return self.names[node]

    def bind_to_reservation(self, app, pod, candidates):
        """An executor with an unbound reservation slot goes to that slot's node when kube-scheduler offers it
        (selectExecutorNode, resource.go:383-470); otherwise the executor is rescheduled and its reservation moves
        (rescheduleExecutor + the reservation update, control plane)."""
        for t, r in enumerate(self.res):
            if r[4] == app and r[5] > 0 and (app, r[5]) not in self.bound:
                self.bound[(app, r[5])] = pod
                if self.names[r[0]] in candidates:
                    self.pods[(app, pod)] = r[0]
                    return self.names[r[0]]
                del self.res[t]                                    # the slot's resources follow the pod
                node = self.schedule_extra_executor(app, pod, (r[1], r[2], r[3]), candidates)
                self.soft.pop()
                self.res.append((self.names.index(node), r[1], r[2], r[3], app, r[5]))
                return node
        return None


@pytest.fixture()
def packer(gangpack):
    p = gangpack.GangPacker()
    yield p
    p.close()


POD = (1000, 1, 0)            # 1 cpu, "1" byte, no gpu
DRV = (1000, 1, 1)            # the driver also asks for the node's only gpu


def test_dynamic_allocation_scenarios(packer):
    nodes = ["node1", "node2"]
    # "creates a reservation when under min executor count" / "... soft reservation for an executor over min" (:181-208)
    x = MiniExtender(packer, nodes, 0)
    d, e = x.schedule_driver("app", DRV, POD, 1, nodes)
    assert (d, e) == ("node1", ["node1"]) and len([r for r in x.res if r[5] > 0]) == 1       # exactly MIN executor slots
    assert x.bind_to_reservation("app", "exec-0", nodes) == "node1"
    assert x.schedule_extra_executor("app", "exec-1", POD, nodes) == "node1"                 # expectedPodToNodeSoftReservationsMap
    # "does not create any reservation for an executor over the max" (:225-241): the second extra executor, node1 again
    assert x.schedule_extra_executor("app", "exec-2", POD, nodes) == "node1"
    # "soft reservations are created on full nodes first" (:209-224): driver restricted to node2
    x = MiniExtender(packer, nodes, 0)
    d, e = x.schedule_driver("app", DRV, POD, 1, ["node2"])
    # the driver may only go to node2; the executor order still holds both (equally free) nodes, node1 first by name, so
    # tightly-pack reserves the executor slot on node1 ...
    assert (d, e) == ("node2", ["node1"])
    # ... but kube-scheduler offers exec-0 only [node2] (`nodeNames[1:]`): its reservation moves with it
    assert x.bind_to_reservation("app", "exec-0", ["node2"]) == "node2"
    # now node2 is the fuller node: it sorts first and takes the extra executor (expectedPodToNodeSoftReservationsMap: node2)
    assert x.schedule_extra_executor("app", "exec-1", POD, nodes) == "node2"
    # "schedules an executor only in the same AZ as the original application" (:263-292): static app on node1, the dynamic
    # app (min 0) on node2; its executors are restricted to the application's zone = {node2} by the control plane
    x = MiniExtender(packer, nodes, 0)
    assert x.schedule_driver("static", DRV, POD, 1, ["node1"]) == ("node1", ["node1"])
    d, e = x.schedule_driver("dyn", DRV, POD, 0, ["node2"])
    assert (d, e) == ("node2", [])
    assert x.schedule_extra_executor("dyn", "exec-0", POD, ["node2"]) == "node2"
    assert x.schedule_extra_executor("dyn", "exec-1", POD, ["node2"]) == "node2"


def test_minimal_fragmentation_flows(packer):
    nodes = ["node1", "node2"]
    # TestMinimalFragmentation (:73-120): static app (2 executors) pinned to node1, dynamic app: driver anywhere -> node2 is the
    # only node with a free gpu; exec-1 bound to its reservation; the extra executor is attracted to node2
    x = MiniExtender(packer, nodes, 2)
    d, e = x.schedule_driver("static", DRV, POD, 2, ["node1"])
    assert d == "node1" and e == ["node1", "node1"]
    d, e = x.schedule_driver("dyn", DRV, POD, 1, nodes)
    assert d == "node2"
    # minimalFragmentation reserves the one executor on the node with the smallest sufficient capacity: node1 (5 free cpu
    # against node2's 7 once the driver sits there) -- minimal_fragmentation.go:106-113
    assert e == ["node1"]
    # "we purposely schedule exec-1 on node2": kube-scheduler offers [node2] only, the reservation's node is not among them,
    # so the executor is rescheduled (rescheduleExecutorWithMinimalFragmentation over [node2]) and lands on node2
    assert x.bind_to_reservation("dyn", "exec-1", ["node2"]) == "node2"
    # the decision under test: [node1, node2] offered, node1 sorts first, node2 already hosts exec-1 -> node2
    assert x.schedule_extra_executor("dyn", "exec-2", POD, nodes) == "node2"
    # TestMinimalFragmentationEdgeCase (:122-165): static driver 4 GiB-"4" mem / 1 cpu on node1, dynamic driver 1 mem / 4 cpu on node2
    # (sizes are (mem, cpu) strings: "4","1" and "1","4"), executor of the dynamic app: mem "1", cpu "3" -> node2 has the smaller
    # capacity for that executor although node1 sorts first
    x = MiniExtender(packer, nodes, 2)
    assert x.schedule_driver("static", (1000, 4, 1), (1000, 1, 0), 0, ["node1"])[0] == "node1"
    assert x.schedule_driver("dyn", (4000, 1, 1), (3000, 1, 0), 0, ["node2"])[0] == "node2"
    assert x.schedule_extra_executor("dyn", "exec-0", (3000, 1, 0), nodes) == "node2"

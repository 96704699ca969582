// Host-layer parity tests, written like the reference's own tests for this path
// (internal/extender/resource_test.go, unschedulablepods_test.go, internal/sort/nodesorting_test.go)
// against the C++ mirror of the plug-in interface (k8s-spark-scheduler_b200/host/gangpack_host.hpp).
// Needs a B200: every BinpackFunc call goes through libgangpack.so to the device.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "gangpack_host.hpp"

static int failures = 0;
#define EXPECT(cond, msg)                                                            \
    do {                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); ++failures; } \
    } while (0)

using resources::CreateResources;
using resources::NodeGroupSchedulingMetadata;
using resources::NodeSchedulingMetadata;
typedef std::vector<std::string> Names;
static const int64_t Gi = 1ll << 30;

// extendertest.NewNode: 8 CPU / 8 GiB / 1 GPU (extender_test_utils.go:239-271)
static NodeSchedulingMetadata NewNode(const std::string& zone) {
    NodeSchedulingMetadata m;
    m.AvailableResources = CreateResources(8000, 8 * Gi, 1);
    m.SchedulableResources = m.AvailableResources;
    m.ZoneLabel = zone;
    return m;
}

static void TestScheduler() {   // resource_test.go:27-51
    NodeGroupSchedulingMetadata md;
    md["node1"] = NewNode("zone1"); md["node2"] = NewNode("zone1");
    Names nodeNames = {"node1", "node2"};
    const binpacker::Binpacker* bp = binpacker::SelectBinpacker("tightly-pack");
    // StaticAllocationSparkPods("2-executor-app", 2): driver 1 cpu / 1 B / 1 gpu, executors 1 cpu / 1 B
    auto r = bp->BinpackFunc(CreateResources(1000, 1, 1), CreateResources(1000, 1, 0), 2, nodeNames, nodeNames, md);
    EXPECT(r.HasCapacity, "There should be enough capacity to schedule the full application");
    EXPECT(r.DriverNode == "node1" && r.ExecutorNodes == Names({"node1", "node1"}), "tightly-pack fills node1 first");
    EXPECT(md["node1"].AvailableResources.CPU == 8000, "BinpackFunc must not mutate the metadata");
}

static void TestUnschedulablePodMarker() {   // unschedulablepods_test.go:24-53
    NodeGroupSchedulingMetadata md;
    md["node1"] = NewNode("zone1"); md["node2"] = NewNode("zone1");
    Names n = {"node1", "node2"};
    for (const char* name : {"tightly-pack", "distribute-evenly"}) {
        const binpacker::Binpacker* bp = binpacker::SelectBinpacker(name);
        auto two = bp->BinpackFunc(CreateResources(1000, 1, 1), CreateResources(1000, 1, 0), 2, n, n, md);
        EXPECT(two.HasCapacity, "The two executor application should fit to the cluster");
        auto hundred = bp->BinpackFunc(CreateResources(1000, 1, 1), CreateResources(1000, 1, 0), 100, n, n, md);
        EXPECT(!hundred.HasCapacity, "The hundred executor application should not fit to the cluster");
        EXPECT(hundred.DriverNode.empty() && hundred.ExecutorNodes.empty(), "EmptyPackingResult on no fit");
    }
}

static void TestSchedulerFailsToScheduleWhenNotEnoughNvidiaGPUs() {   // unschedulablepods_test.go:55-80
    NodeGroupSchedulingMetadata md;
    md["node1"] = NewNode("zone1"); md["node2"] = NewNode("zone1");
    Names n = {"node1", "node2"};
    auto r = binpack::TightlyPack(CreateResources(1000, 1, 1), CreateResources(1000, 1, 1), 2, n, n, md);
    EXPECT(!r.HasCapacity, "There should not be enough capacity to schedule the full application");
}

static void TestSelectBinpacker() {   // internal/binpacker/binpack.go:52-58
    EXPECT(binpacker::SelectBinpacker("tightly-pack")->Name == "tightly-pack", "tightly-pack");
    EXPECT(binpacker::SelectBinpacker("no-such-packer")->Name == "distribute-evenly", "unknown name selects distribute-evenly");
    EXPECT(!binpacker::SelectBinpacker("tightly-pack")->IsSingleAz, "IsSingleAz false");
    EXPECT(binpacker::SelectBinpacker("single-az-tightly-pack") == nullptr, "reference-only packers stay with the Go function");
}

static void TestExecutorNodeOrder() {   // SURVEY App. A.5 V1 / V3 (tests/golden/hotpath_vectors.json)
    NodeGroupSchedulingMetadata md;
    Names n;
    for (int i = 0; i < 4; ++i) {
        std::string name = "n" + std::to_string(i);
        md[name].AvailableResources = CreateResources(8000, 16 * Gi, 0);
        n.push_back(name);
    }
    auto drv = CreateResources(1000, Gi, 0), exe = CreateResources(2000, 4 * Gi, 0);   // README.md:37-41
    auto t = binpack::TightlyPack(drv, exe, 8, n, n, md);
    EXPECT(t.DriverNode == "n0" && t.ExecutorNodes == Names({"n0", "n0", "n0", "n1", "n1", "n1", "n1", "n2"}), "V1 tightly-pack");
    auto e = binpack::DistributeEvenly(drv, exe, 8, n, n, md);
    EXPECT(e.DriverNode == "n0" && e.ExecutorNodes == Names({"n0", "n1", "n2", "n3", "n0", "n1", "n2", "n3"}), "V1 distribute-evenly");
    // orders naming nodes missing from the metadata (binpack.go:68-69, pack_tightly.go:51-52)
    Names dn = {"ghost", "n0"}, en = {"ghost", "n1", "n2"};
    auto g = binpack::TightlyPack(drv, exe, 4, dn, en, md);
    EXPECT(g.HasCapacity && g.DriverNode == "n0" && g.ExecutorNodes == Names({"n1", "n1", "n1", "n1"}), "unknown names are skipped");
}

static void TestNodeSorting() {   // internal/sort/nodesorting_test.go:98-182
    NodeGroupSchedulingMetadata md;
    md["zone1Node1"].AvailableResources = CreateResources(1, 1, 0); md["zone1Node1"].ZoneLabel = "zone1";
    md["zone1Node2"].AvailableResources = CreateResources(1, 2, 0); md["zone1Node2"].ZoneLabel = "zone1";
    md["zone1Node3"].AvailableResources = CreateResources(2, 1, 0); md["zone1Node3"].ZoneLabel = "zone1";
    md["zone2Node1"].AvailableResources = CreateResources(1, 1, 0); md["zone2Node1"].ZoneLabel = "zone2";
    EXPECT(sort::NodeSorter::getNodeNamesInPriorityOrder(md) == Names({"zone2Node1", "zone1Node1", "zone1Node3", "zone1Node2"}),
           "TestAZAwareNodeSorting");
    NodeGroupSchedulingMetadata m2;
    m2["node1"].AvailableResources = CreateResources(2, 1, 0);
    m2["node2"].AvailableResources = CreateResources(2, 2, 0);
    m2["node3"].AvailableResources = CreateResources(1, 1, 0);
    EXPECT(sort::NodeSorter::getNodeNamesInPriorityOrder(m2) == Names({"node3", "node1", "node2"}),
           "TestAZAwareNodeSortingWorksIfZoneLabelIsMissing");
    // TestLabelPrioritySorting "sorts when extra label values" (:203-216), through PotentialNodes
    NodeGroupSchedulingMetadata m3;
    m3["node1"].AvailableResources = CreateResources(1, 1, 0); m3["node1"].AllLabels["test-label"] = "worst";
    m3["node3"].AvailableResources = CreateResources(1, 2, 0); m3["node3"].AllLabels["test-label"] = "best";
    m3["node2"].AvailableResources = CreateResources(1, 3, 0); m3["node2"].AllLabels["test-label"] = "good";
    sort::LabelPriorityOrder lp{"test-label", {"best", "good"}};
    sort::NodeSorter sorter(&lp, &lp);
    Names d, e;
    sorter.PotentialNodes(m3, {"node1", "node2", "node3"}, &d, &e);
    EXPECT(d == Names({"node3", "node2", "node1"}) && e == d, "label priority order");
    // executors must be schedulable and ready; drivers are filtered by the candidate list (:51-58)
    m3["node2"].Unschedulable = true;
    sort::NodeSorter plain;
    plain.PotentialNodes(m3, {"node2", "node3"}, &d, &e);
    EXPECT(d == Names({"node3", "node2"}) && e == Names({"node1", "node3"}), "PotentialNodes filters");
}

static void TestFitEarlierDrivers() {   // resource.go:224-262 with the sparkResourceUsage quirk (V5)
    NodeGroupSchedulingMetadata md;
    Names n;
    for (int i = 0; i < 4; ++i) {
        std::string name = "n" + std::to_string(i);
        md[name].AvailableResources = CreateResources(8000, 16 * Gi, 0);
        n.push_back(name);
    }
    extender::PendingDriver app;
    app.Resources.DriverResources = CreateResources(1000, Gi, 0);
    app.Resources.ExecutorResources = CreateResources(2000, 4 * Gi, 0);
    app.Resources.MinExecutorCount = 8; app.Resources.MaxExecutorCount = 8;
    std::vector<extender::PendingDriver> queue = {app, app, app};
    queue[1].Name = "broken"; queue.insert(queue.begin() + 1, extender::PendingDriver{});
    queue[1].ParseError = true;                       // skipped like a driver whose annotations do not parse
    std::vector<binpack::PackingResult> res;
    bool ok = extender::FitEarlierDrivers(*binpacker::SelectBinpacker("tightly-pack"), queue, n, n, md, &res);
    EXPECT(ok, "all earlier drivers fit");
    EXPECT(res[0].ExecutorNodes == Names({"n0", "n0", "n0", "n1", "n1", "n1", "n1", "n2"}), "app1");
    EXPECT(res[2].ExecutorNodes == Names({"n0", "n0", "n1", "n1", "n1", "n2", "n2", "n2"}), "app2 sees the (under-)charged snapshot");
    EXPECT(res[3].ExecutorNodes == Names({"n0", "n1", "n1", "n2", "n2", "n3", "n3", "n3"}), "app3");
    EXPECT(md["n0"].AvailableResources.CPU == 2000 && md["n0"].AvailableResources.Memory == 4 * Gi, "n0 after the loop");
    EXPECT(md["n3"].AvailableResources.CPU == 6000 && md["n3"].AvailableResources.Memory == 12 * Gi, "n3 after the loop");
    // an old driver that does not fit blocks the queue; a young one is skipped (resource.go:244-253)
    extender::PendingDriver big = app;
    big.Resources.MinExecutorCount = 40;
    std::vector<extender::PendingDriver> q2 = {big};
    NodeGroupSchedulingMetadata md2 = md;
    EXPECT(!extender::FitEarlierDrivers(*binpacker::SelectBinpacker("tightly-pack"), q2, n, n, md2), "old non-fitting driver blocks");
    q2[0].SkipIfNoFit = true;
    EXPECT(extender::FitEarlierDrivers(*binpacker::SelectBinpacker("tightly-pack"), q2, n, n, md2), "young non-fitting driver is skipped");
    EXPECT(md2["n0"].AvailableResources.CPU == md["n0"].AvailableResources.CPU, "skipped driver charges nothing");
    // sparkResourceUsage is an assignment (sparkpods.go:139-146)
    auto usage = extender::SparkResourceUsage(CreateResources(1, 1, 0), CreateResources(2, 2, 0), "a", {"a", "b", "b"});
    EXPECT(usage.size() == 2 && usage["a"].CPU == 2 && usage["b"].CPU == 2, "usage overwrite quirk");
}

static void TestFallbackPolicy() {
    NodeGroupSchedulingMetadata md;
    md["n0"].AvailableResources = CreateResources(8000, 16 * Gi, 0);
    Names n = {"n0"};
    bool threw = false;
    try {
        binpack::TightlyPack(CreateResources(1000, Gi, 0), CreateResources(1000, (int64_t)1 << 62, 0), 1, n, n, md);
    } catch (const gangpack::Error& e) { threw = (e.status == GP_ERR_UNREPRESENTABLE); }
    EXPECT(threw, "no silent CPU path: unrepresentable quantities raise without a fallback");
    gangpack::Fallbacks()[GP_TIGHTLY_PACK] = [](const resources::Resources&, const resources::Resources&, int, const Names&,
                                                const Names&, const NodeGroupSchedulingMetadata&) {
        binpack::PackingResult r; r.DriverNode = "from-fallback"; return r;
    };
    auto r = binpack::TightlyPack(CreateResources(1000, Gi, 0), CreateResources(1000, (int64_t)1 << 62, 0), 1, n, n, md);
    EXPECT(r.DriverNode == "from-fallback", "installed fallback (the original Go packer) serves unrepresentable inputs");
    gangpack::Fallbacks().clear();
}

int main() {
    TestScheduler();
    TestUnschedulablePodMarker();
    TestSchedulerFailsToScheduleWhenNotEnoughNvidiaGPUs();
    TestSelectBinpacker();
    TestExecutorNodeOrder();
    TestNodeSorting();
    TestFitEarlierDrivers();
    TestFallbackPolicy();
    if (failures) { std::printf("%d FAILED\n", failures); return 1; }
    std::printf("host_test: all passed\n");
    return 0;
}

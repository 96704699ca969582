// Host-layer parity tests, written like the reference's own tests for this path
// (internal/extender/resource_test.go, unschedulablepods_test.go, internal/sort/nodesorting_test.go)
// against the C++ mirror of the plug-in interface (k8s-spark-scheduler_b200/host/gangpack_host.hpp).
// Needs a B200: every BinpackFunc call goes through libgangpack.so to the device.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "gangpack_host.hpp"
#include "gangpack_oracle.h"   // the CPU oracle is the checker here (tests/ may link it; the product never does)
#include <random>

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
    EXPECT(binpacker::SelectBinpacker("single-az-tightly-pack")->IsSingleAz, "single-az-tightly-pack is a single-AZ packer (demands.go:169 reads this)");
    EXPECT(!binpacker::SelectBinpacker("az-aware-tightly-pack")->IsSingleAz, "az-aware-tightly-pack is not");
    EXPECT(binpacker::SelectBinpacker("single-az-minimal-fragmentation")->IsSingleAz, "single-az-minimal-fragmentation is a single-AZ packer (binpack.go:48)");
    EXPECT(binpacker::binpackFunctions().size() == 5, "all five binpack: values are served (binpack.go:43-49)");
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

// The reference's harness tests run through single-az-tightly-pack (resource_test.go:34, unschedulablepods_test.go:29,62)
static void TestHarnessThroughSingleAz() {
    NodeGroupSchedulingMetadata md;
    md["node1"] = NewNode("zone1"); md["node2"] = NewNode("zone1");
    Names n = {"node1", "node2"};
    for (const char* name : {"single-az-tightly-pack", "az-aware-tightly-pack", "single-az-minimal-fragmentation"}) {
        const binpacker::Binpacker* bp = binpacker::SelectBinpacker(name);
        EXPECT(bp->BinpackFunc(CreateResources(1000, 1, 1), CreateResources(1000, 1, 0), 2, n, n, md).HasCapacity,
               "TestScheduler: there should be enough capacity to schedule the full application");
        EXPECT(!bp->BinpackFunc(CreateResources(1000, 1, 1), CreateResources(1000, 1, 0), 100, n, n, md).HasCapacity,
               "TestUnschedulablePodMarker: the hundred executor application should not fit");
        EXPECT(!bp->BinpackFunc(CreateResources(1000, 1, 1), CreateResources(1000, 1, 1), 2, n, n, md).HasCapacity,
               "TestSchedulerFailsToScheduleWhenNotEnoughNvidiaGPUs");
    }
}

static NodeSchedulingMetadata Node(int64_t cpu, int64_t mem, int64_t gpu, int64_t scpu, int64_t smem, int64_t sgpu, const std::string& zone) {
    NodeSchedulingMetadata m;
    m.AvailableResources = CreateResources(cpu, mem, gpu);
    m.SchedulableResources = CreateResources(scpu, smem, sgpu);
    m.ZoneLabel = zone;
    return m;
}

static void TestZoneGoldens() {   // tests/golden/hotpath_vectors.json zone_cases Z1..Z3
    NodeGroupSchedulingMetadata md;
    md["a1"] = Node(8000, 16 * Gi, 0, 8000, 16 * Gi, 0, "za"); md["a2"] = Node(2000, 4 * Gi, 0, 8000, 16 * Gi, 0, "za");
    md["b1"] = Node(16000, 64 * Gi, 0, 16000, 64 * Gi, 0, "zb"); md["b2"] = Node(4000, 8 * Gi, 0, 8000, 16 * Gi, 0, "zb");
    md["c1"] = Node(3000, 6 * Gi, 0, 4000, 8 * Gi, 0, "zc");
    Names n = {"a1", "a2", "b1", "b2", "c1"};
    auto drv = CreateResources(1000, Gi, 0), exe = CreateResources(2000, 4 * Gi, 0);
    auto z1 = binpack::SingleAZTightlyPack(drv, exe, 3, n, n, md);
    EXPECT(z1.HasCapacity && z1.DriverNode == "a1" && z1.ExecutorNodes == Names({"a1", "a1", "a1"}), "Z1: chooseBestResult prefers the fuller zone");
    auto z2s = binpack::SingleAZTightlyPack(drv, exe, 9, n, n, md);
    EXPECT(z2s.HasCapacity && z2s.DriverNode == "b1" && z2s.ExecutorNodes.size() == 9 && z2s.ExecutorNodes[8] == "b2", "Z2 single-az");
    auto z2 = binpack::AzAwareTightlyPack(drv, exe, 9, n, n, md);
    EXPECT(z2.HasCapacity && z2.DriverNode == z2s.DriverNode && z2.ExecutorNodes == z2s.ExecutorNodes, "Z2 az-aware == single-az when a zone fits");
    NodeGroupSchedulingMetadata zero;
    zero["a1"] = Node(0, 0, 0, 0, 0, 0, "za"); zero["b1"] = Node(0, 0, 0, 0, 0, 0, "zb");
    Names zn = {"a1", "b1"};
    auto z3 = binpack::SingleAZTightlyPack(CreateResources(0, 0, 0), CreateResources(0, 0, 0), 2, zn, zn, zero);
    EXPECT(!z3.HasCapacity, "Z3: zero efficiency never beats WorstAvgPackingEfficiency (single_az.go:79-94)");
    auto z3a = binpack::AzAwareTightlyPack(CreateResources(0, 0, 0), CreateResources(0, 0, 0), 2, zn, zn, zero);
    EXPECT(z3a.HasCapacity && z3a.DriverNode == "a1" && z3a.ExecutorNodes == Names({"a1", "a1"}), "Z3: az-aware falls back to tightly-pack");
}

// random multi-zone clusters: the device-backed host layer vs the literal CPU oracle, all five packers + plain min-frag
static void TestRandomAgainstOracle() {
    std::mt19937_64 rng(20260922);
    auto U = [&](int64_t lo, int64_t hi) { return (int64_t)(lo + (int64_t)(rng() % (uint64_t)(hi - lo + 1))); };
    int compared = 0;
    for (int trial = 0; trial < 60; ++trial) {
        int n = (int)U(2, 40);
        NodeGroupSchedulingMetadata md;
        std::vector<std::string> names, zones;
        std::vector<int64_t> cpu, mem, gpu, scpu, smem, sgpu;
        for (int i = 0; i < n; ++i) {
            char buf[16]; std::snprintf(buf, sizeof buf, "n%02d", i);
            int64_t sc = U(1, 16) * 1000, sm = U(1, 32) * Gi, sg = U(0, 2);
            int64_t c = U(0, sc / 250) * 250, m = U(0, sm / (Gi / 4)) * (Gi / 4), g = std::min<int64_t>(sg, U(0, 2));
            std::string z = "z" + std::to_string(U(0, 2));
            md[buf] = Node(c, m, g, sc, sm, sg, z);
            names.push_back(buf); zones.push_back(z);
            cpu.push_back(c); mem.push_back(m); gpu.push_back(g); scpu.push_back(sc); smem.push_back(sm); sgpu.push_back(sg);
        }
        Names order = names;
        std::shuffle(order.begin(), order.end(), rng);
        std::vector<const char*> cn, cz, co;
        for (auto& s2 : names) cn.push_back(s2.c_str());
        for (auto& s2 : zones) cz.push_back(s2.c_str());
        for (auto& s2 : order) co.push_back(s2.c_str());
        orc_cluster* cl = orc_cluster_new(n, cn.data(), cpu.data(), mem.data(), gpu.data(), scpu.data(), smem.data(), sgpu.data(),
                                          cz.data(), nullptr, nullptr);
        for (int rep = 0; rep < 4; ++rep) {
            orc_res drv = {U(0, 2) * 500, U(0, 2) * (Gi / 2), U(0, 1)}, exe = {U(1, 4) * 500, U(1, 8) * (Gi / 4), U(0, 1)};
            int k = (int)U(0, rep == 3 ? 30 : 9);
            const char* algos[6] = {"tightly-pack", "distribute-evenly", "single-az-tightly-pack", "az-aware-tightly-pack",
                                    "minimal-fragmentation", "single-az-minimal-fragmentation"};   // index = oracle algo id
            for (int algo = 0; algo < 6; ++algo) {
                int32_t od = -1; std::vector<int32_t> oe((size_t)std::max(k, 1)); double eff[4];
                int ok = orc_binpack(cl, algo, &drv, &exe, k, co.data(), n, co.data(), n, 1, &od, oe.data(), eff);
                const binpack::SparkBinPackFunction& fn = algo == ORC_MINIMAL_FRAGMENTATION ? binpack::MinimalFragmentation
                                                                                            : binpacker::SelectBinpacker(algos[algo])->BinpackFunc;
                auto r = fn(CreateResources(drv.cpu, drv.mem, drv.gpu), CreateResources(exe.cpu, exe.mem, exe.gpu), k, order, order, md);
                bool same = (r.HasCapacity == (ok != 0));
                if (same && ok) {
                    same = r.DriverNode == names[(size_t)od] && (int)r.ExecutorNodes.size() == k;
                    for (int t = 0; same && t < k; ++t) same = r.ExecutorNodes[(size_t)t] == names[(size_t)oe[(size_t)t]];
                }
                if (!same) std::printf("mismatch trial %d rep %d algo %s\n", trial, rep, algos[algo]);
                EXPECT(same, "host layer (device) == literal oracle");
                ++compared;
            }
        }
        orc_cluster_free(cl);
    }
    std::printf("compared %d placements against the oracle\n", compared);
}

// fitEarlierDrivers with the zone-aware packers (resource.go:224-262 calling single_az.go per driver): host loop over
// device-packed zones vs orc_fifo on the string-keyed oracle, incl. the metadata the loop leaves behind
static void TestZoneAwareFifoAgainstOracle() {
    std::mt19937_64 rng(777);
    auto U = [&](int64_t lo, int64_t hi) { return (int64_t)(lo + (int64_t)(rng() % (uint64_t)(hi - lo + 1))); };
    const int ids[3] = {ORC_SINGLE_AZ_TIGHTLY_PACK, ORC_AZ_AWARE_TIGHTLY_PACK, ORC_SINGLE_AZ_MINIMAL_FRAGMENTATION};
    const char* names_of[3] = {"single-az-tightly-pack", "az-aware-tightly-pack", "single-az-minimal-fragmentation"};
    for (int trial = 0; trial < 12; ++trial) {
        const int n = (int)U(3, 24), q = 10;
        std::vector<std::string> names, zones;
        std::vector<int64_t> cpu, mem, gpu, scpu, smem, sgpu;
        for (int i = 0; i < n; ++i) {
            char buf[16]; std::snprintf(buf, sizeof buf, "n%02d", i);
            int64_t sc = U(4, 16) * 1000, sm = U(8, 32) * Gi;
            names.push_back(buf); zones.push_back("z" + std::to_string(U(0, 2)));
            cpu.push_back(U(0, sc / 500) * 500); mem.push_back(U(0, sm / Gi) * Gi); gpu.push_back(0);
            scpu.push_back(sc); smem.push_back(sm); sgpu.push_back(0);
        }
        Names order = names;
        std::shuffle(order.begin(), order.end(), rng);
        std::vector<const char*> cn, cz, co;
        for (auto& s2 : names) cn.push_back(s2.c_str());
        for (auto& s2 : zones) cz.push_back(s2.c_str());
        for (auto& s2 : order) co.push_back(s2.c_str());
        std::vector<orc_res> drv((size_t)q), exe((size_t)q);
        std::vector<int32_t> cnt((size_t)q);
        std::vector<uint8_t> young((size_t)q);
        std::vector<int64_t> off((size_t)q + 1, 0);
        std::vector<extender::PendingDriver> queue((size_t)q);
        for (int i = 0; i < q; ++i) {
            drv[(size_t)i] = {U(0, 2) * 500, U(0, 2) * Gi, 0}; exe[(size_t)i] = {U(1, 4) * 500, U(1, 4) * Gi, 0};
            cnt[(size_t)i] = (int32_t)U(0, 6); young[(size_t)i] = (uint8_t)(U(0, 3) != 0);
            off[(size_t)i + 1] = off[(size_t)i] + cnt[(size_t)i];
            auto& a = queue[(size_t)i];
            a.Resources.DriverResources = CreateResources(drv[(size_t)i].cpu, drv[(size_t)i].mem, 0);
            a.Resources.ExecutorResources = CreateResources(exe[(size_t)i].cpu, exe[(size_t)i].mem, 0);
            a.Resources.MinExecutorCount = cnt[(size_t)i]; a.SkipIfNoFit = young[(size_t)i] != 0;
        }
        for (int p = 0; p < 3; ++p) {
            NodeGroupSchedulingMetadata md;
            for (int i = 0; i < n; ++i) md[names[(size_t)i]] = Node(cpu[(size_t)i], mem[(size_t)i], 0, scpu[(size_t)i], smem[(size_t)i], 0, zones[(size_t)i]);
            orc_cluster* cl = orc_cluster_new(n, cn.data(), cpu.data(), mem.data(), gpu.data(), scpu.data(), smem.data(), sgpu.data(),
                                              cz.data(), nullptr, nullptr);
            std::vector<int32_t> od((size_t)q, -9), oe((size_t)std::max<int64_t>(off[(size_t)q], 1), -1);
            int32_t blocked = orc_fifo(cl, ids[p], ORC_FIFO_REFERENCE, q, drv.data(), exe.data(), cnt.data(), young.data(),
                                       co.data(), n, co.data(), n, 0, off.data(), od.data(), oe.data());
            std::vector<binpack::PackingResult> res;
            bool ok = extender::FitEarlierDrivers(*binpacker::SelectBinpacker(names_of[p]), queue, order, order, md, &res);
            bool same = ok == (blocked < 0);
            for (int i = 0; same && i < q; ++i) {
                if (od[(size_t)i] < 0) { same = !res[(size_t)i].HasCapacity; continue; }
                same = res[(size_t)i].HasCapacity && res[(size_t)i].DriverNode == names[(size_t)od[(size_t)i]];
                for (int t = 0; same && t < cnt[(size_t)i]; ++t)
                    same = res[(size_t)i].ExecutorNodes[(size_t)t] == names[(size_t)oe[(size_t)(off[(size_t)i] + t)]];
            }
            std::vector<int64_t> fc((size_t)n), fm((size_t)n), fg((size_t)n);
            orc_cluster_get_available(cl, fc.data(), fm.data(), fg.data());
            for (int i = 0; same && i < n; ++i)
                same = md[names[(size_t)i]].AvailableResources.CPU == fc[(size_t)i] && md[names[(size_t)i]].AvailableResources.Memory == fm[(size_t)i];
            if (!same) std::printf("zone-aware fifo mismatch trial %d packer %s\n", trial, names_of[p]);
            EXPECT(same, "FitEarlierDrivers (zone-aware packer) == orc_fifo");
            orc_cluster_free(cl);
        }
    }
}

// resource_test.go:73-165: the two reference tests that pin rescheduleExecutorWithMinimalFragmentation's node choice
static void TestMinimalFragmentationReschedule() {
    const binpacker::Binpacker& mf = *binpacker::SelectBinpacker("single-az-minimal-fragmentation");
    {   // TestMinimalFragmentation: static-app (driver + 2 executors) on node1, dyn-app driver + exec-0 on node2
        NodeGroupSchedulingMetadata md;
        md["node1"] = Node(5000, 8 * Gi - 3, 0, 8000, 8 * Gi, 1, "zone1");
        md["node2"] = Node(6000, 8 * Gi - 2, 0, 8000, 8 * Gi, 1, "zone1");
        auto r = extender::RescheduleExecutorNode(mf, CreateResources(1000, 1, 0), {"node1", "node2"}, md, {}, {}, {"node2"});
        EXPECT(r.second && r.first == "node2", "The dynamic pod should be attracted to the node already hosting the first executor");
    }
    {   // TestMinimalFragmentationEdgeCase
        NodeGroupSchedulingMetadata md;
        md["node1"] = Node(7000, 8 * Gi - 4, 0, 8000, 8 * Gi, 1, "zone1");
        md["node2"] = Node(4000, 8 * Gi - 1, 0, 8000, 8 * Gi, 1, "zone1");
        auto r = extender::RescheduleExecutorNode(mf, CreateResources(3000, 1, 0), {"node1", "node2"}, md, {}, {}, {});
        EXPECT(r.second && r.first == "node2", "This pod should be scheduled on node2 as it has the smallest capacity");
        resources::NodeGroupResources avail;
        avail["node1"] = CreateResources(7000, 8 * Gi - 4, 0); avail["node2"] = CreateResources(4000, 8 * Gi - 1, 0);
        auto ff = extender::RescheduleExecutorNode(*binpacker::SelectBinpacker("tightly-pack"), CreateResources(3000, 1, 0),
                                                   {"node1", "node2"}, md, avail, {}, {});
        EXPECT(ff.second && ff.first == "node1", "every other packer takes the first node of the order that fits (resource.go:657-662)");
        auto none = extender::RescheduleExecutorNode(mf, CreateResources(9000, 1, 0), {"node1", "node2"}, md, {}, {}, {});
        EXPECT(!none.second, "not enough capacity to reschedule the executor");
    }
}

int main() {
    TestScheduler();
    TestHarnessThroughSingleAz();
    TestZoneGoldens();
    TestRandomAgainstOracle();
    TestZoneAwareFifoAgainstOracle();
    TestMinimalFragmentationReschedule();
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

// gangpack_host.hpp -- C++ host-side mirror of the reference's plug-in interface for the placement
// hot path, implemented over the C ABI (include/gangpack.h).  The Go toolchain is absent in this
// environment, so this is the host layer a maintainer can compile and test; go/binpacker_gpu.go is
// the same thing as a cgo shim.  Names, argument meaning and error behaviour follow the reference:
//
//   resources::Resources / NodeSchedulingMetadata / NodeGroupSchedulingMetadata
//        vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg/resources/resources.go:103-246
//   binpack::PackingResult / SparkBinPackFunction / TightlyPack / DistributeEvenly
//        vendor/.../pkg/binpack/binpack.go:25-48, pack_tightly.go:25-32, distribute_evenly.go:25-32
//   binpacker::Binpacker / SelectBinpacker        internal/binpacker/binpack.go:37-58
//   sort::NodeSorter::PotentialNodes              internal/sort/nodesorting.go:41-200
//   extender::FitEarlierDrivers / SparkResourceUsage
//        internal/extender/resource.go:224-262, internal/extender/sparkpods.go:139-146
//
// Quantities are int64 (CPU millicores, memory bytes, GPU units) -- the exact-int64 model of
// resource.Quantity.  There is no CPU packer here: if the device path fails the call throws
// gangpack::Error unless a fallback packer has been installed (the Go shim installs the original
// Go function).
#pragma once

#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "gangpack.h"
#include "spark_resources.hpp"

namespace gangpack {
struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};
}  // namespace gangpack

// ------------------------------------------------------------------------------------------------
namespace resources {

struct Resources {                       // resources.go:151-155
    int64_t CPU = 0;        // millicores
    int64_t Memory = 0;     // bytes
    int64_t NvidiaGPU = 0;  // units
    bool GreaterThan(const Resources& o) const { return CPU > o.CPU || Memory > o.Memory || NvidiaGPU > o.NvidiaGPU; }  // :239-241
    bool Eq(const Resources& o) const { return CPU == o.CPU && Memory == o.Memory && NvidiaGPU == o.NvidiaGPU; }          // :244-246
    void Add(const Resources& o) { CPU += o.CPU; Memory += o.Memory; NvidiaGPU += o.NvidiaGPU; }                          // :202-206
    void Sub(const Resources& o) { CPU -= o.CPU; Memory -= o.Memory; NvidiaGPU -= o.NvidiaGPU; }                          // :209-213
};
inline Resources Zero() { return Resources{}; }                                                                            // :177-183
inline Resources CreateResources(int64_t cpuMilli, int64_t memory, int64_t gpus) { return Resources{cpuMilli, memory, gpus}; }

struct NodeSchedulingMetadata {          // resources.go:158-166
    Resources AvailableResources;
    Resources SchedulableResources;
    std::string ZoneLabel = "default";   // zoneLabelPlaceholder, :27
    std::map<std::string, std::string> AllLabels;
    bool Unschedulable = false;
    bool Ready = true;
};

using NodeGroupResources = std::unordered_map<std::string, Resources>;                 // :103
struct NodeGroupSchedulingMetadata : std::unordered_map<std::string, NodeSchedulingMetadata> {  // :106
    // SubtractUsageIfExists, :129-135
    void SubtractUsageIfExists(const NodeGroupResources& used) {
        for (const auto& kv : used) {
            auto it = find(kv.first);
            if (it != end()) it->second.AvailableResources.Sub(kv.second);
        }
    }
};

}  // namespace resources

// ------------------------------------------------------------------------------------------------
namespace binpack {

struct PackingResult {                   // binpack.go:25-30 (PackingEfficiencies: metrics by-product, not produced)
    std::string DriverNode;
    std::vector<std::string> ExecutorNodes;
    bool HasCapacity = false;
};
inline PackingResult EmptyPackingResult() { return PackingResult{}; }    // :33-40

using SparkBinPackFunction = std::function<PackingResult(               // :43-48
    const resources::Resources& driverResources, const resources::Resources& executorResources, int executorCount,
    const std::vector<std::string>& driverNodePriorityOrder, const std::vector<std::string>& executorNodePriorityOrder,
    const resources::NodeGroupSchedulingMetadata& nodesSchedulingMetadata)>;

}  // namespace binpack

// ------------------------------------------------------------------------------------------------
namespace gangpack {

// One device context + the marshalling between name-keyed Go-style maps and the SoA buffers.
class Device {
public:
    static Device& Get() {
        static Device d;
        return d;
    }
    gp_ctx* ctx() { return ctx_; }

    // Marshal metadata + the two orders (names absent from the metadata are dropped: they can host
    // neither a driver, binpack.go:68-69, nor an executor, pack_tightly.go:51-52) and upload.
    void SetSnapshot(const resources::NodeGroupSchedulingMetadata& md, const std::vector<std::string>& driverOrder,
                     const std::vector<std::string>& executorOrder) {
        names_.clear(); index_.clear();
        cpu_.clear(); mem_.clear(); gpu_.clear(); exec_.clear(); drv_.clear();
        auto intern = [&](const std::string& n) -> int32_t {
            auto it = index_.find(n);
            if (it != index_.end()) return it->second;
            auto m = md.find(n);
            if (m == md.end()) return -1;
            int32_t i = (int32_t)names_.size();
            names_.push_back(n);
            index_.emplace(n, i);
            cpu_.push_back(m->second.AvailableResources.CPU);
            mem_.push_back(m->second.AvailableResources.Memory);
            gpu_.push_back(m->second.AvailableResources.NvidiaGPU);
            return i;
        };
        std::vector<uint8_t> seen;
        auto add_unique = [&](std::vector<int32_t>& dst, const std::vector<std::string>& src) {
            seen.assign(src.size() + names_.size() + 1, 0);
            for (const auto& n : src) {
                int32_t i = intern(n);
                if (i < 0) continue;
                if ((size_t)i >= seen.size()) seen.resize((size_t)i + 1, 0);
                if (seen[i]) throw Error(GP_ERR_INVALID, "node listed twice in a priority order: " + n);
                seen[i] = 1;
                dst.push_back(i);
            }
        };
        add_unique(exec_, executorOrder);
        add_unique(drv_, driverOrder);
        int32_t eoff[2] = {0, (int32_t)exec_.size()}, doff[2] = {0, (int32_t)drv_.size()};
        gp_nodes n{};
        n.n_nodes = (int32_t)names_.size();
        n.avail_cpu_milli = cpu_.data(); n.avail_mem_bytes = mem_.data(); n.avail_gpu = gpu_.data();
        n.n_groups = 1;
        n.exec_off = eoff; n.exec_order = exec_.data();
        n.drv_off = doff; n.drv_order = drv_.data();
        check(gp_set_snapshot(ctx_, &n), "gp_set_snapshot");
    }
    // One instance group per entry of `groups` (driver order, executor order): used by the single-AZ packers,
    // where a "group" is a zone.  Nodes are interned once; a node may appear in one group only.
    void SetSnapshotGroups(const resources::NodeGroupSchedulingMetadata& md,
                           const std::vector<std::pair<std::vector<std::string>, std::vector<std::string>>>& groups) {
        names_.clear(); index_.clear();
        cpu_.clear(); mem_.clear(); gpu_.clear(); exec_.clear(); drv_.clear();
        std::vector<int32_t> eoff{0}, doff{0};
        auto intern = [&](const std::string& n) -> int32_t {
            auto it = index_.find(n);
            if (it != index_.end()) return it->second;
            auto m = md.find(n);
            if (m == md.end()) return -1;
            int32_t i = (int32_t)names_.size();
            names_.push_back(n);
            index_.emplace(n, i);
            cpu_.push_back(m->second.AvailableResources.CPU);
            mem_.push_back(m->second.AvailableResources.Memory);
            gpu_.push_back(m->second.AvailableResources.NvidiaGPU);
            return i;
        };
        for (const auto& g : groups) {
            for (const auto& n : g.second) { int32_t i = intern(n); if (i >= 0) exec_.push_back(i); }
            for (const auto& n : g.first) { int32_t i = intern(n); if (i >= 0) drv_.push_back(i); }
            eoff.push_back((int32_t)exec_.size());
            doff.push_back((int32_t)drv_.size());
        }
        gp_nodes n{};
        n.n_nodes = (int32_t)names_.size();
        n.avail_cpu_milli = cpu_.data(); n.avail_mem_bytes = mem_.data(); n.avail_gpu = gpu_.data();
        n.n_groups = (int32_t)groups.size();
        n.exec_off = eoff.data(); n.exec_order = exec_.data();
        n.drv_off = doff.data(); n.drv_order = drv_.data();
        check(gp_set_snapshot(ctx_, &n), "gp_set_snapshot");
    }
    // NodeSchedulingMetadata.SchedulableResources of the interned nodes (the packing efficiencies read them)
    void SetSchedulable(const resources::NodeGroupSchedulingMetadata& md) {
        std::vector<int64_t> c(names_.size()), m(names_.size()), g(names_.size());
        for (size_t i = 0; i < names_.size(); ++i) {
            const auto& s = md.at(names_[i]).SchedulableResources;
            c[i] = s.CPU; m[i] = s.Memory; g[i] = s.NvidiaGPU;
        }
        check(gp_set_schedulable(ctx_, c.data(), m.data(), g.data()), "gp_set_schedulable");
    }
    const std::string& name(int32_t i) const { return names_.at((size_t)i); }
    size_t n_nodes() const { return names_.size(); }
    int32_t index(const std::string& n) const { auto it = index_.find(n); return it == index_.end() ? -1 : it->second; }
    void check(int st, const char* what) {
        if (st != GP_OK) throw Error(st, std::string(what) + ": " + gp_last_error(ctx_));
    }

private:
    Device() {
        if (gp_create(&ctx_, nullptr) != GP_OK) throw Error(GP_ERR_NO_DEVICE, std::string("gp_create: ") + gp_last_error(nullptr));
    }
    ~Device() { gp_destroy(ctx_); }
    gp_ctx* ctx_ = nullptr;
    std::vector<std::string> names_;
    std::unordered_map<std::string, int32_t> index_;
    std::vector<int64_t> cpu_, mem_, gpu_;
    std::vector<int32_t> exec_, drv_;
};

// optional CPU fallback installed by the embedding application (the Go shim installs the original packer)
inline std::map<int, binpack::SparkBinPackFunction>& Fallbacks() {
    static std::map<int, binpack::SparkBinPackFunction> f;
    return f;
}

inline binpack::PackingResult PackOne(gp_algo algo, const resources::Resources& drv, const resources::Resources& exe, int count,
                                      const std::vector<std::string>& driverOrder, const std::vector<std::string>& executorOrder,
                                      const resources::NodeGroupSchedulingMetadata& md) {
    try {
        Device& d = Device::Get();
        d.SetSnapshot(md, driverOrder, executorOrder);
        std::vector<int32_t> nodes((size_t)std::max(count, 1));
        int32_t has = 0, driver = -1;
        d.check(gp_pack_one(d.ctx(), algo, drv.CPU, drv.Memory, drv.NvidiaGPU, exe.CPU, exe.Memory, exe.NvidiaGPU, count, &has,
                            &driver, nodes.data()),
                "gp_pack_one");
        binpack::PackingResult r;
        if (!has) return r;                       // EmptyPackingResult
        r.HasCapacity = true;
        r.DriverNode = d.name(driver);
        r.ExecutorNodes.reserve((size_t)count);
        for (int i = 0; i < count; ++i) r.ExecutorNodes.push_back(d.name(nodes[(size_t)i]));
        return r;
    } catch (const Error&) {
        auto f = Fallbacks().find((int)algo);
        if (f != Fallbacks().end() && f->second) return f->second(drv, exe, count, driverOrder, executorOrder, md);
        throw;
    }
}

}  // namespace gangpack

namespace binpack {
// binpack.TightlyPack (pack_tightly.go:25-32) / binpack.DistributeEvenly (distribute_evenly.go:25-32)
inline const SparkBinPackFunction TightlyPack = [](const resources::Resources& d, const resources::Resources& e, int c,
                                                   const std::vector<std::string>& dord, const std::vector<std::string>& eord,
                                                   const resources::NodeGroupSchedulingMetadata& md) {
    return gangpack::PackOne(GP_TIGHTLY_PACK, d, e, c, dord, eord, md);
};
inline const SparkBinPackFunction DistributeEvenly = [](const resources::Resources& d, const resources::Resources& e, int c,
                                                        const std::vector<std::string>& dord, const std::vector<std::string>& eord,
                                                        const resources::NodeGroupSchedulingMetadata& md) {
    return gangpack::PackOne(GP_DISTRIBUTE_EVENLY, d, e, c, dord, eord, md);
};
// binpack.MinimalFragmentation (minimal_fragmentation.go:27-35)
inline const SparkBinPackFunction MinimalFragmentation = [](const resources::Resources& d, const resources::Resources& e, int c,
                                                            const std::vector<std::string>& dord, const std::vector<std::string>& eord,
                                                            const resources::NodeGroupSchedulingMetadata& md) {
    return gangpack::PackOne(GP_MINIMAL_FRAGMENTATION, d, e, c, dord, eord, md);
};
}  // namespace binpack

// ------------------------------------------------------------------------------------------------
// Packing efficiencies (LIB/binpack/efficiency.go) and the zone-aware tightly-pack variants
// (SURVEY §8f row f3).  Host-side float64, same operation order as the Go code; the placements
// themselves come from the device: one instance group per zone, one independent decision per zone.
namespace binpack {

struct PackingEfficiency {                 // efficiency.go:51-56
    std::string NodeName;
    double CPU = 0, Memory = 0, GPU = 0;
    double Max() const { return std::max(GPU, std::max(CPU, Memory)); }           // :59-61
};
struct AvgPackingEfficiency {              // efficiency.go:24-29
    double CPU = 0, Memory = 0, GPU = 0, Max = 0;
    bool LessThan(const AvgPackingEfficiency& o) const { return Max < o.Max; }    // :33-35
};
inline AvgPackingEfficiency WorstAvgPackingEfficiency() { return AvgPackingEfficiency{}; }   // :38-45

// Quantity.Value() of a milli-scaled CPU: whole cores, inexact values rounded away from zero
// (k8s apimachinery quantity.go:732-734 -> math.go:166-199)
inline int64_t cpuValue(int64_t milli) {
    int64_t q = milli / 1000, rem = milli % 1000;
    return rem > 0 ? q + 1 : (rem < 0 ? q - 1 : q);
}
inline int64_t normalizeResource(int64_t v) { return v == 0 ? 1 : v; }            // :104-109

// computePackingEfficiency, efficiency.go:79-102
inline PackingEfficiency computePackingEfficiency(const std::string& nodeName, const resources::NodeSchedulingMetadata& m,
                                                  const resources::NodeGroupResources& reserved) {
    resources::Resources r = m.SchedulableResources;
    r.Sub(m.AvailableResources);
    auto it = reserved.find(nodeName);
    if (it != reserved.end()) r.Add(it->second);
    PackingEfficiency e;
    e.NodeName = nodeName;
    if (m.SchedulableResources.NvidiaGPU != 0)
        e.GPU = (double)r.NvidiaGPU / (double)normalizeResource(m.SchedulableResources.NvidiaGPU);
    e.CPU = (double)cpuValue(r.CPU) / (double)normalizeResource(cpuValue(m.SchedulableResources.CPU));
    e.Memory = (double)r.Memory / (double)normalizeResource(m.SchedulableResources.Memory);
    return e;
}

// ComputeAvgPackingEfficiency, efficiency.go:114-156
inline AvgPackingEfficiency ComputeAvgPackingEfficiency(const resources::NodeGroupSchedulingMetadata& md,
                                                        const std::vector<PackingEfficiency>& effs) {
    if (effs.empty()) return WorstAvgPackingEfficiency();
    double cpuSum = 0, gpuSum = 0, memorySum = 0, maxSum = 0;
    int nodesWithGPU = 0;
    for (const auto& e : effs) {
        cpuSum += e.CPU;
        memorySum += e.Memory;
        if (md.at(e.NodeName).SchedulableResources.NvidiaGPU != 0) { gpuSum += e.GPU; nodesWithGPU++; }
        maxSum += e.Max();
    }
    double length = std::max((double)effs.size(), 1.0);
    AvgPackingEfficiency a;
    a.CPU = cpuSum / length; a.Memory = memorySum / length;
    a.GPU = nodesWithGPU == 0 ? 1.0 : gpuSum / (double)nodesWithGPU;
    a.Max = maxSum / length;
    return a;
}

// reserved map of SparkBinPack (binpack.go:72-75) rebuilt from a placement
inline resources::NodeGroupResources ReservedOf(const resources::Resources& drv, const resources::Resources& exe,
                                                const PackingResult& r) {
    resources::NodeGroupResources reserved;
    reserved[r.DriverNode] = drv;
    for (const auto& n : r.ExecutorNodes) reserved[n].Add(exe);
    return reserved;
}

// groupNodesByZone, single_az.go:57-73
inline void groupNodesByZone(const std::vector<std::string>& nodeNames, const resources::NodeGroupSchedulingMetadata& md,
                             std::vector<std::string>* zonesInOrder, std::unordered_map<std::string, std::vector<std::string>>* byZone) {
    for (const auto& n : nodeNames) {
        auto m = md.find(n);
        if (m == md.end()) continue;
        auto it = byZone->find(m->second.ZoneLabel);
        if (it == byZone->end()) { zonesInOrder->push_back(m->second.ZoneLabel); it = byZone->emplace(m->second.ZoneLabel, std::vector<std::string>{}).first; }
        it->second.push_back(n);
    }
}

// getSingleAZSparkBinFunction(fn) + chooseBestResult, single_az.go:23-55,75-97, fn = tightlyPackExecutors or
// minimalFragmentation: every candidate zone is packed in ONE device batch (zone = instance group), the best result
// is chosen on the host by average packing efficiency over [driver] + ExecutorNodes.  The efficiencies come from
// SparkBinPack's `reserved` map (binpack.go:72-77): tightlyPackExecutors adds every executor to it,
// minimalFragmentation never touches it -> for that packer the map holds the driver only (kept as is).
inline PackingResult SingleAZPackImpl(gp_algo algo, const resources::Resources& drv, const resources::Resources& exe, int count,
                                      const std::vector<std::string>& driverOrder, const std::vector<std::string>& executorOrder,
                                      const resources::NodeGroupSchedulingMetadata& md) {
    std::vector<std::string> dzOrder, ezOrder;
    std::unordered_map<std::string, std::vector<std::string>> dz, ez;
    groupNodesByZone(driverOrder, md, &dzOrder, &dz);
    groupNodesByZone(executorOrder, md, &ezOrder, &ez);
    std::vector<std::pair<std::vector<std::string>, std::vector<std::string>>> groups;
    for (const auto& z : dzOrder)
        if (ez.count(z)) groups.emplace_back(dz[z], ez[z]);                         // :36-41
    if (groups.empty()) return EmptyPackingResult();
    gangpack::Device& d = gangpack::Device::Get();
    d.SetSnapshotGroups(md, groups);
    const size_t Z = groups.size();
    const int64_t k = count > 0 ? count : 0;
    std::vector<int64_t> dc(Z, drv.CPU), dm(Z, drv.Memory), dg(Z, drv.NvidiaGPU), ec(Z, exe.CPU), em(Z, exe.Memory), eg(Z, exe.NvidiaGPU), off(Z + 1);
    std::vector<int32_t> cnt(Z, count), grp(Z), driver(Z, -1), exec((size_t)std::max<int64_t>(k * (int64_t)Z, 1));
    for (size_t z = 0; z < Z; ++z) { grp[z] = (int32_t)z; off[z] = k * (int64_t)z; }
    off[Z] = k * (int64_t)Z;
    gp_apps a{};
    a.n_apps = (int32_t)Z;
    a.drv_cpu_milli = dc.data(); a.drv_mem_bytes = dm.data(); a.drv_gpu = dg.data();
    a.exe_cpu_milli = ec.data(); a.exe_mem_bytes = em.data(); a.exe_gpu = eg.data();
    a.exe_count = cnt.data(); a.group = grp.data(); a.exec_out_off = off.data();
    gp_results r{};
    r.driver_node = driver.data(); r.executor_nodes = exec.data(); r.executor_nodes_cap = (int64_t)exec.size();
    d.check(gp_pack_batch(d.ctx(), &a, algo, GP_MODE_INDEPENDENT, &r), "gp_pack_batch");
    PackingResult best = EmptyPackingResult();                                      // :79
    AvgPackingEfficiency bestAvg = WorstAvgPackingEfficiency();                     // :80
    for (size_t z = 0; z < Z; ++z) {
        if (driver[z] < 0) continue;                                                // :44-46
        PackingResult res;
        res.HasCapacity = true;
        res.DriverNode = d.name(driver[z]);
        for (int64_t t = off[z]; t < off[z + 1]; ++t) res.ExecutorNodes.push_back(d.name(exec[(size_t)t]));
        resources::NodeGroupResources reserved;
        if (algo == GP_MINIMAL_FRAGMENTATION) reserved[res.DriverNode] = drv;
        else reserved = ReservedOf(drv, exe, res);
        std::vector<PackingEfficiency> effs;                                        // :83-89: [driver] + executors, duplicates kept
        effs.push_back(computePackingEfficiency(res.DriverNode, md.at(res.DriverNode), reserved));
        for (const auto& n : res.ExecutorNodes) effs.push_back(computePackingEfficiency(n, md.at(n), reserved));
        AvgPackingEfficiency avg = ComputeAvgPackingEfficiency(md, effs);
        if (bestAvg.LessThan(avg)) { best = res; bestAvg = avg; }                   // :91-94
    }
    return best;
}

inline PackingResult GuardedSingleAZ(bool azAware, const resources::Resources& d, const resources::Resources& e, int c,
                                     const std::vector<std::string>& dord, const std::vector<std::string>& eord,
                                     const resources::NodeGroupSchedulingMetadata& md, gp_algo algo = GP_TIGHTLY_PACK) {
    const int key = algo == GP_MINIMAL_FRAGMENTATION ? 102 : (azAware ? 101 : 100);
    try {
        PackingResult r = SingleAZPackImpl(algo, d, e, c, dord, eord, md);
        if (r.HasCapacity || !azAware) return r;
        return gangpack::PackOne(GP_TIGHTLY_PACK, d, e, c, dord, eord, md);          // az_aware_pack_tightly.go:33-37
    } catch (const gangpack::Error&) {
        auto f = gangpack::Fallbacks().find(key);
        if (f != gangpack::Fallbacks().end() && f->second) return f->second(d, e, c, dord, eord, md);
        throw;
    }
}

// binpack.SingleAZTightlyPack (single_az_pack_tightly.go) / binpack.AzAwareTightlyPack (az_aware_pack_tightly.go:27-38)
inline const SparkBinPackFunction SingleAZTightlyPack = [](const resources::Resources& d, const resources::Resources& e, int c,
                                                           const std::vector<std::string>& dord, const std::vector<std::string>& eord,
                                                           const resources::NodeGroupSchedulingMetadata& md) {
    return GuardedSingleAZ(false, d, e, c, dord, eord, md);
};
inline const SparkBinPackFunction AzAwareTightlyPack = [](const resources::Resources& d, const resources::Resources& e, int c,
                                                          const std::vector<std::string>& dord, const std::vector<std::string>& eord,
                                                          const resources::NodeGroupSchedulingMetadata& md) {
    return GuardedSingleAZ(true, d, e, c, dord, eord, md);
};
// binpack.SingleAZMinimalFragmentation (single_az_minimal_fragmentation.go:20)
inline const SparkBinPackFunction SingleAZMinimalFragmentation = [](const resources::Resources& d, const resources::Resources& e, int c,
                                                                    const std::vector<std::string>& dord, const std::vector<std::string>& eord,
                                                                    const resources::NodeGroupSchedulingMetadata& md) {
    return GuardedSingleAZ(false, d, e, c, dord, eord, md, GP_MINIMAL_FRAGMENTATION);
};

}  // namespace binpack

// ------------------------------------------------------------------------------------------------
namespace binpacker {

struct Binpacker {                        // internal/binpacker/binpack.go:37-41
    std::string Name;
    binpack::SparkBinPackFunction BinpackFunc;
    bool IsSingleAz;
    int algo;                             // gp_algo behind BinpackFunc
};

inline const char* const tightlyPack = "tightly-pack";            // :23
inline const char* const distributeEvenly = "distribute-evenly";  // :22
inline const char* const azAwareTightlyPack = "az-aware-tightly-pack";      // :24
inline const char* const SingleAzTightlyPack = "single-az-tightly-pack";    // :29
inline const char* const SingleAzMinimalFragmentation = "single-az-minimal-fragmentation";   // :33

// binpackFunctions (:43-49): all five `binpack:` values.  `algo` is the device packer behind the function; the
// zone-aware ones (needsHostLoop) choose a zone per application on the host, so the FIFO loop calls them per driver.
inline const std::map<std::string, Binpacker>& binpackFunctions() {
    static const std::map<std::string, Binpacker> m = {
        {tightlyPack, {tightlyPack, binpack::TightlyPack, false, GP_TIGHTLY_PACK}},
        {distributeEvenly, {distributeEvenly, binpack::DistributeEvenly, false, GP_DISTRIBUTE_EVENLY}},
        {azAwareTightlyPack, {azAwareTightlyPack, binpack::AzAwareTightlyPack, false, GP_TIGHTLY_PACK}},
        {SingleAzTightlyPack, {SingleAzTightlyPack, binpack::SingleAZTightlyPack, true, GP_TIGHTLY_PACK}},
        {SingleAzMinimalFragmentation, {SingleAzMinimalFragmentation, binpack::SingleAZMinimalFragmentation, true, GP_MINIMAL_FRAGMENTATION}},
    };
    return m;
}

// packers whose placement is not ONE device packer over ONE pair of orders (they pick a zone on the host)
inline bool NeedsHostLoop(const Binpacker& b) { return b.Name != tightlyPack && b.Name != distributeEvenly; }

// SelectBinpacker (:52-58): unknown names select distribute-evenly, exactly like the reference.
inline const Binpacker* SelectBinpacker(const std::string& name) {
    auto it = binpackFunctions().find(name);
    if (it == binpackFunctions().end()) return &binpackFunctions().at(distributeEvenly);
    return &it->second;
}

}  // namespace binpacker

// ------------------------------------------------------------------------------------------------
namespace sort {

struct LabelPriorityOrder {               // config.LabelPriorityOrder, config/config.go:79-84
    std::string Name;
    std::vector<std::string> DescendingPriorityValues;
};

// NodeSorter, internal/sort/nodesorting.go:25-64.  Host-side (the step before the hot path).
class NodeSorter {
public:
    NodeSorter(const LabelPriorityOrder* driverLabel = nullptr, const LabelPriorityOrder* executorLabel = nullptr) {
        if (driverLabel) driver_ = std::make_unique<LabelPriorityOrder>(*driverLabel);
        if (executorLabel) executor_ = std::make_unique<LabelPriorityOrder>(*executorLabel);
    }

    // PotentialNodes (:41-64)
    void PotentialNodes(const resources::NodeGroupSchedulingMetadata& md, const std::vector<std::string>& nodeNames,
                        std::vector<std::string>* driverNodes, std::vector<std::string>* executorNodes) const {
        std::vector<std::string> order = getNodeNamesInPriorityOrder(md);
        std::unordered_map<std::string, int> cand;
        for (const auto& n : nodeNames) cand.emplace(n, 0);
        driverNodes->clear(); executorNodes->clear();
        for (const auto& n : order) {
            if (cand.count(n)) driverNodes->push_back(n);                                   // :52-54
            const auto& m = md.at(n);
            if (!m.Unschedulable && m.Ready) executorNodes->push_back(n);                   // :55-57
        }
        sortByLabel(*driverNodes, md, driver_.get());                                       // :61
        sortByLabel(*executorNodes, md, executor_.get());                                   // :62
    }

    // resourcesLessThan (:74-80)
    static bool resourcesLessThan(const resources::Resources& l, const resources::Resources& r) {
        if (l.Memory != r.Memory) return l.Memory < r.Memory;
        return l.CPU < r.CPU;
    }

    // getNodeNamesInPriorityOrder (:95-122); ties the reference leaves to an unstable sort are broken
    // by zone label / stable order here.
    static std::vector<std::string> getNodeNamesInPriorityOrder(const resources::NodeGroupSchedulingMetadata& md) {
        std::map<std::string, resources::Resources> az;                                     // :124-134
        for (const auto& kv : md) az[kv.second.ZoneLabel].Add(kv.second.AvailableResources);
        std::vector<std::string> labels;
        for (const auto& kv : az) labels.push_back(kv.first);
        std::stable_sort(labels.begin(), labels.end(),
                         [&](const std::string& a, const std::string& b) { return resourcesLessThan(az[a], az[b]); });  // :102-104
        std::unordered_map<std::string, int> prio;
        for (size_t i = 0; i < labels.size(); ++i) prio[labels[i]] = (int)i;
        std::vector<std::string> names;
        names.reserve(md.size());
        for (const auto& kv : md) names.push_back(kv.first);
        std::sort(names.begin(), names.end());   // deterministic starting order (Go: map order)
        std::stable_sort(names.begin(), names.end(), [&](const std::string& a, const std::string& b) {       // :83-93, :117-119
            const auto &ma = md.at(a), &mb = md.at(b);
            int pa = prio[ma.ZoneLabel], pb = prio[mb.ZoneLabel];
            if (pa != pb) return pa < pb;
            if (!ma.AvailableResources.Eq(mb.AvailableResources)) return resourcesLessThan(ma.AvailableResources, mb.AvailableResources);
            return a < b;
        });
        return names;
    }

private:
    // createLabelLessThanFunction + sortNodesByMetadataLessThanFunction (:161-200)
    static void sortByLabel(std::vector<std::string>& names, const resources::NodeGroupSchedulingMetadata& md,
                            const LabelPriorityOrder* cfg) {
        if (!cfg) return;
        std::unordered_map<std::string, int> rank;
        for (size_t i = 0; i < cfg->DescendingPriorityValues.size(); ++i) rank[cfg->DescendingPriorityValues[i]] = (int)i;
        auto rank_of = [&](const std::string& n, int* r) {
            const auto& labels = md.at(n).AllLabels;
            auto v = labels.find(cfg->Name);
            if (v == labels.end()) return false;
            auto k = rank.find(v->second);
            if (k == rank.end()) return false;
            *r = k->second;
            return true;
        };
        std::stable_sort(names.begin(), names.end(), [&](const std::string& a, const std::string& b) {
            int ra = 0, rb = 0;
            if (!rank_of(a, &ra)) return false;
            if (!rank_of(b, &rb)) return true;
            return ra < rb;
        });
    }
    std::unique_ptr<LabelPriorityOrder> driver_, executor_;
};

}  // namespace sort

// ------------------------------------------------------------------------------------------------
namespace extender {

struct SparkApplicationResources {        // internal/types/types.go:22-27
    resources::Resources DriverResources;
    resources::Resources ExecutorResources;
    int MinExecutorCount = 0;
    int MaxExecutorCount = 0;
};

struct PendingDriver {                    // what fitEarlierDrivers reads off a queued driver pod
    std::string Name;
    SparkApplicationResources Resources;
    bool ParseError = false;              // sparkResources failed -> skipped (resource.go:232-237)
    bool SkipIfNoFit = false;             // shouldSkipDriverFifo (resource.go:264-270): younger than enforce-after-pod-age
};

// One queued driver pod -> what fitEarlierDrivers needs of it: sparkResources (sparkpods.go:73-137) over its annotations
// (spark_resources.hpp).  A parse error marks the entry like resource.go:232-237 does (skipped); a quantity outside the
// exact-int64 model is reported through *exact = false so that the embedding runs the original Go loop instead.
inline PendingDriver PendingDriverFromAnnotations(const std::string& name, const std::map<std::string, std::string>& annotations,
                                                  bool skipIfNoFit, bool* exact = nullptr) {
    PendingDriver d;
    d.Name = name;
    d.SkipIfNoFit = skipIfNoFit;
    ParsedSparkResources r;
    d.ParseError = !SparkResources(annotations, &r).empty();
    if (exact) *exact = r.Exact;
    if (!d.ParseError) {
        d.Resources.DriverResources = resources::CreateResources(r.DriverCPUMilli, r.DriverMemoryBytes, r.DriverNvidiaGPUs);
        d.Resources.ExecutorResources = resources::CreateResources(r.ExecutorCPUMilli, r.ExecutorMemoryBytes, r.ExecutorNvidiaGPUs);
        d.Resources.MinExecutorCount = (int)r.MinExecutorCount;
        d.Resources.MaxExecutorCount = (int)r.MaxExecutorCount;
    }
    return d;
}

// sparkResourceUsage (sparkpods.go:139-146): ASSIGNMENT semantics, kept bug-for-bug.
inline resources::NodeGroupResources SparkResourceUsage(const resources::Resources& drv, const resources::Resources& exe,
                                                        const std::string& driverNode, const std::vector<std::string>& executorNodes) {
    resources::NodeGroupResources res;
    res[driverNode] = drv;
    for (const auto& n : executorNodes) res[n] = exe;
    return res;
}

// fitEarlierDrivers (resource.go:224-262) as ONE device batch: every earlier driver is packed in queue
// order against the mutating snapshot on the GPU (GP_MODE_FIFO_REFERENCE); on return `metadata` has the
// usage of every fitted driver subtracted, exactly as the Go loop leaves it.  `results` (optional)
// receives the per-driver packing results.  Returns false when a non-skippable driver does not fit.
inline bool FitEarlierDrivers(const binpacker::Binpacker& packer, const std::vector<PendingDriver>& drivers,
                              const std::vector<std::string>& nodeNames, const std::vector<std::string>& executorNodeNames,
                              resources::NodeGroupSchedulingMetadata& metadata,
                              std::vector<binpack::PackingResult>* results = nullptr) {
    std::vector<size_t> live;   // parse errors are skipped before the packer is called (:232-237)
    for (size_t i = 0; i < drivers.size(); ++i)
        if (!drivers[i].ParseError) live.push_back(i);
    if (results) results->assign(drivers.size(), binpack::EmptyPackingResult());
    if (live.empty()) return true;
    if (packer.IsSingleAz) {
        // single-AZ packers: the whole queue in ONE launch (gp_pack_fifo_zones) -- zone = instance group, the zone choice of
        // driver i feeds driver i+1 on the device
        std::vector<std::string> dzOrder, ezOrder;
        std::unordered_map<std::string, std::vector<std::string>> dz, ez;
        binpack::groupNodesByZone(nodeNames, metadata, &dzOrder, &dz);
        binpack::groupNodesByZone(executorNodeNames, metadata, &ezOrder, &ez);
        std::vector<std::pair<std::vector<std::string>, std::vector<std::string>>> groups;
        for (const auto& z : dzOrder)
            if (ez.count(z)) groups.emplace_back(dz[z], ez[z]);                 // single_az.go:36-41
        const size_t q = live.size();
        if (groups.empty()) {                                                   // EmptyPackingResult for every driver
            for (size_t j = 0; j < q; ++j) if (!drivers[live[j]].SkipIfNoFit) return false;
            return true;
        }
        gangpack::Device& d = gangpack::Device::Get();
        d.SetSnapshotGroups(metadata, groups);
        d.SetSchedulable(metadata);
        std::vector<int64_t> dc(q), dm(q), dg(q), ec(q), em(q), eg(q);
        std::vector<int32_t> cnt(q), zone(q, -1), driver(q, -1);
        std::vector<uint8_t> skip(q);
        int64_t total = 0;
        for (size_t j = 0; j < q; ++j) {
            const auto& a = drivers[live[j]];
            dc[j] = a.Resources.DriverResources.CPU; dm[j] = a.Resources.DriverResources.Memory; dg[j] = a.Resources.DriverResources.NvidiaGPU;
            ec[j] = a.Resources.ExecutorResources.CPU; em[j] = a.Resources.ExecutorResources.Memory; eg[j] = a.Resources.ExecutorResources.NvidiaGPU;
            cnt[j] = a.Resources.MinExecutorCount;
            skip[j] = a.SkipIfNoFit ? 1 : 0;
            total += std::max(cnt[j], 0);
        }
        std::vector<int32_t> exec((size_t)std::max<int64_t>(total, 1));
        gp_apps a{};
        a.n_apps = (int32_t)q;
        a.drv_cpu_milli = dc.data(); a.drv_mem_bytes = dm.data(); a.drv_gpu = dg.data();
        a.exe_cpu_milli = ec.data(); a.exe_mem_bytes = em.data(); a.exe_gpu = eg.data();
        a.exe_count = cnt.data(); a.skip_if_no_fit = skip.data();
        gp_zone_results r{};
        r.zone = zone.data(); r.driver_node = driver.data(); r.executor_nodes = exec.data(); r.executor_nodes_cap = (int64_t)exec.size();
        d.check(gp_pack_fifo_zones(d.ctx(), &a, (gp_algo)packer.algo, GP_MODE_FIFO_REFERENCE, &r), "gp_pack_fifo_zones");
        std::vector<int64_t> cpu(d.n_nodes()), mem(d.n_nodes()), gpu(d.n_nodes());
        d.check(gp_get_snapshot(d.ctx(), cpu.data(), mem.data(), gpu.data()), "gp_get_snapshot");
        for (size_t i = 0; i < d.n_nodes(); ++i) {
            auto& m = metadata.at(d.name((int32_t)i)).AvailableResources;
            m.CPU = cpu[i]; m.Memory = mem[i]; m.NvidiaGPU = gpu[i];
        }
        bool ok = true;
        int64_t at = 0;
        for (size_t j = 0; j < q; ++j) {
            if (driver[j] == -2 || (driver[j] == -1 && !skip[j])) ok = false;   // :250-252
            if (results && driver[j] >= 0) {
                auto& pr = (*results)[live[j]];
                pr.HasCapacity = true;
                pr.DriverNode = d.name(driver[j]);
                for (int64_t t = 0; t < std::max(cnt[j], 0); ++t) pr.ExecutorNodes.push_back(d.name(exec[(size_t)(at + t)]));
            }
            at += std::max(cnt[j], 0);
        }
        return ok;
    }
    if (binpacker::NeedsHostLoop(packer)) {
        // az-aware-tightly-pack: the reference's own loop (:230-259), every BinpackFunc call being one device batch over
        // the zones (+ the undivided fallback); usage is subtracted on the host between calls (SubtractUsageIfExists,
        // resources.go:129-135)
        for (size_t i : live) {
            const auto& a = drivers[i];
            binpack::PackingResult pr = packer.BinpackFunc(a.Resources.DriverResources, a.Resources.ExecutorResources,
                                                           a.Resources.MinExecutorCount, nodeNames, executorNodeNames, metadata);
            if (!pr.HasCapacity) {
                if (a.SkipIfNoFit) continue;                                   // :245-249
                return false;                                                  // :250-252
            }
            for (const auto& u : SparkResourceUsage(a.Resources.DriverResources, a.Resources.ExecutorResources, pr.DriverNode, pr.ExecutorNodes)) {
                auto m = metadata.find(u.first);
                if (m != metadata.end()) m->second.AvailableResources.Sub(u.second);
            }
            if (results) (*results)[i] = pr;
        }
        return true;
    }
    gangpack::Device& d = gangpack::Device::Get();
    d.SetSnapshot(metadata, nodeNames, executorNodeNames);
    const size_t q = live.size();
    std::vector<int64_t> dc(q), dm(q), dg(q), ec(q), em(q), eg(q), off(q + 1, 0);
    std::vector<int32_t> cnt(q);
    std::vector<uint8_t> skip(q);
    for (size_t j = 0; j < q; ++j) {
        const auto& a = drivers[live[j]];
        dc[j] = a.Resources.DriverResources.CPU; dm[j] = a.Resources.DriverResources.Memory; dg[j] = a.Resources.DriverResources.NvidiaGPU;
        ec[j] = a.Resources.ExecutorResources.CPU; em[j] = a.Resources.ExecutorResources.Memory; eg[j] = a.Resources.ExecutorResources.NvidiaGPU;
        cnt[j] = a.Resources.MinExecutorCount;                     // MIN count enters the packer (:242)
        skip[j] = a.SkipIfNoFit ? 1 : 0;
        off[j + 1] = off[j] + std::max(cnt[j], 0);
    }
    std::vector<int32_t> driver(q, -1), exec((size_t)std::max<int64_t>(off[q], 1));
    gp_apps a{};
    a.n_apps = (int32_t)q;
    a.drv_cpu_milli = dc.data(); a.drv_mem_bytes = dm.data(); a.drv_gpu = dg.data();
    a.exe_cpu_milli = ec.data(); a.exe_mem_bytes = em.data(); a.exe_gpu = eg.data();
    a.exe_count = cnt.data(); a.skip_if_no_fit = skip.data(); a.exec_out_off = off.data();
    gp_results r{};
    r.driver_node = driver.data(); r.executor_nodes = exec.data(); r.executor_nodes_cap = (int64_t)exec.size();
    d.check(gp_pack_batch(d.ctx(), &a, (gp_algo)packer.algo, GP_MODE_FIFO_REFERENCE, &r), "gp_pack_batch");
    // pull the charged snapshot back into the caller's metadata (SubtractUsageIfExists, resources.go:129-135)
    std::vector<int64_t> cpu(d.n_nodes()), mem(d.n_nodes()), gpu(d.n_nodes());
    d.check(gp_get_snapshot(d.ctx(), cpu.data(), mem.data(), gpu.data()), "gp_get_snapshot");
    for (size_t i = 0; i < d.n_nodes(); ++i) {
        auto& m = metadata.at(d.name((int32_t)i)).AvailableResources;
        m.CPU = cpu[i]; m.Memory = mem[i]; m.NvidiaGPU = gpu[i];
    }
    bool ok = true;
    for (size_t j = 0; j < q; ++j) {
        if (driver[j] == -2 || (driver[j] == -1 && !skip[j])) ok = false;   // :250-252
        if (results && driver[j] >= 0) {
            auto& pr = (*results)[live[j]];
            pr.HasCapacity = true;
            pr.DriverNode = d.name(driver[j]);
            for (int64_t t = off[j]; t < off[j + 1]; ++t) pr.ExecutorNodes.push_back(d.name(exec[(size_t)t]));
        }
    }
    return ok;
}

// The node choice of rescheduleExecutor (resource.go:652-662, 675-705) on the device.  `availableNodesSchedulingMetadata`
// is the metadata of :640 (minimal-fragmentation branch), `availableResources` the map of :643 (first-fit branch),
// `overhead` the map the reference hands to GetNodeCapacities as "reserved" (:682).  Returns (node, ok) like
// rescheduleExecutorWithMinimalFragmentation; ok == false is failureFit (:672).
inline std::pair<std::string, bool> RescheduleExecutorNode(const binpacker::Binpacker& packer, const resources::Resources& executorResources,
                                                           const std::vector<std::string>& executorNodeNames,
                                                           const resources::NodeGroupSchedulingMetadata& availableNodesSchedulingMetadata,
                                                           const resources::NodeGroupResources& availableResources,
                                                           const resources::NodeGroupResources& overhead,
                                                           const std::set<std::string>& nodesWithExecutorsBelongingToThisApp) {
    gangpack::Device& d = gangpack::Device::Get();
    const bool minFrag = packer.Name == binpacker::SingleAzMinimalFragmentation;           // :652
    if (minFrag) {
        d.SetSnapshot(availableNodesSchedulingMetadata, {}, executorNodeNames);
    } else {
        resources::NodeGroupSchedulingMetadata md;
        for (const auto& kv : availableResources) md[kv.first].AvailableResources = kv.second;
        d.SetSnapshot(md, {}, executorNodeNames);
    }
    const size_t N = d.n_nodes();
    int64_t ec = executorResources.CPU, em = executorResources.Memory, eg = executorResources.NvidiaGPU;
    std::vector<int64_t> rc(N, 0), rm(N, 0), rg(N, 0);
    std::vector<int32_t> hosting;
    int64_t hoff[2] = {0, 0};
    gp_reschedule in{};
    in.n_execs = 1; in.exe_cpu_milli = &ec; in.exe_mem_bytes = &em; in.exe_gpu = &eg; in.min_frag = minFrag ? 1 : 0;
    if (minFrag) {
        for (const auto& kv : overhead) {
            int32_t i = d.index(kv.first);
            if (i >= 0) { rc[(size_t)i] = kv.second.CPU; rm[(size_t)i] = kv.second.Memory; rg[(size_t)i] = kv.second.NvidiaGPU; }
        }
        for (const auto& n : nodesWithExecutorsBelongingToThisApp) { int32_t i = d.index(n); if (i >= 0) hosting.push_back(i); }
        hoff[1] = (int64_t)hosting.size();
        in.reserved_cpu_milli = rc.data(); in.reserved_mem_bytes = rm.data(); in.reserved_gpu = rg.data();
        in.host_off = hoff; in.host_nodes = hosting.data();
    }
    int32_t node = -1;
    if (N == 0) return {std::string(), false};
    d.check(gp_reschedule_executors(d.ctx(), &in, &node), "gp_reschedule_executors");
    if (node < 0) return {std::string(), false};
    return {d.name(node), true};
}

}  // namespace extender

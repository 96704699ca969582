// CPU-only tests of the host layer's input side (k8s-spark-scheduler_b200/host/spark_resources.hpp): the driver
// annotations -> application tuple step of the reference (internal/extender/sparkpods.go:73-137), written like its own
// sparkpods_test.go:38-117.  No device, no libgangpack.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "spark_resources.hpp"

static int failures = 0;
#define EXPECT(cond, msg)                                                            \
    do {                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); ++failures; } \
    } while (0)

using extender::ParsedSparkResources;
using extender::SparkResources;
typedef std::map<std::string, std::string> Annotations;
static const int64_t Mi = 1ll << 20;

static void TestSparkResources() {   // sparkpods_test.go:38-117, the three table entries
    ParsedSparkResources r;
    Annotations staticApp = {{"spark-driver-cpu", "1"}, {"spark-driver-mem", "2432Mi"}, {"spark-driver-nvidia.com/gpu", "1"},
                             {"spark-executor-cpu", "2"}, {"spark-executor-mem", "6758Mi"}, {"spark-executor-nvidia.com/gpu", "1"},
                             {"spark-executor-count", "2"}};
    EXPECT(SparkResources(staticApp, &r).empty(), "parses static allocation pod annotations into resources");
    EXPECT(r.Exact && r.DriverCPUMilli == 1000 && r.DriverMemoryBytes == 2432 * Mi && r.DriverNvidiaGPUs == 1, "driver resources");
    EXPECT(r.ExecutorCPUMilli == 2000 && r.ExecutorMemoryBytes == 6758 * Mi && r.ExecutorNvidiaGPUs == 1, "executor resources");
    EXPECT(r.MinExecutorCount == 2 && r.MaxExecutorCount == 2, "minExecutorCount == maxExecutorCount == ExecutorCount in static allocation");

    Annotations dynamicApp = staticApp;
    dynamicApp.erase("spark-executor-count");
    dynamicApp["spark-dynamic-allocation-enabled"] = "true";
    dynamicApp["spark-dynamic-allocation-min-executor-count"] = "2";
    dynamicApp["spark-dynamic-allocation-max-executor-count"] = "5";
    EXPECT(SparkResources(dynamicApp, &r).empty(), "parses dynamic allocation pod annotations into resources");
    EXPECT(r.MinExecutorCount == 2 && r.MaxExecutorCount == 5 && r.ExecutorMemoryBytes == 6758 * Mi, "dynamic allocation counts");

    Annotations noGpu = {{"spark-driver-cpu", "1"}, {"spark-driver-mem", "2432Mi"}, {"spark-executor-cpu", "2"},
                         {"spark-executor-mem", "6758Mi"}, {"spark-executor-count", "2"}};
    EXPECT(SparkResources(noGpu, &r).empty(), "parses static allocation pod annotations when no gpu annotation is present");
    EXPECT(r.DriverNvidiaGPUs == 0 && r.ExecutorNvidiaGPUs == 0 && r.DriverCPUMilli == 1000, "gpu annotations are optional");
}

static void TestSparkResourcesErrors() {   // sparkpods.go:76-104: the error branches
    ParsedSparkResources r;
    Annotations a = {{"spark-driver-cpu", "1"}, {"spark-driver-mem", "1"}, {"spark-executor-cpu", "1"}, {"spark-executor-mem", "1"}};
    EXPECT(SparkResources(a, &r) == "annotation ExecutorCount is required when DynamicAllocationEnabled is false", "static allocation needs a count");
    a["spark-dynamic-allocation-enabled"] = "true";
    EXPECT(SparkResources(a, &r) == "annotation spark-dynamic-allocation-min-executor-count is required when DynamicAllocationEnabled is true",
           "dynamic allocation needs min");
    a["spark-dynamic-allocation-min-executor-count"] = "1";
    EXPECT(SparkResources(a, &r) == "annotation spark-dynamic-allocation-max-executor-count is required when DynamicAllocationEnabled is true",
           "dynamic allocation needs max");
    a["spark-dynamic-allocation-max-executor-count"] = "3";
    EXPECT(SparkResources(a, &r).empty() && r.MinExecutorCount == 1 && r.MaxExecutorCount == 3, "dynamic allocation without ExecutorCount");
    a["spark-dynamic-allocation-enabled"] = "maybe";
    EXPECT(SparkResources(a, &r) == "annotation DynamicAllocationEnabled could not be parsed as a boolean", "ParseBool");
    a["spark-dynamic-allocation-enabled"] = "false";
    a["spark-executor-count"] = "4";
    a.erase("spark-executor-mem");
    EXPECT(SparkResources(a, &r) == "annotation spark-executor-mem is missing from driver", "required annotation");
    a["spark-executor-mem"] = "4 GiB";
    EXPECT(SparkResources(a, &r) == "annotation spark-executor-mem does not have a parseable value 4 GiB", "unparseable value");
    a["spark-executor-mem"] = "1Gi";
    a["spark-executor-cpu"] = "0.0005";     // half a millicore: a valid Quantity, outside the exact-int64 model
    EXPECT(SparkResources(a, &r).empty() && !r.Exact, "sub-millicore requests are flagged, not rounded");
}

static void TestParseQuantity() {   // grammar of quantity.go:147-300, suffix.go:113-132
    using resource::ParseQuantityScaled;
    using resource::ParseStatus;
    struct Case { const char* s; int scale; ParseStatus st; int64_t v; };
    const Case cases[] = {
        {"0", 0, ParseStatus::Ok, 0},           {"1", 3, ParseStatus::Ok, 1000},           {"500m", 3, ParseStatus::Ok, 500},
        {"1.5", 3, ParseStatus::Ok, 1500},      {"0.1", 3, ParseStatus::Ok, 100},          {"100m", 0, ParseStatus::Unrepresentable, 0},
        {"2432Mi", 0, ParseStatus::Ok, 2432 * Mi}, {"1Gi", 0, ParseStatus::Ok, 1ll << 30},  {"1.5Gi", 0, ParseStatus::Ok, 3ll << 29},
        {"1G", 0, ParseStatus::Ok, 1000000000},  {"1k", 0, ParseStatus::Ok, 1000},          {"1Ki", 0, ParseStatus::Ok, 1024},
        {"1e3", 0, ParseStatus::Ok, 1000},       {"1E3", 0, ParseStatus::Ok, 1000},         {"12e-1", 3, ParseStatus::Ok, 1200},
        {"+7", 0, ParseStatus::Ok, 7},           {"-2", 3, ParseStatus::Ok, -2000},         {"007", 0, ParseStatus::Ok, 7},
        {"1.", 0, ParseStatus::Ok, 1},           {".5", 3, ParseStatus::Ok, 500},           {"1.G", 0, ParseStatus::Ok, 1000000000},
        {"1u", 3, ParseStatus::Unrepresentable, 0}, {"1n", 3, ParseStatus::Unrepresentable, 0}, {"1000u", 3, ParseStatus::Ok, 1},
        {"", 0, ParseStatus::ErrFormatWrong, 0}, {"1 Gi", 0, ParseStatus::ErrFormatWrong, 0}, {"abc", 0, ParseStatus::ErrFormatWrong, 0},
        {"1Zi", 0, ParseStatus::ErrFormatWrong, 0}, {"1Kii", 0, ParseStatus::ErrSuffix, 0}, {"1mm", 0, ParseStatus::ErrSuffix, 0},
        {"1e", 0, ParseStatus::ErrSuffix, 0},    {"8Ei", 0, ParseStatus::Unrepresentable, 0}, {"4E", 3, ParseStatus::Unrepresentable, 0},
        {"1Ei", 0, ParseStatus::Ok, 1ll << 60},  {"2Ei", 0, ParseStatus::Unrepresentable, 0},
    };
    for (const Case& c : cases) {
        int64_t v = -1;
        ParseStatus st = ParseQuantityScaled(c.s, c.scale, &v);
        if (st != c.st || (st == ParseStatus::Ok && v != c.v)) {
            std::printf("FAIL quantity '%s' scale %d: status %d value %lld (want %d, %lld)\n", c.s, c.scale, (int)st, (long long)v, (int)c.st,
                        (long long)c.v);
            ++failures;
        }
    }
    int64_t v = 0;   // Quantity.Value() rounds a fractional count away from zero (quantity.go:732-734)
    EXPECT(ParseQuantityScaled("2.5", 0, &v, true) == ParseStatus::Ok && v == 3, "Value() of 2.5 is 3");
    EXPECT(ParseQuantityScaled("2500m", 0, &v, true) == ParseStatus::Ok && v == 3, "Value() of 2500m is 3");
    EXPECT(ParseQuantityScaled("2", 0, &v, true) == ParseStatus::Ok && v == 2, "Value() of 2 is 2");
}

static extender::QueuedDriverPod Pod(int64_t seconds, const char* uid) {   // createPod, sparkpods_test.go:125-157
    extender::QueuedDriverPod p;
    p.UID = uid; p.CreationSeconds = seconds; p.HasInstanceGroup = true; p.InstanceGroup = "instance-group-foobar";
    return p;
}

static void TestIsEarliest() {   // sparkpods_test.go:174-226
    auto uids = [](const std::vector<extender::QueuedDriverPod>& v) { std::vector<std::string> u; for (auto& p : v) u.push_back(p.UID); return u; };
    typedef std::vector<std::string> U;
    using extender::FilterToEarliestAndSort;
    EXPECT(uids(FilterToEarliestAndSort(Pod(100, "1"), {Pod(101, "3"), Pod(150, "2"), Pod(100, "1")})) == U{}, "selects earliest unassigned");
    EXPECT(uids(FilterToEarliestAndSort(Pod(100, "1"), {Pod(101, "2")})) == U{}, "selects if earliest and not in cache");
    EXPECT(uids(FilterToEarliestAndSort(Pod(100, "1"), {Pod(101, "3"), Pod(99, "2"), Pod(100, "1")})) == U{"2"}, "does not select when not earliest");
    EXPECT(uids(FilterToEarliestAndSort(Pod(100, "1"), {Pod(99, "3"), Pod(101, "2")})) == U{"3"}, "does not select when not earliest and not in cache");
    // the other filters of sparkpods.go:59-64 and the ordering of :70-72
    auto scheduled = Pod(50, "s"); scheduled.NodeName = "n1";
    auto other = Pod(60, "o"); other.InstanceGroup = "another-group";
    auto dying = Pod(70, "d"); dying.Deleting = true;
    auto foreign = Pod(80, "f"); foreign.SchedulerName = "default-scheduler";
    EXPECT(uids(FilterToEarliestAndSort(Pod(100, "1"), {Pod(90, "b"), scheduled, other, dying, foreign, Pod(10, "a")})) == (U{"a", "b"}),
           "unscheduled, same scheduler, same instance group, not deleting; oldest first");
}

// CLI used by tests/test_host_cpp.py to cross-check this C++ restatement with the Python one (oracle/pyref.py):
//   host_cpu_test quantity <scale> <roundUp 0|1> <string>...   -> one "<status> <value>" line per string
//   host_cpu_test annotations k=v k=v ...                      -> error text, or "ok drv exe min max exact" numbers
static int Cli(int argc, char** argv) {
    const std::string mode = argv[1];
    if (mode == "quantity" && argc >= 4) {
        const int scale = std::atoi(argv[2]);
        const bool up = std::atoi(argv[3]) != 0;
        static const char* const names[] = {"ok", "ErrFormatWrong", "ErrSuffix", "ErrNumeric", "unrepresentable"};
        for (int i = 4; i < argc; ++i) {
            int64_t v = 0;
            resource::ParseStatus st = resource::ParseQuantityScaled(argv[i], scale, &v, up);
            std::printf("%s %lld\n", names[(int)st], (long long)v);
        }
        return 0;
    }
    if (mode == "annotations") {
        Annotations a;
        for (int i = 2; i < argc; ++i) {
            std::string kv = argv[i];
            size_t eq = kv.find('=');
            a[kv.substr(0, eq)] = eq == std::string::npos ? "" : kv.substr(eq + 1);
        }
        ParsedSparkResources r;
        std::string err = SparkResources(a, &r);
        if (!err.empty()) { std::printf("%s\n", err.c_str()); return 0; }
        std::printf("ok %lld %lld %lld %lld %lld %lld %lld %lld %d\n", (long long)r.DriverCPUMilli, (long long)r.DriverMemoryBytes,
                    (long long)r.DriverNvidiaGPUs, (long long)r.ExecutorCPUMilli, (long long)r.ExecutorMemoryBytes,
                    (long long)r.ExecutorNvidiaGPUs, (long long)r.MinExecutorCount, (long long)r.MaxExecutorCount, r.Exact ? 1 : 0);
        return 0;
    }
    std::printf("usage: host_cpu_test [quantity <scale> <roundUp> <str>... | annotations k=v ...]\n");
    return 2;
}

int main(int argc, char** argv) {
    if (argc > 1) return Cli(argc, argv);
    TestIsEarliest();
    TestSparkResources();
    TestSparkResourcesErrors();
    TestParseQuantity();
    if (failures) { std::printf("%d FAILED\n", failures); return 1; }
    std::printf("host_cpu_test: all passed\n");
    return 0;
}

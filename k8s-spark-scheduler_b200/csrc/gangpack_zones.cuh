// gangpack_zones.cuh -- SURVEY §8f row f3 on the device: the float64 packing efficiencies and chooseBestResult of the
// single-AZ packers.
//
// Reference (LIB = /root/reference/vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg):
//   computePackingEfficiency / ComputeAvgPackingEfficiency      LIB/binpack/efficiency.go:79-156
//   getSingleAZSparkBinFunction / chooseBestResult              LIB/binpack/single_az.go:23-55, 75-97
//   SparkBinPack's `reserved` map (what the efficiencies see)   LIB/binpack/binpack.go:72-77
//   Quantity.Value() of a milli quantity (ceil away from zero)  vendor/k8s.io/apimachinery/pkg/api/resource/quantity.go:732-734
//
// Every application has been packed once per candidate zone (zone = instance group; row = app * Z + z of an independent
// batch).  One warp per application: lane z walks row (app, z) SEQUENTIALLY in the reference's order -- [driver] +
// ExecutorNodes, duplicates kept -- so the float64 sums round exactly like the Go loop; the warp then takes the arg-max of
// AvgPackingEfficiency.Max with the reference's tie rule (strict LessThan against a running best that starts at 0.0:
// the first zone with the highest average wins, and nothing wins when every average is 0) and copies the winning row.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gp {

struct ZoneChooseIn {
    const long long* avail[3];        // node-table order
    const long long* sched[3];        // SchedulableResources, node-table order (gpu may be NULL = 0)
    const int64_t* drv[3];            // per application (device columns, int64)
    const int64_t* exe[3];
    const int32_t* count;
    const int64_t* out_off;           // [n_apps + 1] CSR of the chosen placements
    const int32_t* row_driver;        // [n_apps * Z]
    const int32_t* row_exec;          // row (app, z) at Z * out_off[app] + z * count[app]
    int32_t n_apps, n_zones;
    int32_t executors_reserved;       // 1: tightly-pack (every executor is in `reserved`), 0: minimal-fragmentation (driver only)
};

__device__ __forceinline__ long long cpu_value(long long milli) {       // Quantity.Value(): whole cores, inexact -> away from zero
    const long long q = milli / 1000, rem = milli % 1000;
    return rem > 0 ? q + 1 : (rem < 0 ? q - 1 : q);
}
__device__ __forceinline__ long long normalize_resource(long long v) { return v == 0 ? 1 : v; }       // efficiency.go:104-109

// computePackingEfficiency (efficiency.go:79-102) for a node with SchedulableResources (sc, sm, sg), AvailableResources
// (ac, am, ag) and `r*` reserved on it; returns max(GPU, max(CPU, Memory)) and the three components
__device__ __forceinline__ double node_efficiency_values(long long sc, long long sm, long long sg, long long ac, long long am, long long ag,
                                                         long long rc, long long rm, long long rg,
                                                         double& cpu, double& mem, double& gpu, bool& has_gpu) {
    const long long uc = sc - ac + rc, um = sm - am + rm, ug = sg - ag + rg;
    has_gpu = sg != 0;
    gpu = has_gpu ? (double)ug / (double)normalize_resource(sg) : 0.0;
    cpu = (double)cpu_value(uc) / (double)normalize_resource(cpu_value(sc));
    mem = (double)um / (double)normalize_resource(sm);
    return fmax(gpu, fmax(cpu, mem));
}

// availability in node-table order (an immutable copy: independent batches)
struct NodeTableAvail {
    const long long* a[3];
    __device__ __forceinline__ void load(int32_t n, long long& c, long long& m, long long& g) const { c = a[0][n]; m = a[1][n]; g = a[2][n]; }
};

// ComputeAvgPackingEfficiency (efficiency.go:111-156) over [driver] + ExecutorNodes of ONE packing result, walked
// sequentially by the calling thread in the reference's order (duplicates kept) so the float64 sums round like the Go loop.
// AV: where the current AvailableResources of a node come from.
template <class AV>
__device__ __forceinline__ void zone_row_average(const AV& av, const long long* const* sched, int32_t d, const int32_t* ex, int32_t k,
                                                 long long dc, long long dm, long long dg, long long ec, long long em, long long eg,
                                                 bool executors_reserved, double* a4) {
    auto eff = [&](int32_t n, long long rc, long long rm, long long rg, double& c, double& m, double& g, bool& hg) {
        long long ac, am, ag;
        av.load(n, ac, am, ag);
        return node_efficiency_values(sched[0][n], sched[1][n], sched[2] ? sched[2][n] : 0, ac, am, ag, rc, rm, rg, c, m, g, hg);
    };
    // executors on the driver's node (reserved[driver] = driver + its executors, binpack.go:72-75 + pack_tightly.go:50)
    long long on_driver = 0;
    if (executors_reserved) for (int32_t t = 0; t < k; ++t) on_driver += ex[t] == d ? 1 : 0;
    double cpuSum = 0.0, memSum = 0.0, gpuSum = 0.0, maxSum = 0.0, c, m, g;
    int nodesWithGPU = 0;
    bool hg;
    double mx = eff(d, dc + on_driver * ec, dm + on_driver * em, dg + on_driver * eg, c, m, g, hg);
    cpuSum += c; memSum += m; if (hg) { gpuSum += g; nodesWithGPU++; } maxSum += mx;
    int32_t t = 0;
    while (t < k) {
        const int32_t n = ex[t];
        // executors of this placement on node n (ExecutorNodes may list a node in several runs: count them all)
        long long cnt = 0;
        if (executors_reserved) for (int32_t u = 0; u < k; ++u) cnt += ex[u] == n ? 1 : 0;
        const long long isd = n == d ? 1 : 0;
        mx = eff(n, isd * dc + cnt * ec, isd * dm + cnt * em, isd * dg + cnt * eg, c, m, g, hg);
        // every entry of the run adds the same efficiency again (ComputeAvgPackingEfficiency loops over entries)
        int32_t run = 1;
        while (t + run < k && ex[t + run] == n) ++run;
        for (int32_t u = 0; u < run; ++u) { cpuSum += c; memSum += m; if (hg) { gpuSum += g; nodesWithGPU++; } maxSum += mx; }
        t += run;
    }
    const double length = fmax((double)(k + 1), 1.0);
    a4[0] = cpuSum / length; a4[1] = memSum / length;
    a4[2] = nodesWithGPU == 0 ? 1.0 : gpuSum / (double)nodesWithGPU;
    a4[3] = maxSum / length;
}

__global__ void __launch_bounds__(256) gp_zone_choose(ZoneChooseIn in, int32_t* __restrict__ zone_out, int32_t* __restrict__ driver_out,
                                                      int32_t* __restrict__ exec_out, double* __restrict__ avg_out /* [n_apps][4] or NULL */) {
    const int lane = threadIdx.x & 31;
    const int32_t app = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (app >= in.n_apps) return;
    const int32_t Z = in.n_zones;
    const int32_t k = in.count[app] > 0 ? in.count[app] : 0;
    const int64_t base = (int64_t)Z * in.out_off[app];
    const long long dc = in.drv[0][app], dm = in.drv[1][app], dg = in.drv[2] ? in.drv[2][app] : 0;
    const long long ec = in.exe[0][app], em = in.exe[1][app], eg = in.exe[2] ? in.exe[2][app] : 0;
    NodeTableAvail av;
    av.a[0] = in.avail[0]; av.a[1] = in.avail[1]; av.a[2] = in.avail[2];
    double best = 0.0;                 // WorstAvgPackingEfficiency().Max
    int32_t best_z = -1;
    double best4[4] = {0.0, 0.0, 0.0, 0.0};
    for (int32_t z0 = 0; z0 < Z; z0 += 32) {
        const int32_t z = z0 + lane;
        double avg_max = 0.0, a4[4] = {0.0, 0.0, 0.0, 0.0};
        bool fits = false;
        if (z < Z) {
            const int32_t d = in.row_driver[(int64_t)app * Z + z];
            if (d >= 0) {
                fits = true;
                zone_row_average(av, in.sched, d, in.row_exec + base + (int64_t)z * k, k, dc, dm, dg, ec, em, eg, in.executors_reserved != 0, a4);
                avg_max = a4[3];
            }
        }
        // arg-max in zone order with strict '<' against the running best (single_az.go:91-94)
        for (int32_t t = 0; t < 32 && z0 + t < Z; ++t) {
            const bool f = __shfl_sync(0xffffffffu, fits ? 1 : 0, t) != 0;
            const double v = __shfl_sync(0xffffffffu, avg_max, t);
            const double v0 = __shfl_sync(0xffffffffu, a4[0], t), v1 = __shfl_sync(0xffffffffu, a4[1], t), v2 = __shfl_sync(0xffffffffu, a4[2], t);
            if (f && best < v) { best = v; best_z = z0 + t; best4[0] = v0; best4[1] = v1; best4[2] = v2; best4[3] = v; }
        }
    }
    if (lane == 0) {
        zone_out[app] = best_z;
        driver_out[app] = best_z >= 0 ? in.row_driver[(int64_t)app * Z + best_z] : -1;
        if (avg_out) { avg_out[4 * (int64_t)app + 0] = best4[0]; avg_out[4 * (int64_t)app + 1] = best4[1]; avg_out[4 * (int64_t)app + 2] = best4[2]; avg_out[4 * (int64_t)app + 3] = best4[3]; }
    }
    if (best_z >= 0) {
        const int32_t* ex = in.row_exec + base + (int64_t)best_z * k;
        int32_t* out = exec_out + in.out_off[app];
        for (int32_t t = lane; t < k; t += 32) out[t] = ex[t];
    }
}

// rows (app, z): the application's tuple with group = z, executor slice at Z * off[app] + z * count[app]
struct SixCols { const int64_t* p[6]; };      // drv cpu, drv mem, drv gpu, exe cpu, exe mem, exe gpu (NULL = 0)
__global__ void gp_zone_expand(int32_t n_apps, int32_t Z, SixCols src,
                               const int32_t* __restrict__ count, const int64_t* __restrict__ off,
                               int64_t* __restrict__ dst /* [6][n_apps * Z] */, int32_t* __restrict__ rcount, int32_t* __restrict__ rgroup,
                               int64_t* __restrict__ roff /* [n_apps * Z + 1] */) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t R = (int64_t)n_apps * Z;
    if (r > R) return;
    if (r == R) { roff[R] = (int64_t)Z * off[n_apps]; return; }
    const int32_t app = (int32_t)(r / Z), z = (int32_t)(r % Z);
#pragma unroll
    for (int c = 0; c < 6; ++c) dst[(int64_t)c * R + r] = src.p[c] ? src.p[c][app] : 0;
    const int32_t k = count[app];
    rcount[r] = k;
    rgroup[r] = z;
    roff[r] = (int64_t)Z * off[app] + (int64_t)z * (k > 0 ? k : 0);
}

// ---- reservation table of a batch of placements + the device-resident snapshot kept current ----
// newResourceReservation (internal/extender/resourcereservations.go:491-528): reservations["driver"] = {driver node, driver
// resources}, reservations["executor-<i>"] = {ExecutorNodes[i-1], executor resources} (executorReservationName, :530-533).
// One warp per placed application writes its 1 + k rows (slot 0 = "driver", slot i = "executor-i") and, when asked to,
// subtracts every pod from the availability the device keeps -- what UsageForNodes (LIB/resources/resources.go:31-43) will
// add up from these very reservations on the next Predicate.  Integer adds commute: atomics are exact.
struct ReserveIn {
    SixCols cols;                     // int64 device columns of the applications
    const int32_t* count;
    const int64_t* off;               // ExecutorNodes offsets
    const int64_t* row_off;           // first row of each application (placed ones only advance it)
    const int32_t* driver;
    const int32_t* exec;
    int32_t n_apps;
    int32_t subtract;
};
struct ReserveOut { int32_t* app; int32_t* slot; int32_t* node; long long* cpu; long long* mem; long long* gpu; };

__device__ __forceinline__ void charge_node(int32_t node, long long c, long long m, long long g, const int32_t* __restrict__ node_slot,
                                            longlong2* pair, long long* sgpu, long long* ncpu, long long* nmem, long long* ngpu) {
    const int32_t sl = node_slot[node];
    if (sl >= 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&pair[sl].x), (unsigned long long)(-c));
        atomicAdd(reinterpret_cast<unsigned long long*>(&pair[sl].y), (unsigned long long)(-m));
        if (g != 0) atomicAdd(reinterpret_cast<unsigned long long*>(sgpu + sl), (unsigned long long)(-g));
    } else {                          // a node outside every order lives in the node-table copy only
        atomicAdd(reinterpret_cast<unsigned long long*>(ncpu + node), (unsigned long long)(-c));
        atomicAdd(reinterpret_cast<unsigned long long*>(nmem + node), (unsigned long long)(-m));
        if (g != 0) atomicAdd(reinterpret_cast<unsigned long long*>(ngpu + node), (unsigned long long)(-g));
    }
}

__global__ void __launch_bounds__(256) gp_reserve_rows(ReserveIn in, ReserveOut out, const int32_t* __restrict__ node_slot, longlong2* pair,
                                                       long long* sgpu, long long* ncpu, long long* nmem, long long* ngpu) {
    const int lane = threadIdx.x & 31;
    const int32_t i = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (i >= in.n_apps) return;
    const int32_t d = in.driver[i];
    if (d < 0) return;
    const int32_t k = in.count[i] > 0 ? in.count[i] : 0;
    const int64_t r0 = in.row_off[i];
    const long long dc = in.cols.p[0][i], dm = in.cols.p[1][i], dg = in.cols.p[2] ? in.cols.p[2][i] : 0;
    const long long ec = in.cols.p[3][i], em = in.cols.p[4][i], eg = in.cols.p[5] ? in.cols.p[5][i] : 0;
    if (lane == 0) {
        out.app[r0] = i; out.slot[r0] = 0; out.node[r0] = d; out.cpu[r0] = dc; out.mem[r0] = dm; out.gpu[r0] = dg;
        if (in.subtract) charge_node(d, dc, dm, dg, node_slot, pair, sgpu, ncpu, nmem, ngpu);
    }
    const int32_t* ex = in.exec + in.off[i];
    for (int32_t t = lane; t < k; t += 32) {
        const int64_t r = r0 + 1 + t;
        const int32_t n = ex[t];
        out.app[r] = i; out.slot[r] = t + 1; out.node[r] = n; out.cpu[r] = ec; out.mem[r] = em; out.gpu[r] = eg;
        if (in.subtract) charge_node(n, ec, em, eg, node_slot, pair, sgpu, ncpu, nmem, ngpu);
    }
}

// availability[node] -= sign * (cpu, mem, gpu) for a list of (node, resources) rows: reservations that appeared (sign +1)
// or went away (sign -1) since the snapshot was laid out.  Keeps SnapMeta::max_avail an upper bound.
__global__ void gp_usage_delta(int64_t n_rows, const int32_t* __restrict__ node, const long long* __restrict__ cpu, const long long* __restrict__ mem,
                               const long long* __restrict__ gpu, int sign, int32_t n_nodes, const int32_t* __restrict__ node_slot, longlong2* pair,
                               long long* sgpu, long long* ncpu, long long* nmem, long long* ngpu) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t n = node[r];
    if (n < 0 || n >= n_nodes) return;           // a reservation on a node that left: ignored (resources.go:67-75)
    charge_node(n, sign * cpu[r], sign * mem[r], gpu ? sign * gpu[r] : 0, node_slot, pair, sgpu, ncpu, nmem, ngpu);
}
// after a delta that may have RAISED availabilities: refresh the per-dimension maxima and the negative-gpu flag
__global__ void gp_refresh_meta(int32_t n_slots, const longlong2* __restrict__ pair, const long long* __restrict__ sgpu, const int32_t* __restrict__ slot_node,
                                int* flags, long long* max_avail) {
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots || slot_node[s] < 0) return;
    const longlong2 v = pair[s];
    const long long g = sgpu[s];
    if (v.x > max_avail[0]) atomicMax(max_avail + 0, v.x);
    if (v.y > max_avail[1]) atomicMax(max_avail + 1, v.y);
    if (g > max_avail[2]) atomicMax(max_avail + 2, g);
    if (g < 0) atomicOr(flags, 1);
}

}  // namespace gp

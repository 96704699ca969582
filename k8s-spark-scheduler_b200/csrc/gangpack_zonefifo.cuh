// gangpack_zonefifo.cuh -- the FIFO loop with a single-AZ packer, in ONE launch.
//
// Reference: fitEarlierDrivers (internal/extender/resource.go:224-262) calls binpacker.BinpackFunc once per queued driver;
// with `binpack: single-az-tightly-pack` / `single-az-minimal-fragmentation` that function packs the application in EVERY
// zone and keeps the result with the best average packing efficiency (LIB/binpack/single_az.go:23-55, 75-97), and the
// usage of the winner is subtracted before the next driver is tried (sparkResourceUsage, EXT/sparkpods.go:139-146 +
// SubtractUsageIfExists, LIB/resources/resources.go:129-135).  The zone choice of driver i therefore feeds driver i+1:
// the queue is one sequential chain over the whole cluster (zones are NOT independent queues).
//
// One persistent CTA walks the queue.  Per application:
//   1. warp w packs it in zones w, w + W, ... (snapshot instance group = zone) with the warp-per-application scan of the
//      independent kernels -- here with coherent loads, the availabilities change inside the launch -- into a scratch row
//      per zone; lane 0 of the warp then walks its row for the float64 averages in the reference's operation order;
//   2. __syncthreads; every thread takes the arg-max in zone order (strict '<' against a best that starts at 0.0);
//   3. the winner's warp copies the row out and charges the slots: FIFO_MODE 1 = the reference's map assignment (every
//      DISTINCT executor node once, the driver only when its node hosts no executor), 2 = every pod;
//   4. __syncthreads, next application.  A driver that fits nowhere blocks the queue unless it is young
//      (resource.go:244-253); the applications behind it report -2 like gp_pack_fifo_cta.
#pragma once

#include "gangpack_kernels.cuh"
#include "gangpack_minfrag.cuh"
#include "gangpack_zones.cuh"

namespace gp {

constexpr int kZoneFifoWarps = 16;
constexpr int kZoneFifoThreads = kZoneFifoWarps * 32;
constexpr int kZoneFifoMaxZones = 64;

struct ZoneFifoIn {
    const long long* sched[3];        // SchedulableResources, node-table order (gpu may be NULL = 0)
    const int32_t* node_slot;         // node -> global slot
    int32_t* row_exec;                // [Z][row_pitch] ExecutorNodes of the application in every zone
    int2* row_list;                   // [Z][row_pitch] consumed-node list of minimal-fragmentation, NULL for tightly-pack
    int64_t row_pitch;                // >= max exe_count of the batch
    int32_t n_apps, n_zones;
};

// AvailableResources of a node as the loop currently holds them (the slots are charged inside the launch)
struct SlotAvail {
    const longlong2* pair;
    const int64_t* gpu;
    const int32_t* node_slot;
    __device__ __forceinline__ void load(int32_t n, long long& c, long long& m, long long& g) const {
        const int32_t sl = node_slot[n];
        const longlong2 v = load_pair<true>(pair + sl);
        c = v.x; m = v.y; g = load_gpu<true>(gpu + sl);
    }
};

__device__ __forceinline__ void zone_charge(const Snapshot& s, int32_t slot, long long mult, long long c, long long m, long long g) {
    longlong2 v = load_pair<true>(s.pair + slot);
    v.x -= mult * c; v.y -= mult * m;
    s.pair[slot] = v;
    if (g != 0) s.gpu[slot] = load_gpu<true>(s.gpu + slot) - mult * g;
}

// ALGO: 0 tightly-pack, 2 minimal-fragmentation.  FIFO_MODE: 1 reference accounting, 2 exact.
template <int ALGO, int FIFO_MODE>
__global__ void __launch_bounds__(kZoneFifoThreads, 1) gp_pack_fifo_zones_cta(Snapshot s, const PrepApp* __restrict__ prep, ZoneFifoIn in,
                                                                              int32_t* __restrict__ zone_out, int32_t* __restrict__ driver_out,
                                                                              int32_t* __restrict__ exec_out, double* __restrict__ avg_out,
                                                                              unsigned long long* __restrict__ stats) {
    __shared__ uint16_t caches[kZoneFifoWarps][kCapCache];
    __shared__ double z_avg[kZoneFifoMaxZones][4];
    __shared__ int32_t z_drv[kZoneFifoMaxZones];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t Z = in.n_zones;
    const GroupDesc g0 = s.groups[0];
    SlotAvail av;
    av.pair = s.pair; av.gpu = s.gpu; av.node_slot = in.node_slot;
    WarpStats st{0, 0};
    const int snap_flags = s.meta->flags;     // a placement only lands where it fits: no availability turns negative inside the loop
    bool blocked = false;
    for (int32_t app = 0; app < in.n_apps; ++app) {
        const PrepApp* pa = prep + app;
        const uint32_t fl = pa->flags;
        if (blocked || (fl & kAppInvalid)) {
            if (threadIdx.x == 0) {
                zone_out[app] = -1;
                driver_out[app] = blocked ? -2 : -1;                 // never evaluated (resource.go:252)
                if (avg_out) for (int t = 0; t < 4; ++t) avg_out[4 * (int64_t)app + t] = 0.0;
            }
            continue;
        }
        const int32_t k = pa->count;
        // ---- 1. every zone: pack + averages ---------------------------------------------------------------------------
        for (int32_t z = warp; z < Z; z += kZoneFifoWarps) {
            const int64_t row = (int64_t)z * in.row_pitch;
            int32_t d;
            if (ALGO == 0)
                d = pack_app_impl<0, false, false, false, int32_t, true>(s, pa, in.row_exec, nullptr, caches[warp], st, lane, snap_flags, g0, z, row);
            else
                d = pack_app_minfrag<false, true>(s, pa, in.row_exec, in.row_list, st, lane, snap_flags, z, row);
            __syncwarp();
            if (lane == 0) {
                z_drv[z] = d;
                if (d >= 0)
                    zone_row_average(av, in.sched, d, in.row_exec + row, k, pa->drv[0], pa->drv[1], pa->drv[2],
                                     pa->div[0].e, pa->div[1].e, pa->div[2].e, ALGO == 0, z_avg[z]);
            }
        }
        __syncthreads();
        // ---- 2. chooseBestResult (single_az.go:75-97), identically in every thread ------------------------------------
        double best = 0.0;
        int32_t best_z = -1;
        for (int32_t z = 0; z < Z; ++z)
            if (z_drv[z] >= 0 && best < z_avg[z][3]) { best = z_avg[z][3]; best_z = z; }
        // ---- 3. the winner's warp emits and charges ---------------------------------------------------------------------
        if (best_z >= 0 && warp == best_z % kZoneFifoWarps) {
            const int32_t* ex = in.row_exec + (int64_t)best_z * in.row_pitch;
            int32_t* out = exec_out + pa->out_off;
            const int32_t d = z_drv[best_z];
            const long long ec = pa->div[0].e, em = pa->div[1].e, eg = pa->div[2].e;
            bool hosts = false;                                      // the driver's node hosts an executor
            for (int32_t t0 = 0; t0 < k; t0 += 32) {
                const int32_t t = t0 + lane;
                if (t < k) {
                    const int32_t n = ex[t];
                    out[t] = n;
                    hosts = hosts || n == d;
                    // one writer per node: the first entry that names it, with the number of entries that do
                    bool first = t == 0 || ex[t - 1] != n;
                    for (int32_t u = 0; first && u + 1 < t; ++u) first = ex[u] != n;
                    if (first) {
                        long long cnt = 1;
                        if (FIFO_MODE == 2) { cnt = 0; for (int32_t u = t; u < k; ++u) cnt += ex[u] == n ? 1 : 0; }
                        zone_charge(s, in.node_slot[n], cnt, ec, em, eg);
                    }
                }
            }
            hosts = __any_sync(kFull, hosts);
            __syncwarp();
            if (lane == 0) {
                if (FIFO_MODE == 2 || !hosts) zone_charge(s, in.node_slot[d], 1, pa->drv[0], pa->drv[1], pa->drv[2]);
                zone_out[app] = best_z;
                driver_out[app] = d;
                if (avg_out) for (int t = 0; t < 4; ++t) avg_out[4 * (int64_t)app + t] = z_avg[best_z][t];
            }
        } else if (best_z < 0 && threadIdx.x == 0) {
            zone_out[app] = -1;
            driver_out[app] = -1;
            if (avg_out) for (int t = 0; t < 4; ++t) avg_out[4 * (int64_t)app + t] = 0.0;
        }
        if (best_z < 0 && !(fl & kAppSkipIfNoFit)) blocked = true;    // resource.go:244-253
        __syncthreads();
    }
    if (lane == 0 && stats) { atomicAdd(stats + 0, st.nodes); atomicAdd(stats + 1, st.drivers); }
}

}  // namespace gp

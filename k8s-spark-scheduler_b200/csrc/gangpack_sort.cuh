// gangpack_sort.cuh -- SURVEY §8f row f1: NodeSorter.PotentialNodes (internal/sort/nodesorting.go:41-122,
// 161-200) on the device -- the step immediately before the packing hot path.
//
//   node priority order = ascending (AZ priority, available memory, available CPU, node name)      :83-93
//   AZ priority         = rank of the zone by (sum of available memory, sum of available CPU)       :102-115
//   driver candidates   = that order restricted to kube-scheduler's NodeNames                        :52-54
//   executor candidates = that order restricted to schedulable && ready nodes                        :55-57
//   optional stable re-sort of either list by the rank of a configured label value                  :61-62, 161-200
//
// The key is a strict total order (node names are unique), so the position of a node is simply the number
// of nodes with a smaller key: an enumeration sort.  N^2 comparisons, but every one is three integer
// compares on data staged in shared memory, every node is independent, there is no multi-pass data
// movement and the result is deterministic -- 10^8 comparisons for 10 000 nodes are a few tens of
// microseconds on 148 SMs.  (A radix sort is the better tool from ~10^5 nodes on.)
//
// Where the reference's comparator leaves the order undefined (equal (memory, cpu) but different gpu:
// scheduleContextLessThan is not a strict weak order there; equal zone totals) this
// implementation orders by node name / by zone id.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gp {

struct SortKey {          // 24 bytes
    long long mem;
    long long cpu;
    int az;               // AZ priority
    int name_rank;        // rank of the node name (unique)
};

__device__ __forceinline__ bool key_less(const SortKey& a, const SortKey& b) {
    if (a.az != b.az) return a.az < b.az;
    if (a.mem != b.mem) return a.mem < b.mem;
    if (a.cpu != b.cpu) return a.cpu < b.cpu;
    return a.name_rank < b.name_rank;
}

// per-zone totals of available memory / cpu (getAvailableResourcesByAZ, :124-134)
__global__ void gp_zone_totals(int32_t n, const long long* __restrict__ cpu, const long long* __restrict__ mem,
                               const int32_t* __restrict__ zone, unsigned long long* __restrict__ tot /* [2*Z]: mem, cpu */) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t z = zone[i];
    atomicAdd(tot + 2 * z + 0, (unsigned long long)mem[i]);     // two's complement: signed sums wrap correctly
    atomicAdd(tot + 2 * z + 1, (unsigned long long)cpu[i]);
}

// AZ priority = number of zones with smaller (mem, cpu, id)   (sort.Slice over zones, :102-104; ties by zone id)
__global__ void gp_zone_priority(int32_t n_zones, const unsigned long long* __restrict__ tot, int32_t* __restrict__ prio) {
    int32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= n_zones) return;
    const long long m = (long long)tot[2 * z], c = (long long)tot[2 * z + 1];
    int32_t p = 0;
    for (int32_t o = 0; o < n_zones; ++o) {
        const long long om = (long long)tot[2 * o], oc = (long long)tot[2 * o + 1];
        const bool less = (om != m) ? (om < m) : ((oc != c) ? (oc < c) : (o < z));
        p += less ? 1 : 0;
    }
    prio[z] = p;
}

__global__ void gp_make_keys(int32_t n, const long long* __restrict__ cpu, const long long* __restrict__ mem,
                             const int32_t* __restrict__ zone, const int32_t* __restrict__ zone_prio,
                             const int32_t* __restrict__ name_rank, SortKey* __restrict__ keys, int32_t* __restrict__ pos) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SortKey k;
    k.mem = mem[i]; k.cpu = cpu[i]; k.az = zone_prio[zone[i]]; k.name_rank = name_rank[i];
    keys[i] = k;
    pos[i] = 0;
}

// pos[i] += #{ j in this CTA's j-tile : key_j < key_i }     grid = (ceil(n/256), ceil(n/kTileJ))
constexpr int kTileJ = 2048;
__global__ void __launch_bounds__(256) gp_rank_nodes(int32_t n, const SortKey* __restrict__ keys, int32_t* __restrict__ pos) {
    __shared__ SortKey tile[kTileJ];
    const int32_t j0 = blockIdx.y * kTileJ;
    const int32_t jn = min(kTileJ, n - j0);
    for (int32_t t = threadIdx.x; t < jn; t += blockDim.x) tile[t] = keys[j0 + t];
    __syncthreads();
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SortKey me = keys[i];
    int32_t cnt = 0;
#pragma unroll 4
    for (int32_t t = 0; t < jn; ++t) cnt += key_less(tile[t], me) ? 1 : 0;     // broadcast reads: no bank conflicts
    if (cnt) atomicAdd(pos + i, cnt);
}

__global__ void gp_scatter_order(int32_t n, const int32_t* __restrict__ pos, int32_t* __restrict__ order) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[pos[i]] = i;
}

// Order-preserving compaction of the priority order into the driver and executor candidate lists (one CTA).
// counts[0] = #driver candidates, counts[1] = #executor candidates.
__global__ void __launch_bounds__(1024) gp_split_candidates(int32_t n, const int32_t* __restrict__ order,
                                                            const uint8_t* __restrict__ is_candidate,
                                                            const uint8_t* __restrict__ unschedulable, const uint8_t* __restrict__ ready,
                                                            int32_t* __restrict__ drv, int32_t* __restrict__ exe, int32_t* __restrict__ counts) {
    __shared__ int32_t wsum[2][32];
    __shared__ int32_t base[2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) { base[0] = 0; base[1] = 0; }
    __syncthreads();
    for (int32_t p0 = 0; p0 < n; p0 += blockDim.x) {
        const int32_t p = p0 + tid;
        int32_t node = -1;
        bool fd = false, fe = false;
        if (p < n) {
            node = order[p];
            fd = is_candidate ? is_candidate[node] != 0 : true;
            fe = !(unschedulable && unschedulable[node]) && (ready ? ready[node] != 0 : true);
        }
        const unsigned bd = __ballot_sync(0xffffffffu, fd), be = __ballot_sync(0xffffffffu, fe);
        if (lane == 0) { wsum[0][w] = __popc(bd); wsum[1][w] = __popc(be); }
        __syncthreads();
        int32_t od = base[0], oe = base[1];
        for (int k = 0; k < w; ++k) { od += wsum[0][k]; oe += wsum[1][k]; }
        const unsigned below = (1u << lane) - 1u;
        if (fd) drv[od + __popc(bd & below)] = node;
        if (fe) exe[oe + __popc(be & below)] = node;
        __syncthreads();
        if (tid == 0) {
            int32_t td = 0, te = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { td += wsum[0][k]; te += wsum[1][k]; }
            base[0] += td; base[1] += te;
        }
        __syncthreads();
    }
    if (tid == 0) { counts[0] = base[0]; counts[1] = base[1]; }
}

// Stable re-sort of a candidate list by label rank (createLabelLessThanFunction :161-180: unknown rank sorts
// last, ties keep their order): out[#{j : (rank_j, j) < (rank_i, i)}] = list[i].
__global__ void __launch_bounds__(256) gp_label_sort(const int32_t* __restrict__ count_ptr, const int32_t* __restrict__ list,
                                                     const int32_t* __restrict__ label_rank, int32_t* __restrict__ out) {
    const int32_t m = *count_ptr;
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t node = list[i];
    const int32_t lr = label_rank[node];
    const uint32_t ri = lr < 0 ? 0x7fffffffu : (uint32_t)lr;
    int32_t cnt = 0;
    for (int32_t j = 0; j < m; ++j) {
        const int32_t lj = label_rank[list[j]];
        const uint32_t rj = lj < 0 ? 0x7fffffffu : (uint32_t)lj;
        cnt += (rj < ri || (rj == ri && j < i)) ? 1 : 0;
    }
    out[cnt] = node;
}

// ---- the availability snapshot from reservations -------------------------------------------
// usage[node] = sum of the resources of every (hard or soft) reservation on that node: UsageForNodes
// (LIB/resources/resources.go:31-43) + UsedSoftReservationResources (internal/cache/softreservations.go:155-170),
// summed as in GetReservedResources (EXT/resourcereservations.go:258-263).  A segmented reduction over
// (reservation -> node) pairs: thread per reservation, 64-bit atomics (integer adds commute: exact).
__global__ void gp_usage_scatter(long long n_res, const int32_t* __restrict__ res_node, const long long* __restrict__ res_cpu,
                                 const long long* __restrict__ res_mem, const long long* __restrict__ res_gpu, int32_t n_nodes,
                                 unsigned long long* __restrict__ usage /* [3][n_nodes] */) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= n_res) return;
    const int32_t node = res_node[r];
    if (node < 0 || node >= n_nodes) return;            // reservation on a node that is not in the list: ignored (:67-75)
    atomicAdd(usage + node, (unsigned long long)res_cpu[r]);
    atomicAdd(usage + n_nodes + node, (unsigned long long)res_mem[r]);
    if (res_gpu) atomicAdd(usage + 2 * (size_t)n_nodes + node, (unsigned long long)res_gpu[r]);
}

// NodeSchedulingMetadataForNodes (resources.go:61-100): available = allocatable - (usage + overhead),
// schedulable = allocatable - overhead.
__global__ void gp_availability(int32_t n, const long long* __restrict__ alloc_cpu, const long long* __restrict__ alloc_mem,
                                const long long* __restrict__ alloc_gpu, const long long* __restrict__ over_cpu,
                                const long long* __restrict__ over_mem, const long long* __restrict__ over_gpu,
                                const unsigned long long* __restrict__ usage, long long* __restrict__ avail /* [3][n] */,
                                long long* __restrict__ sched /* [3][n] */) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a[3] = {alloc_cpu[i], alloc_mem[i], alloc_gpu ? alloc_gpu[i] : 0};
    const long long o[3] = {over_cpu ? over_cpu[i] : 0, over_mem ? over_mem[i] : 0, over_gpu ? over_gpu[i] : 0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const long long u = (long long)usage[(size_t)d * n + i];
        avail[(size_t)d * n + i] = a[d] - (u + o[d]);
        sched[(size_t)d * n + i] = a[d] - o[d];
    }
}

}  // namespace gp

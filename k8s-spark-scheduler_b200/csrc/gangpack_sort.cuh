// gangpack_sort.cuh -- SURVEY §8f row f1: NodeSorter.PotentialNodes (internal/sort/nodesorting.go:41-122,
// 161-200) on the device -- the step immediately before the packing hot path.
//
//   node priority order = ascending (AZ priority, available memory, available CPU, node name)      :83-93
//   AZ priority         = rank of the zone by (sum of available memory, sum of available CPU)       :102-115
//   driver candidates   = that order restricted to kube-scheduler's NodeNames                        :52-54
//   executor candidates = that order restricted to schedulable && ready nodes                        :55-57
//   optional stable re-sort of either list by the rank of a configured label value                  :61-62, 161-200
//
// The key is a strict total order (node names are unique), so the position of a node is the number of nodes with a
// smaller key.  Two kernels: (1) every CTA sorts one tile of 4 096 keys in shared memory (bitonic network, 96 KB);
// (2) every node binary-searches its key in every sorted tile and adds up the lower bounds -- N x (N / 4096) x 12
// dependent steps instead of the N^2 comparisons of an enumeration sort (10 000 nodes: 3 tiles; 50 000: 13; 500 000:
// 123), no atomics, no multi-pass data movement, deterministic.  The same primitive re-sorts a candidate list stably
// by label rank (key = (rank, position)).
//
// Where the reference's comparator leaves the order undefined (equal (memory, cpu) but different gpu:
// scheduleContextLessThan is not a strict weak order there; equal zone totals) this
// implementation orders by node name / by zone id -- and REPORTS it (gp_sort_input.undefined_ties) so that the shim can
// route such a Predicate to the Go sorter if it wants the reference's (unspecified) choice.
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

// zone == NULL: one zone (priority 0); name_rank == NULL: index order
__global__ void gp_make_keys(int32_t n, const long long* __restrict__ cpu, const long long* __restrict__ mem,
                             const int32_t* __restrict__ zone, const int32_t* __restrict__ zone_prio,
                             const int32_t* __restrict__ name_rank, SortKey* __restrict__ keys) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SortKey k;
    k.mem = mem[i]; k.cpu = cpu[i]; k.az = zone ? zone_prio[zone[i]] : 0; k.name_rank = name_rank ? name_rank[i] : i;
    keys[i] = k;
}

// ---- sort primitive: tiles sorted in shared memory, ranks by binary search across tiles ---------------------------
constexpr int kSortTile = 4096;
constexpr int kSortThreads = 1024;

struct LabelKey {          // stable re-sort by label rank: (rank, position in the list) is unique
    uint32_t rank;
    int32_t pos;
};
__device__ __forceinline__ bool key_less(const LabelKey& a, const LabelKey& b) { return a.rank != b.rank ? a.rank < b.rank : a.pos < b.pos; }
__device__ __forceinline__ SortKey key_max(const SortKey*) { SortKey k; k.mem = 0x7fffffffffffffffLL; k.cpu = 0x7fffffffffffffffLL; k.az = 0x7fffffff; k.name_rank = 0x7fffffff; return k; }
__device__ __forceinline__ LabelKey key_max(const LabelKey*) { LabelKey k; k.rank = 0xffffffffu; k.pos = 0x7fffffff; return k; }

// (1) sorted[tile] = the tile's keys in ascending order.  count_ptr: the number of keys lives on the device (candidate
// lists) or is `n` when NULL.  Dynamic shared memory: kSortTile * sizeof(K).
template <class K>
__global__ void __launch_bounds__(kSortThreads) gp_sort_tiles(int32_t n, const int32_t* __restrict__ count_ptr, const K* __restrict__ keys,
                                                              K* __restrict__ sorted) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    K* t = reinterpret_cast<K*>(sort_smem);
    if (count_ptr) n = *count_ptr;
    const int32_t base = blockIdx.x * kSortTile;
    if (base >= n) return;
    const int32_t m = min(kSortTile, n - base);
    for (int32_t i = threadIdx.x; i < kSortTile; i += blockDim.x) t[i] = i < m ? keys[base + i] : key_max((const K*)nullptr);
    __syncthreads();
    for (int32_t size = 2; size <= kSortTile; size <<= 1) {
        for (int32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (int32_t p = threadIdx.x; p < kSortTile / 2; p += blockDim.x) {
                const int32_t lo = 2 * p - (p & (stride - 1));          // index with bit `stride` clear
                const int32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const K a = t[lo], b = t[hi];
                if (key_less(b, a) == up) { t[lo] = b; t[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int32_t i = threadIdx.x; i < m; i += blockDim.x) sorted[base + i] = t[i];
}

// (2) pos[i] = #{ j : key_j < key_i } = sum over tiles of lower_bound(tile, key_i)
template <class K>
__global__ void __launch_bounds__(256) gp_rank_by_search(int32_t n, const int32_t* __restrict__ count_ptr, const K* __restrict__ keys,
                                                         const K* __restrict__ sorted, int32_t* __restrict__ pos) {
    if (count_ptr) n = *count_ptr;
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const K me = keys[i];
    int32_t cnt = 0;
    for (int32_t base = 0; base < n; base += kSortTile) {
        const K* t = sorted + base;
        int32_t lo = 0, hi = min(kSortTile, n - base);              // first index whose key is not < me
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (key_less(t[mid], me)) lo = mid + 1; else hi = mid;
        }
        cnt += lo;
    }
    pos[i] = cnt;
}

__global__ void gp_scatter_order(int32_t n, const int32_t* __restrict__ pos, int32_t* __restrict__ order) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[pos[i]] = i;
}

// Order-preserving compaction of the priority order into the driver and executor candidate lists (one CTA, every
// thread owns 8 consecutive positions per step).  counts[0] = #driver candidates, counts[1] = #executor candidates,
// counts[2] = #adjacent pairs of the order that the reference's comparator leaves undefined (same zone priority, memory
// and cpu but different gpu: SURVEY App. B6) -- 0 when no gpu column was given.
constexpr int kSplitItems = 8;
__global__ void __launch_bounds__(1024) gp_split_candidates(int32_t n, const int32_t* __restrict__ order,
                                                            const uint8_t* __restrict__ is_candidate,
                                                            const uint8_t* __restrict__ unschedulable, const uint8_t* __restrict__ ready,
                                                            const SortKey* __restrict__ keys, const long long* __restrict__ gpu,
                                                            int32_t* __restrict__ drv, int32_t* __restrict__ exe, int32_t* __restrict__ counts) {
    __shared__ int32_t wsum[2][32];
    __shared__ int32_t base[2];
    __shared__ int32_t ties;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) { base[0] = 0; base[1] = 0; ties = 0; }
    __syncthreads();
    int32_t my_ties = 0;
    for (int32_t p0 = 0; p0 < n; p0 += blockDim.x * kSplitItems) {
        const int32_t pb = p0 + tid * kSplitItems;
        int32_t node[kSplitItems];
        uint32_t fd = 0, fe = 0;
#pragma unroll
        for (int t = 0; t < kSplitItems; ++t) {
            const int32_t p = pb + t;
            node[t] = -1;
            if (p < n) {
                node[t] = order[p];
                if (is_candidate ? is_candidate[node[t]] != 0 : true) fd |= 1u << t;
                if (!(unschedulable && unschedulable[node[t]]) && (ready ? ready[node[t]] != 0 : true)) fe |= 1u << t;
                if (gpu && p > 0) {
                    const int32_t prev = t > 0 ? node[t - 1] : order[p - 1];
                    const SortKey a = keys[prev], b = keys[node[t]];
                    if (a.az == b.az && a.mem == b.mem && a.cpu == b.cpu && gpu[prev] != gpu[node[t]]) ++my_ties;
                }
            }
        }
        const int32_t cd = __popc(fd), ce = __popc(fe);
        // exclusive scan of (cd, ce) over the block
        int32_t id = cd, ie = ce;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t ud = __shfl_up_sync(0xffffffffu, id, d), ue = __shfl_up_sync(0xffffffffu, ie, d);
            if (lane >= d) { id += ud; ie += ue; }
        }
        if (lane == 31) { wsum[0][w] = id; wsum[1][w] = ie; }
        __syncthreads();
        int32_t od = base[0] + id - cd, oe = base[1] + ie - ce;
        for (int k = 0; k < w; ++k) { od += wsum[0][k]; oe += wsum[1][k]; }
#pragma unroll
        for (int t = 0; t < kSplitItems; ++t) {
            if (fd & (1u << t)) drv[od++] = node[t];
            if (fe & (1u << t)) exe[oe++] = node[t];
        }
        __syncthreads();
        if (tid == 0) {
            int32_t td = 0, te = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { td += wsum[0][k]; te += wsum[1][k]; }
            base[0] += td; base[1] += te;
        }
        __syncthreads();
    }
    if (my_ties) atomicAdd(&ties, my_ties);
    __syncthreads();
    if (tid == 0) { counts[0] = base[0]; counts[1] = base[1]; counts[2] = ties; }
}

// Stable re-sort of a candidate list by label rank (createLabelLessThanFunction :161-180: unknown rank sorts
// last, ties keep their order): keys (rank, position) through the sort primitive above, then out[pos_i] = list[i].
__global__ void gp_label_keys(const int32_t* __restrict__ count_ptr, const int32_t* __restrict__ list,
                              const int32_t* __restrict__ label_rank, LabelKey* __restrict__ keys) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_ptr) return;
    const int32_t lr = label_rank[list[i]];
    LabelKey k; k.rank = lr < 0 ? 0x7fffffffu : (uint32_t)lr; k.pos = i;
    keys[i] = k;
}
__global__ void gp_label_scatter(const int32_t* __restrict__ count_ptr, const int32_t* __restrict__ list,
                                 const int32_t* __restrict__ pos, int32_t* __restrict__ out) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *count_ptr) out[pos[i]] = list[i];
}

// ---- the availability snapshot from reservations -------------------------------------------
// usage[node] = sum of the resources of every (hard or soft) reservation on that node: UsageForNodes
// (LIB/resources/resources.go:31-43) + UsedSoftReservationResources (internal/cache/softreservations.go:155-170),
// summed as in GetReservedResources (EXT/resourcereservations.go:258-263).  A segmented reduction over
// (reservation -> node) pairs: thread per reservation, 64-bit atomics (integer adds commute: exact).
__global__ void gp_usage_scatter(long long n_res, const int32_t* __restrict__ res_node, const long long* __restrict__ res_cpu,
                                 const long long* __restrict__ res_mem, const long long* __restrict__ res_gpu, int32_t n_nodes,
                                 unsigned long long* __restrict__ usage /* [3][n_nodes] */, unsigned int* __restrict__ has_entry /* [n_nodes] or NULL */) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= n_res) return;
    const int32_t node = res_node[r];
    if (node < 0 || node >= n_nodes) return;            // reservation on a node that is not in the list: ignored (:67-75)
    if (has_entry) has_entry[node] = 1u;                // the usage map has an entry for this node (whatever its value)
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
                                long long* __restrict__ sched /* [3][n] */, const unsigned int* __restrict__ has_entry = nullptr,
                                long long* __restrict__ resched /* [3][n] or NULL */ = nullptr) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a[3] = {alloc_cpu[i], alloc_mem[i], alloc_gpu ? alloc_gpu[i] : 0};
    const long long o[3] = {over_cpu ? over_cpu[i] : 0, over_mem ? over_mem[i] : 0, over_gpu ? over_gpu[i] : 0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const long long u = (long long)usage[(size_t)d * n + i];
        avail[(size_t)d * n + i] = a[d] - (u + o[d]);
        sched[(size_t)d * n + i] = a[d] - o[d];
        // availableResources of rescheduleExecutor (EXT/resource.go:638-643): NodeSchedulingMetadataForNodes has already added
        // the overhead into the usage map's EXISTING entries in place when usage.Add(overhead) adds it again
        if (resched) resched[(size_t)d * n + i] = a[d] - (u + o[d] * (has_entry[i] ? 2 : 1));
    }
}

}  // namespace gp

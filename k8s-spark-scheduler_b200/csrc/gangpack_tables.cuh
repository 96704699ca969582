// gangpack_tables.cuh -- independent mode (GP_MODE_INDEPENDENT) for tightly-pack / distribute-evenly:
// per-executor-shape capacity tables + the fused (prep + pack) warp-per-application kernel.
//
// Reference semantics (all under /root/reference; LIB = vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg):
//   SparkBinPack driver loop        LIB/binpack/binpack.go:60-87
//   tightlyPackExecutors            LIB/binpack/pack_tightly.go:34-63
//   distributeExecutorsEvenly       LIB/binpack/distribute_evenly.go:34-73
//   node capacity                   LIB/capacity/capacity.go:36-75
//
// Why tables.  Every application of an independent batch is packed against the SAME snapshot
// (internal/extender/unschedulablepods.go:132-166; each Predicate's own pack, resource.go:321), and a node's executor
// capacity depends only on the executor request -- the "shape" (cpu, mem, gpu).  Real queues hold few shapes (the
// synthetic one: 12).  So instead of every application re-scanning the node order, each distinct shape is scanned ONCE:
//   K1 gp_classify_apps   thread per application: intern its shape in a small hash table (atomicCAS on a 64-bit
//                         fingerprint; the full tuple is verified again by the consumer, so a fingerprint collision costs
//                         a fallback, never a wrong answer) and derive the ExecutorNodes offsets when the caller gave none;
//   K2 gp_build_shape_tables  one CTA per (shape, instance group): the group's (cpu, mem) slots are staged tile by tile
//                         into shared memory by TMA bulk copies (cp.async.bulk + mbarrier, double-buffered), every
//                         thread evaluates 4 consecutive nodes, a block-wide exclusive scan turns capacities into the
//                         prefix table  S[i] = sum_{n<i} min(cap(n|0), CLAMP)   (tightly-pack) or
//                                       M[i] = #{n<i : cap(n|0) >= 1}            (distribute-evenly);
//   K3 gp_pack_tables     one warp per application, inputs read raw (no prepared record round-trips HBM): feasibility
//                         is  S[ne] - delta(d) >= k  per driver candidate d (delta = what the driver displaces on its
//                         own node: O(1) per candidate, the reference's loop binpack.go:67-85 verbatim), the first
//                         hosting node is found by a 32-ary search of the table, and ExecutorNodes is expanded from the
//                         prefix values of 32 nodes at a time (shuffle binary search -> coalesced stores).
// What the tables cannot answer exactly is handed to the scan path of gangpack_kernels.cuh INSIDE the same kernel
// (warp-uniform branch): more than kMaxShapes distinct shapes in a batch, executor counts above the table clamp,
// distribute-evenly placements that need more than one round.  GANGPACK_TABLES=0 forces that path for everything.
#pragma once

#include "gangpack_kernels.cuh"
#include "gangpack_fifo.cuh"     // TMA / mbarrier helpers

namespace gp {

constexpr int kShapeSlots = 1024;     // hash slots (power of two)
constexpr int kShapeProbes = 16;      // linear probes before giving up (-> scan path)
constexpr int kMaxShapes = 64;        // dense tables per batch
constexpr int kTabThreads = 1024;
constexpr int kTabPerThread = 4;
constexpr int kTabTile = kTabThreads * kTabPerThread;     // 4096 nodes per tile: 64 KB of (cpu, mem)

struct __align__(16) ShapeEntry {     // 96 bytes
    unsigned long long key;           // fingerprint, never 0; 0 = empty (claimed with atomicCAS)
    int32_t id;                       // dense table id in claim order; -1: more than kMaxShapes shapes in this batch
    uint32_t flags;                   // bit0: gpu request != 0
    DimDiv div[3];                    // executor cpu, mem, gpu: the request itself (div[t].e) and how to divide by it
};
static_assert(sizeof(ShapeEntry) == 96, "ShapeEntry layout");

// device-side header of one table set
struct ShapeHeader {
    int32_t n_shapes;                 // shapes claimed so far (may exceed kMaxShapes)
    int32_t fallback_apps;            // statistics: applications that took the scan path
    int32_t id_slot[kMaxShapes];      // dense id -> hash slot
};

struct ShapeTables {
    ShapeEntry* entries;              // [kShapeSlots]
    ShapeHeader* hdr;
    uint32_t* table;                  // [kMaxShapes][pitch] exclusive prefix per instance group, indexed by slot
    uint32_t* total;                  // [kMaxShapes][n_groups]
    int32_t pitch;                    // row length (>= n_slots, multiple of 4)
    int32_t n_groups;
};

// largest executor count / per-node capacity the uint32 prefix of a group can carry without wrapping
__device__ __forceinline__ uint32_t table_clamp(int32_t ne) {
    const uint32_t by_len = 0xFFFFFFFFu / (uint32_t)(ne > 0 ? ne : 1);
    return by_len < (uint32_t)kMaxCount ? by_len : (uint32_t)kMaxCount;
}

__device__ __forceinline__ unsigned long long shape_fingerprint(int64_t a, int64_t b, int64_t c) {
    unsigned long long z = (unsigned long long)a * 0x9E3779B97F4A7C15ull;
    z ^= ((unsigned long long)b + 0xBF58476D1CE4E5B9ull) * 0x94D049BB133111EBull;
    z ^= ((unsigned long long)c + 0x2545F4914F6CDD1Dull) * 0xD6E8FEB86659FD93ull;
    z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    return z ? z : 1ull;
}

// ---- raw application columns in either wire width --------------------------------------------------------------
// 64-bit layout: exact int64 quantities (gp_apps).  32-bit layout (gp_apps_wire.quantity_bits == 32): int32 millicores /
// int32 (bytes >> mem_shift) / int32 gpu units -- 28 bytes per application instead of 60.
struct AppColumns {
    const void* q[6];                 // drv cpu, drv mem, drv gpu, exe cpu, exe mem, exe gpu (gpu columns may be NULL = 0)
    const int32_t* count;
    const int32_t* group;             // or NULL (= 0)
    const int64_t* off;               // [n+1] ExecutorNodes offsets (the caller's, or the ones K1 derived)
    int32_t bits;                     // 64 or 32
    int32_t mem_shift;                // 32-bit layout only
    __device__ __forceinline__ int64_t load(int c, int32_t i) const {
        const void* p = q[c];
        if (!p) return 0;
        if (bits == 64) return static_cast<const int64_t*>(p)[i];
        const int64_t v = static_cast<const int32_t*>(p)[i];
        return (c == 1 || c == 4) ? (v < 0 ? v : (v << mem_shift)) : v;     // negative stays negative -> validation error
    }
};

// ---------------------------------------------------------------------------------------------------------------
// K1: intern executor shapes, derive offsets
// ---------------------------------------------------------------------------------------------------------------
constexpr int kClassifyThreads = 256;
__global__ void __launch_bounds__(kClassifyThreads) gp_classify_apps(int32_t n_apps, AppColumns cols, ShapeTables tabs,
                                                                     const SnapMeta* __restrict__ meta,
                                                                     int64_t off_base, int64_t* __restrict__ off_out /* or NULL */,
                                                                     int32_t* __restrict__ app_slot, int use_tables) {
    __shared__ unsigned long long s_part[kClassifyThreads / 32];
    const int32_t block0 = blockIdx.x * blockDim.x;
    const int32_t i = block0 + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // ---- ExecutorNodes offsets = exclusive prefix sum of max(count, 0), when the caller passed none.  Each CTA sums
    // the counts before its own block itself (L2-resident, <= n_apps loads per CTA): no inter-CTA dependency.
    if (off_out) {
        unsigned long long before = 0;
        for (int32_t t = threadIdx.x; t < block0; t += blockDim.x) { const int32_t c = cols.count[t]; before += c > 0 ? (unsigned)c : 0u; }
        const int32_t mine = (i < n_apps) ? max(cols.count[i], 0) : 0;
        // block reduce of `before`, block exclusive scan of `mine`
        unsigned long long wb = before;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) wb += __shfl_xor_sync(kFull, wb, d);
        uint32_t incl = warp_incl_scan((uint32_t)mine, lane);
        __shared__ unsigned long long s_before[kClassifyThreads / 32];
        if (lane == 31) s_part[w] = incl;
        if (lane == 0) s_before[w] = wb;
        __syncthreads();
        unsigned long long base = (unsigned long long)off_base, prev = 0;
        for (int t = 0; t < kClassifyThreads / 32; ++t) { base += s_before[t]; if (t < w) prev += s_part[t]; }
        if (i < n_apps) off_out[i] = (int64_t)(base + prev + incl - (uint32_t)mine);
        if (i == n_apps - 1) off_out[n_apps] = (int64_t)(base + prev + incl);
        __syncthreads();
    }
    if (i >= n_apps) return;
    if (!use_tables) { app_slot[i] = -1; return; }

    const int64_t e0 = cols.load(3, i), e1 = cols.load(4, i), e2 = cols.load(5, i);
    int32_t found = -1;
    if (e0 >= 0 && e1 >= 0 && e2 >= 0 && e0 < kMaxQuantity && e1 < kMaxQuantity && e2 < kMaxQuantity) {
        const unsigned long long fp = shape_fingerprint(e0, e1, e2);
        int32_t slot = (int32_t)(fp & (kShapeSlots - 1));
        for (int p = 0; p < kShapeProbes && found < 0; ++p) {
            ShapeEntry* en = tabs.entries + slot;
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&en->key);
            if (cur == 0) {
                cur = atomicCAS(&en->key, 0ull, fp);
                if (cur == 0) {
                    // this thread owns the entry: dense id, division recipes (verified by every consumer against its own tuple)
                    int32_t id = atomicAdd(&tabs.hdr->n_shapes, 1);
                    if (id >= kMaxShapes) id = -1;
                    int bad = 0; uint64_t l; bool fast = true;
                    en->div[0] = prep_dim(0, e0, 0, meta->max_avail[0], bad, l, fast);
                    en->div[1] = prep_dim(0, e1, 1, meta->max_avail[1], bad, l, fast);
                    en->div[2] = prep_dim(0, e2, 2, meta->max_avail[2], bad, l, fast);
                    en->flags = e2 != 0 ? 1u : 0u;
                    en->id = id;
                    if (id >= 0) tabs.hdr->id_slot[id] = slot;
                    cur = fp;
                }
            }
            if (cur == fp) found = slot;
            slot = (slot + 1) & (kShapeSlots - 1);
        }
    }
    app_slot[i] = found;
}

// ---------------------------------------------------------------------------------------------------------------
// K2: capacity prefix tables, one CTA per (shape, instance group)
// ---------------------------------------------------------------------------------------------------------------
struct TabScratch {
    unsigned long long bar[2];
    uint32_t part[2][kTabThreads / 32];
};
constexpr size_t kTabSmemBytes = 1024 + 2 * (size_t)kTabTile * sizeof(longlong2);      // 1 KB scratch + 2 x 64 KB tiles

template <int ALGO>
__global__ void __launch_bounds__(kTabThreads, 1) gp_build_shape_tables(Snapshot s, ShapeTables tabs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TabScratch& sh = *reinterpret_cast<TabScratch*>(smem_raw);
    longlong2* tile[2] = {reinterpret_cast<longlong2*>(smem_raw + 1024), reinterpret_cast<longlong2*>(smem_raw + 1024) + kTabTile};
    const int id = blockIdx.x, grp = blockIdx.y;
    const int n_shapes = min(tabs.hdr->n_shapes, kMaxShapes);
    if (id >= n_shapes) return;
    const ShapeEntry& en = tabs.entries[tabs.hdr->id_slot[id]];
    const DimDiv dc = en.div[0], dm = en.div[1], dg = en.div[2];
    const bool ug = (en.flags & 1u) || (s.meta->flags & kSnapGpuNegative);
    const GroupDesc g = s.groups[grp];
    const int32_t ne = g.ne;
    const uint32_t clamp = table_clamp(ne);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const longlong2* gpair = s.pair + g.sbase;
    const int64_t* ggpu = s.gpu + g.sbase;
    uint32_t* out = tabs.table + (size_t)id * tabs.pitch + g.sbase;

    if (tid == 0) { mbar_init(&sh.bar[0], 1); mbar_init(&sh.bar[1], 1); }
    __syncthreads();
    const int n_tiles = (ne + kTabTile - 1) / kTabTile;
    auto issue = [&](int t) {          // thread 0: TMA bulk copies of tile t into buffer t & 1 (<= 32 KB per copy)
        const int32_t lo = t * kTabTile;
        const uint32_t bytes = (uint32_t)min(kTabTile, ne - lo) * (uint32_t)sizeof(longlong2);
        mbar_expect_tx(&sh.bar[t & 1], bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
            tma_load_1d(reinterpret_cast<unsigned char*>(tile[t & 1]) + off, reinterpret_cast<const unsigned char*>(gpair + lo) + off, n,
                        &sh.bar[t & 1]);
        }
    };
    if (tid == 0 && n_tiles > 0) issue(0);
    uint32_t carry = 0;
    for (int t = 0; t < n_tiles; ++t) {
        if (tid == 0 && t + 1 < n_tiles) issue(t + 1);           // buffer (t+1)&1 was released by the barrier that ended tile t-1
        mbar_wait(&sh.bar[t & 1], (uint32_t)((t >> 1) & 1));
        const longlong2* tp = tile[t & 1];
        const int32_t lo = t * kTabTile;
        const int32_t i0 = lo + tid * kTabPerThread;
        uint32_t v[kTabPerThread];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < kTabPerThread; ++j) {
            const int32_t i = i0 + j;
            uint32_t c = 0;
            if (i < ne) {
                const longlong2 a = tp[i - lo];
                c = min(cap_dim(a.x, dc, clamp), cap_dim(a.y, dm, clamp));
                if (ug) c = min(c, cap_dim(__ldg(ggpu + i), dg, clamp));
                if (ALGO == 1) c = c != 0 ? 1u : 0u;
            }
            v[j] = sum;             // exclusive within the thread
            sum += c;
        }
        // block-wide exclusive scan of the per-thread sums
        const uint32_t incl = warp_incl_scan(sum, lane);
        if (lane == 31) sh.part[t & 1][w] = incl;
        __syncthreads();            // also: every thread is done reading tile[t & 1]... of the PREVIOUS use (see issue())
        const uint32_t pt = sh.part[t & 1][lane];
        const uint32_t pti = warp_incl_scan(pt, lane);
        const uint32_t before = __shfl_sync(kFull, pti - pt, w);
        const uint32_t tile_total = __shfl_sync(kFull, pti, 31);
        const uint32_t base = carry + before + incl - sum;
        if (i0 + kTabPerThread <= ne && ((g.sbase + i0) & 3) == 0) {
            *reinterpret_cast<uint4*>(out + i0) = make_uint4(base + v[0], base + v[1], base + v[2], base + v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < kTabPerThread; ++j) if (i0 + j < ne) out[i0 + j] = base + v[j];
        }
        carry += tile_total;
        __syncthreads();            // tile[t & 1] and part[t & 1] may be overwritten from here on
    }
    if (tid == 0) tabs.total[(size_t)id * tabs.n_groups + grp] = carry;
}

// ---------------------------------------------------------------------------------------------------------------
// K3: one warp per application
// ---------------------------------------------------------------------------------------------------------------
struct TabStats { unsigned long long probes, drivers, fallback; };

// prefix value S(i), i in [0, ne]
__device__ __forceinline__ uint32_t tab_at(const uint32_t* __restrict__ tab, int32_t i, int32_t ne, uint32_t total) {
    return i < ne ? __ldg(tab + i) : total;
}

// Largest i in [0, ne] with S(i) <= 0, narrowed to a window of <= 32 entries: returns lo such that the first node with
// a non-zero table increment lies in [lo, lo + 32).  Requires total > 0.
__device__ __forceinline__ int32_t tab_first_window(const uint32_t* __restrict__ tab, int32_t ne, uint32_t total, int lane, TabStats& st) {
    int32_t lo = 0, hi = ne;       // S(lo) == 0, S(hi) > 0
    while (hi - lo > 32) {
        const int32_t step = (hi - lo + 31) / 32;
        const int32_t p = lo + (lane + 1) * step;
        const bool zero = p < hi && tab_at(tab, p, ne, total) == 0;
        const unsigned z = __ballot_sync(kFull, zero);
        const int c = __ffs(~z) - 1;           // leading lanes whose probe is still 0 (S is non-decreasing)
        lo += c * step;
        hi = min(hi, lo + step);
        st.probes += 32;
    }
    return lo;
}

template <int ALGO, class OUT>
__device__ __forceinline__ int32_t pack_app_tables(const Snapshot& s, const GroupDesc& g, const uint32_t* __restrict__ tab, uint32_t total,
                                                   const ShapeEntry& en, int64_t d_cpu, int64_t d_mem, int64_t d_gpu, uint32_t k,
                                                   OUT* __restrict__ out, TabStats& st, int lane, int snap_flags, bool& need_scan) {
    const int32_t ne = g.ne;
    const uint32_t clamp = table_clamp(ne);
    need_scan = false;
    if (k > clamp) { need_scan = true; return -1; }
    if (ALGO == 0) { if (k != 0 && total < k) return -1; }                        // not even without a driver (pack_tightly.go:62)
    else if (k != 0 && total < k + 1) { need_scan = true; return -1; }            // fewer than k+1 hosting nodes: rounds / exact test

    const bool drv_gpu = d_gpu != 0 || (snap_flags & kSnapGpuNegative);
    const bool cap_gpu = drv_gpu || (en.flags & 1u);
    const longlong2* gpair = s.pair + g.sbase;
    const int64_t* ggpu = s.gpu + g.sbase;

    // ---- driver loop (binpack.go:67-85): first candidate that fits and leaves room for k executors ----------------
    int32_t dslot = -1;
    uint32_t cd = 0, c0d = 0, Sp = 0;
    for (int32_t j0 = 0; j0 < g.nd && dslot < 0; j0 += kWarp) {
        const int32_t j = j0 + lane;
        bool feasible = false;
        int32_t ls = -1;
        uint32_t my_cd = 0, my_c0 = 0, my_sp = 0;
        if (j < g.nd) {
            ls = s.drv_slot[g.dbase + j];
            const longlong2 v = __ldg(gpair + ls);
            const int64_t gv = drv_gpu || cap_gpu ? __ldg(ggpu + ls) : 0;
            feasible = !(d_cpu > v.x) && !(d_mem > v.y) && !(drv_gpu && d_gpu > gv);
            if (feasible && ls < ne && k != 0) {
                my_sp = __ldg(tab + ls);
                my_c0 = tab_at(tab, ls + 1, ne, total) - my_sp;
                if (my_c0 != 0) {        // what the node can still take once the driver sits on it
                    uint32_t c = min(cap_dim(v.x - d_cpu, en.div[0], clamp), cap_dim(v.y - d_mem, en.div[1], clamp));
                    if (cap_gpu) c = min(c, cap_dim(gv - d_gpu, en.div[2], clamp));
                    if (ALGO == 1) c = c != 0 ? 1u : 0u;
                    my_cd = c;
                }
                if (ALGO == 0) feasible = (total - (my_c0 - my_cd) >= k);
                // distribute-evenly: total >= k+1 hosting nodes, the driver removes at most its own -> always feasible
            }
        }
        const unsigned vote = __ballot_sync(kFull, feasible);
        st.drivers += (unsigned long long)((g.nd - j0) < kWarp ? (g.nd - j0) : kWarp);
        if (vote) {
            const int src = __ffs(vote) - 1;
            dslot = __shfl_sync(kFull, ls, src);
            cd = __shfl_sync(kFull, my_cd, src);
            c0d = __shfl_sync(kFull, my_c0, src);
            Sp = __shfl_sync(kFull, my_sp, src);
        }
    }
    if (dslot < 0) return -1;
    const int32_t driver_node = s.slot_node[g.sbase + dslot];
    if (k == 0) return driver_node;

    // ---- emission: output j belongs to the node whose adjusted prefix range holds j ------------------------------
    // adjusted inclusive prefix of node i:  S'(i+1) = S(i+1) - (i >= dslot ? delta : 0),  delta = c0(d) - cd(d)
    const uint32_t delta = (dslot < ne) ? c0d - cd : 0u;
    const int32_t dpos = (dslot < ne) ? dslot : 0x7fffffff;
    (void)Sp;
    int32_t pos = tab_first_window(tab, ne, total, lane, st);
    uint32_t placed = 0;
    const int32_t* slot_node = s.slot_node + g.sbase;
    while (placed < k && pos < ne) {
        const int32_t i = pos + lane;
        uint32_t Sv = 0xFFFFFFFFu;                        // lanes beyond the order: never selected
        int32_t node = -1;
        if (i < ne) {
            Sv = tab_at(tab, i + 1, ne, total);
            if (i >= dpos) Sv -= delta;
            node = __ldg(slot_node + i);
        }
        st.probes += 32;
        // last valid lane's value bounds what this window can place
        const int lastl = min(kWarp - 1, ne - 1 - pos);
        const uint32_t wend = __shfl_sync(kFull, Sv, lastl);
        const uint32_t endv = wend < k ? wend : k;
        for (uint32_t j0 = placed; j0 < endv; j0 += kWarp) {
            const uint32_t j = j0 + lane;
            int lo = 0;                                   // first lane whose adjusted inclusive prefix exceeds j
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(kFull, Sv, lo + step - 1);
                if (v <= j) lo += step;
            }
            const int32_t nd = __shfl_sync(kFull, node, lo & 31);
            if (j < endv) out[j] = (OUT)nd;
        }
        placed = endv > placed ? endv : placed;
        pos += kWarp;
    }
    return driver_node;
}

// The node-order scan of gangpack_kernels.cuh for ONE application inside the fused kernel (out of line: its register
// needs must not weigh on the table path).  The prepared record is built in shared memory, lanes 0..2 one dimension each.
template <int ALGO, class OUT>
__device__ __noinline__ int32_t scan_path_app(const Snapshot& s, PrepApp* pa, int64_t d_cpu, int64_t d_mem, int64_t d_gpu,
                                              int64_t e_cpu, int64_t e_mem, int64_t e_gpu, int32_t kk, int32_t grp,
                                              OUT* __restrict__ out, int2* __restrict__ list, uint16_t* __restrict__ wcache,
                                              WarpStats& st, int lane, int snap_flags) {
    int b2 = 0; uint64_t l = 0; bool fast = true;
    if (lane < 3) {
        const int64_t dd = lane == 0 ? d_cpu : (lane == 1 ? d_mem : d_gpu);
        const int64_t ee = lane == 0 ? e_cpu : (lane == 1 ? e_mem : e_gpu);
        pa->div[lane] = prep_dim(dd, ee, lane, s.meta->max_avail[lane], b2, l, fast);
        pa->drv[lane] = dd;
    }
    unsigned long long lm = l;
    lm = max(lm, __shfl_xor_sync(kFull, lm, 1)); lm = max(lm, __shfl_xor_sync(kFull, lm, 2));   // lanes 0..3 (lane 3 holds 0)
    const bool all_fast = __all_sync(kFull, fast);
    __syncwarp();
    if (lane == 0) {
        pa->out_off = 0;                    // `out` / `list` already point at this application's slice
        pa->count = kk; pa->group = grp;
        pa->lmax = (int32_t)(lm < (unsigned long long)(uint32_t)kk ? lm : (unsigned long long)(uint32_t)kk);
        const bool f32 = prep_fast32(all_fast, pa->div[0], pa->div[1], s.meta);
        pa->flags = ((d_gpu != 0 || e_gpu != 0) ? kAppUsesGpu : 0u) | (all_fast ? kAppFast : 0u) | (f32 ? kAppFast32 : 0u);
    }
    __syncwarp();
    const GroupDesc g0 = s.groups[0];
    const int32_t d = pack_app<ALGO, OUT>(s, pa, out, list, wcache, st, lane, snap_flags, g0);
    __syncwarp();
    return d;
}

constexpr int kPackTabThreads = 256;
#ifndef GP_TAB_MIN_BLOCKS
#define GP_TAB_MIN_BLOCKS 4
#endif
// cols: raw application columns; app_slot from K1; `force_scan` routes every application to the scan path.
template <int ALGO, class OUT>
__global__ void __launch_bounds__(kPackTabThreads, GP_TAB_MIN_BLOCKS) gp_pack_tables(Snapshot s, AppColumns cols, ShapeTables tabs,
                                                                                     const int32_t* __restrict__ app_slot, int32_t n_apps,
                                                                                     int64_t out_cap, int32_t* __restrict__ driver_node,
                                                                                     OUT* __restrict__ executor_nodes,
                                                                                     int2* __restrict__ scratch,
                                                                                     unsigned long long* __restrict__ stats,
                                                                                     unsigned int* __restrict__ next_app,
                                                                                     int* __restrict__ err, volatile int* __restrict__ err_host) {
    __shared__ uint16_t cap_cache[kPackTabThreads / 32][kCapCache];
    __shared__ PrepApp prep_sm[kPackTabThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint16_t* wcache = cap_cache[w];
    PrepApp* pa = &prep_sm[w];
    TabStats ts{0, 0, 0};
    WarpStats st{0, 0};
    const int snap_flags = s.meta->flags;
    const GroupDesc g0 = s.groups[0];
    unsigned int i = 0, n1 = 0;
    if (lane == 0) { i = atomicAdd(next_app, 1u); n1 = atomicAdd(next_app, 1u); }
    i = __shfl_sync(kFull, i, 0);
    n1 = __shfl_sync(kFull, n1, 0);
    while (i < (unsigned int)n_apps) {
        unsigned int n2 = 0;
        if (lane == 0) n2 = atomicAdd(next_app, 1u);
        // ---- the raw tuple: one field per lane, then broadcast ------------------------------------------------
        int64_t f = 0;
        if (lane < 6) f = cols.load(lane, (int32_t)i);
        else if (lane == 6) f = cols.count[i];
        else if (lane == 7) f = cols.group ? cols.group[i] : 0;
        else if (lane == 8) f = cols.off[i];
        else if (lane == 9) f = cols.off[i + 1];
        else if (lane == 10) f = app_slot[i];
        const int64_t d_cpu = __shfl_sync(kFull, f, 0), d_mem = __shfl_sync(kFull, f, 1), d_gpu = __shfl_sync(kFull, f, 2);
        const int64_t e_cpu = __shfl_sync(kFull, f, 3), e_mem = __shfl_sync(kFull, f, 4), e_gpu = __shfl_sync(kFull, f, 5);
        const int32_t kk = (int32_t)__shfl_sync(kFull, f, 6);
        const int32_t grp = (int32_t)__shfl_sync(kFull, f, 7);
        const int64_t off = __shfl_sync(kFull, f, 8), off1 = __shfl_sync(kFull, f, 9);
        const int32_t slot = (int32_t)__shfl_sync(kFull, f, 10);
        // ---- validation (types.SparkApplicationResources must be non-negative and inside the exact-int64 domain) ----
        int bad = 0;
        if (lane < 6) { if (f < 0) bad |= kErrNegativeRequest; if (f >= kMaxQuantity) bad |= kErrUnrepresentable; }
        if (kk < 0) bad |= kErrNegativeRequest;
        if (kk > kMaxCount) bad |= kErrUnrepresentable;
        if (grp < 0 || grp >= s.n_groups) bad |= kErrBadGroup;
        if (off < 0 || off1 - off != (int64_t)(kk > 0 ? kk : 0) || off1 > out_cap) bad |= kErrBadOffsets;
        bad = __reduce_or_sync(kFull, bad);
        int32_t d = -1;
        if (bad) {
            if (lane == 0) { atomicOr(err, bad); *err_host = bad; }
        } else {
            const GroupDesc g = grp == 0 ? g0 : s.groups[grp];
            const uint32_t k = (uint32_t)kk;
            bool need_scan = true;
            if (slot >= 0) {
                const ShapeEntry& en = tabs.entries[slot];
                // the full tuple decides, not the fingerprint
                if (en.id >= 0 && en.div[0].e == e_cpu && en.div[1].e == e_mem && en.div[2].e == e_gpu) {
                    const uint32_t* tab = tabs.table + (size_t)en.id * tabs.pitch + g.sbase;
                    const uint32_t total = tabs.total[(size_t)en.id * tabs.n_groups + grp];
                    d = pack_app_tables<ALGO, OUT>(s, g, tab, total, en, d_cpu, d_mem, d_gpu, k, executor_nodes + off, ts, lane,
                                                   snap_flags, need_scan);
                }
            }
            if (need_scan) {
                ts.fallback += 1;
                d = scan_path_app<ALGO, OUT>(s, pa, d_cpu, d_mem, d_gpu, e_cpu, e_mem, e_gpu, kk, grp, executor_nodes + off,
                                             scratch ? scratch + off : nullptr, wcache, st, lane, snap_flags);
            }
        }
        if (lane == 0) driver_node[i] = d;
        i = n1;
        n1 = __shfl_sync(kFull, n2, 0);
    }
    if (lane == 0) {
        atomicAdd(stats + 0, st.nodes + ts.probes);
        atomicAdd(stats + 1, st.drivers + ts.drivers);
        atomicAdd(stats + 2, ts.fallback);
        atomicAdd(stats + 3, st.nodes);
    }
}

}  // namespace gp

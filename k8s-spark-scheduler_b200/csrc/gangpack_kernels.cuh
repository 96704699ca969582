// gangpack_kernels.cuh -- sm_100a device code of the gang-placement hot path.
//
// What is computed (reference, all under /root/reference; LIB = vendor/github.com/palantir/
// k8s-spark-scheduler-lib/pkg):
//   SparkBinPack driver loop            LIB/binpack/binpack.go:60-87
//   tightlyPackExecutors                LIB/binpack/pack_tightly.go:34-63
//   distributeExecutorsEvenly           LIB/binpack/distribute_evenly.go:34-73
//   Resources.GreaterThan / Add / Sub   LIB/resources/resources.go:239-241, 202-213
//   fitEarlierDrivers + sparkResourceUsage + SubtractUsageIfExists
//                                       internal/extender/resource.go:224-262,
//                                       internal/extender/sparkpods.go:139-146,
//                                       LIB/resources/resources.go:129-135
//
// How (B200-first, not a translation): the reference re-runs the executor loop for every driver
// candidate over string-keyed maps.  Here a node's executor capacity is a closed form
// (the reference's own LIB/capacity/capacity.go:36-75):
//     cap_dim(n|r) = 0 if r > avail ; INF if exe == 0 ; floor((avail - r) / exe)
//     cap(n|r)     = min over cpu, mem, gpu
// so one WARP owns one pending application, lanes own consecutive nodes of the executor
// priority order (coalesced 128-bit (cpu,mem) loads), a warp reduction/prefix scan counts the
// executors that fit, lanes trial-place the driver on 32 candidates at a time and
// __ballot_sync picks the first feasible one, and a second pass emits ExecutorNodes.
// Division by the (warp-uniform) executor request is a multiply-high by a per-app magic
// number computed once by gp_prep_apps (exact, see cap_dim()).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gp {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int32_t kMaxCount = 1 << 24;            // exe_count limit (keeps 32-lane int32 sums exact)
constexpr int64_t kMaxQuantity = (int64_t)1 << 61; // |quantity| limit of the exact-int64 domain

// error bits raised by device-side validation (read back by the host API)
enum : int {
    kErrNegativeRequest = 1,   // driver/executor request < 0 or exe_count < 0
    kErrUnrepresentable = 2,   // |quantity| >= 2^61 or exe_count > 2^24
    kErrBadGroup = 4,          // app group outside [0, n_groups)
    kErrBadOffsets = 8,        // exec_out_off inconsistent with exe_count / capacity
};

// snapshot flags (device int)
enum : int { kSnapGpuNegative = 1 };

// How to divide by one executor-request dimension.
enum : uint32_t { kDivInf = 0, kDivMagic = 1, kDivShift = 2, kDivSlow = 3 };

struct DimDiv {
    uint64_t magic;  // floor((2^64-1)/e') + 1 for e = e' << sh, e' < 2^32   (kDivMagic)
    int64_t e;       // the request itself
    uint32_t sh;     // trailing zero bits of e
    uint32_t kind;
};

// One pending application, prepared (128 B).
struct __align__(16) PrepApp {
    int64_t drv[3];      // driver cpu, mem, gpu
    DimDiv div[3];       // executor cpu, mem, gpu
    int64_t out_off;     // first ExecutorNodes slot of this app
    int32_t count;       // MinExecutorCount
    int32_t group;
    int32_t lmax;        // max executors the driver can displace on its own node (<= count)
    uint32_t flags;      // bit0: uses gpu dim, bit1: skip_if_no_fit, bit2: invalid
};
static_assert(sizeof(PrepApp) == 128, "PrepApp layout");

struct GroupDesc {
    int32_t sbase;   // first slot of the group
    int32_t ne;      // executor-order length: slots [sbase, sbase+ne) in priority order
    int32_t dbase;   // first entry in drv_slot
    int32_t nd;      // driver-order length
};

// Device snapshot: "slots" = per group the executor order followed by one spare slot per driver
// candidate (used only when that candidate is not an executor candidate).
struct Snapshot {
    longlong2* pair;       // [n_slots] (avail cpu milli, avail mem bytes)
    int64_t* gpu;          // [n_slots]
    int32_t* slot_node;    // [n_slots] caller's node index, -1 = unused spare
    const int32_t* drv_slot; // [n_drv] group-local slot of each driver candidate
    const GroupDesc* groups;
    const int* flags;      // kSnap*
    int32_t n_groups;
    int32_t n_slots;
};

// ---------------------------------------------------------------------------------------------
// capacity arithmetic
// ---------------------------------------------------------------------------------------------

// Executors of request `p.e` that fit into `a` free units, clamped to k.
// Exactness: a >= 0, e = e' << sh  =>  floor(a/e) = floor((a >> sh)/e').  For xs = a >> sh < 2^32
// and e' < 2^32, floor(xs/e') = mulhi64(magic, xs) with magic = floor((2^64-1)/e') + 1
// (Lemire/Kaser/Kurz 2019, N = 32).  Anything else takes the 64-bit divide.
__device__ __forceinline__ uint32_t cap_dim(int64_t a, const DimDiv& p, uint32_t k) {
    if (a < 0) return 0;                       // reserved(=0) > avail in this dim (resources.go:239)
    if (p.kind == kDivInf) return k;           // zero request never limits
    uint64_t xs = (uint64_t)a >> p.sh;
    uint64_t q;
    if (p.kind == kDivShift) q = xs;
    else if (p.kind == kDivMagic && (xs >> 32) == 0) q = __umul64hi(p.magic, xs);
    else q = (uint64_t)a / (uint64_t)p.e;
    return q < (uint64_t)k ? (uint32_t)q : k;
}

// cap(n | reserved = r) for a single node with an arbitrary reservation (the driver's node).
__device__ __forceinline__ uint32_t cap_dim_reserved(int64_t a, int64_t r, int64_t e, uint32_t k) {
    if (r > a) return 0;
    if (e == 0) return k;
    uint64_t q = (uint64_t)(a - r) / (uint64_t)e;
    return q < (uint64_t)k ? (uint32_t)q : k;
}

template <bool MUTABLE>
__device__ __forceinline__ longlong2 load_pair(const longlong2* p) {
    if (MUTABLE) {
        longlong2 v;
        asm volatile("ld.global.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
        return v;
    } else {
        return __ldg(p);
    }
}
template <bool MUTABLE>
__device__ __forceinline__ int64_t load_gpu(const int64_t* p) {
    if (MUTABLE) {
        int64_t v;
        asm volatile("ld.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    } else {
        return __ldg(p);
    }
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(kFull, v); }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// per-application state shared by the phases
// ---------------------------------------------------------------------------------------------
struct AppRegs {
    int64_t d_cpu, d_mem, d_gpu;
    DimDiv cpu, mem, gpu;
    int64_t out_off;
    uint32_t k;
    uint32_t lmax;
    bool use_gpu;
};

template <bool MUTABLE>
__device__ __forceinline__ uint32_t cap0_at(const Snapshot& s, int32_t slot, const AppRegs& a) {
    longlong2 v = load_pair<MUTABLE>(s.pair + slot);
    uint32_t c = cap_dim(v.x, a.cpu, a.k);
    uint32_t m = cap_dim(v.y, a.mem, a.k);
    c = c < m ? c : m;
    if (a.use_gpu) {
        uint32_t g = cap_dim(load_gpu<MUTABLE>(s.gpu + slot), a.gpu, a.k);
        c = c < g ? c : g;
    }
    return c;
}

// cap(d | drv) for the chosen driver's slot
template <bool MUTABLE>
__device__ __forceinline__ uint32_t capd_at(const Snapshot& s, int32_t slot, const AppRegs& a) {
    longlong2 v = load_pair<MUTABLE>(s.pair + slot);
    uint32_t c = cap_dim_reserved(v.x, a.d_cpu, a.cpu.e, a.k);
    uint32_t m = cap_dim_reserved(v.y, a.d_mem, a.mem.e, a.k);
    c = c < m ? c : m;
    // the gpu dim must be honoured whenever it can bind; with !use_gpu it cannot (request 0, avail >= 0)
    if (a.use_gpu) {
        uint32_t g = cap_dim_reserved(load_gpu<MUTABLE>(s.gpu + slot), a.d_gpu, a.gpu.e, a.k);
        c = c < g ? c : g;
    }
    return c;
}

// driverResources.GreaterThan(available) == false  (binpack.go:69)
template <bool MUTABLE>
__device__ __forceinline__ bool driver_fits(const Snapshot& s, int32_t slot, const AppRegs& a) {
    longlong2 v = load_pair<MUTABLE>(s.pair + slot);
    bool ok = !(a.d_cpu > v.x) && !(a.d_mem > v.y);
    if (a.use_gpu) ok = ok && !(a.d_gpu > load_gpu<MUTABLE>(s.gpu + slot));
    return ok;
}

struct WarpStats { unsigned long long nodes; unsigned long long drivers; };

// ---------------------------------------------------------------------------------------------
// One application, one warp.  Returns the driver's node index (>= 0) or -1.
// ALGO: 0 tightly-pack, 1 distribute-evenly.  FIFO_MODE: 0 none, 1 reference usage, 2 exact usage
// (then the snapshot is read with plain loads and the placement is subtracted from it).
// ---------------------------------------------------------------------------------------------
template <int ALGO, int FIFO_MODE>
__device__ int32_t pack_app(const Snapshot& s, const PrepApp* __restrict__ pa, int32_t* __restrict__ executor_nodes,
                            int2* __restrict__ scratch, WarpStats& st, int lane) {
    constexpr bool MUT = FIFO_MODE != 0;
    AppRegs a;
    a.d_cpu = pa->drv[0]; a.d_mem = pa->drv[1]; a.d_gpu = pa->drv[2];
    a.cpu = pa->div[0]; a.mem = pa->div[1]; a.gpu = pa->div[2];
    a.out_off = pa->out_off;
    a.k = (uint32_t)pa->count;
    a.lmax = (uint32_t)pa->lmax;
    const uint32_t flags = pa->flags;
    a.use_gpu = (flags & 1u) || ((*s.flags) & kSnapGpuNegative);
    const GroupDesc g = s.groups[pa->group];
    const uint32_t k = a.k;
    const int32_t ne = g.ne;
    int32_t* out = executor_nodes + a.out_off;
    int2* list = scratch ? scratch + a.out_off : nullptr;   // distribute-evenly candidate list

    // ---- phase 1: how many executors fit, walking the executor priority order --------------------
    // P = sum min(cap(n|0), k) over the scanned prefix; m1 = #nodes with cap >= 1.
    // Early exit is exact: once P >= k + lmax (or m1 >= k+1) every driver candidate that fits
    // leaves >= k executor slots inside the prefix (the driver displaces <= lmax of them, all on one
    // node).  distribute-evenly may only stop on m1 >= k+1 (then one round places everything).
    unsigned long long P = 0;
    uint32_t m1 = 0;
    int32_t pos = 0;
    bool early = (k == 0);
    const unsigned long long need = (unsigned long long)k + a.lmax;
    while (!early && pos < ne) {
        int32_t i = pos + lane;
        uint32_t c = (i < ne) ? cap0_at<MUT>(s, g.sbase + i, a) : 0u;
        unsigned has = __ballot_sync(kFull, c != 0);
        if (ALGO == 1) {
            // remember the first k nodes that can host at all: (position, cap)
            uint32_t r = m1 + __popc(has & ((1u << lane) - 1u));
            if (c != 0 && r < k) list[r] = make_int2(i, (int)c);
        }
        P += warp_sum(c);
        m1 += __popc(has);
        pos += kWarp;
        if (ALGO == 0) early = (P >= need) || (m1 >= k + 1);
        else early = (m1 >= k + 1);
    }
    st.nodes += (unsigned long long)(pos < ne ? pos : ne);
    const bool exact_total = !early;          // scanned everything: P == S0
    if (exact_total && P < k) return -1;      // not even without a driver

    // ---- phase 2: first feasible driver candidate (binpack.go:67-85) ----------------------------
    int32_t dslot = -1;
    uint32_t cd = 0;      // cap(d | drv), clamped
    for (int32_t j0 = 0; j0 < g.nd && dslot < 0; j0 += kWarp) {
        int32_t j = j0 + lane;
        bool feasible = false;
        int32_t ls = -1;
        uint32_t lcd = 0;
        if (j < g.nd) {
            ls = s.drv_slot[g.dbase + j];
            feasible = driver_fits<MUT>(s, g.sbase + ls, a);
            if (feasible && ls < ne) {
                lcd = capd_at<MUT>(s, g.sbase + ls, a);
                if (exact_total) {
                    uint32_t c0 = cap0_at<MUT>(s, g.sbase + ls, a);
                    feasible = (P - c0 + lcd >= k);
                }
            }
        }
        unsigned vote = __ballot_sync(kFull, feasible);
        st.drivers += (unsigned long long)((g.nd - j0) < kWarp ? (g.nd - j0) : kWarp);
        if (vote) {
            int src = __ffs(vote) - 1;
            dslot = __shfl_sync(kFull, ls, src);
            cd = __shfl_sync(kFull, lcd, src);
        }
    }
    if (dslot < 0) return -1;
    const int32_t driver_node = s.slot_node[g.sbase + dslot];

    // ---- phase 3: emit ExecutorNodes -------------------------------------------------------------
    bool driver_hosts_executor = false;
    if (k != 0) {
        if (ALGO == 0) {
            // node-major: node n receives min(cap_d(n), remaining)  (pack_tightly.go:45-61)
            uint32_t placed = 0;
            for (int32_t p0 = 0; placed < k && p0 < ne; p0 += kWarp) {
                st.nodes += (unsigned long long)((ne - p0) < kWarp ? (ne - p0) : kWarp);
                int32_t i = p0 + lane;
                uint32_t c = 0;
                if (i < ne) c = (i == dslot) ? cd : cap0_at<MUT>(s, g.sbase + i, a);
                uint32_t incl = warp_incl_scan(c, lane);
                uint32_t total = __shfl_sync(kFull, incl, kWarp - 1);
                uint32_t room = k - placed;
                uint32_t T = total < room ? total : room;
                uint32_t excl = incl - c;
                uint32_t take = excl >= T ? 0u : ((c < T - excl) ? c : (T - excl));
                int32_t node = (i < ne) ? s.slot_node[g.sbase + i] : -1;
                // cooperative expansion: output j belongs to the first lane with incl > j
                for (uint32_t j = lane; j < ((T + kWarp - 1) & ~(uint32_t)(kWarp - 1)); j += kWarp) {
                    int lo = 0;
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        uint32_t v = __shfl_sync(kFull, incl, lo + step - 1);
                        if (v <= j) lo += step;
                    }
                    int32_t nd = __shfl_sync(kFull, node, lo & 31);
                    if (j < T) out[placed + j] = nd;
                }
                if (FIFO_MODE != 0 && take != 0) {
                    if (i == dslot) driver_hosts_executor = true;
                    long long mult = (FIFO_MODE == 1) ? 1 : (long long)take;
                    longlong2* pp = s.pair + g.sbase + i;
                    longlong2 v = *pp;
                    v.x -= mult * a.cpu.e; v.y -= mult * a.mem.e;
                    *pp = v;
                    s.gpu[g.sbase + i] -= mult * a.gpu.e;
                }
                placed += T;
            }
        } else if (early) {
            // one round: the first k nodes with cap_d >= 1, in order  (distribute_evenly.go:49-70)
            uint32_t placed = 0;
            for (int32_t p0 = 0; placed < k && p0 < ne; p0 += kWarp) {
                st.nodes += (unsigned long long)((ne - p0) < kWarp ? (ne - p0) : kWarp);
                int32_t i = p0 + lane;
                uint32_t c = 0;
                if (i < ne) c = (i == dslot) ? cd : cap0_at<MUT>(s, g.sbase + i, a);
                unsigned has = __ballot_sync(kFull, c != 0);
                uint32_t r = placed + __popc(has & ((1u << lane) - 1u));
                if (c != 0 && r < k) {
                    out[r] = s.slot_node[g.sbase + i];
                    if (FIFO_MODE != 0) {
                        if (i == dslot) driver_hosts_executor = true;
                        longlong2* pp = s.pair + g.sbase + i;
                        longlong2 v = *pp;
                        v.x -= a.cpu.e; v.y -= a.mem.e;
                        *pp = v;
                        s.gpu[g.sbase + i] -= a.gpu.e;
                    }
                }
                placed += __popc(has);
            }
        } else {
            // general rounds over the complete candidate list (m1 <= k entries, in order):
            // R* = min r with sum min(c, r) >= k; node gets min(c, R*-1) (+1 for the first
            // k - sum min(c, R*-1) nodes with c >= R*); ExecutorNodes is round-major.
            __syncwarp();
            const uint32_t m = m1;
            // patch the driver's own entry with cap(d|drv)
            for (uint32_t t = lane; t < m; t += kWarp) {
                int2 e = list[t];
                if (e.x == dslot) { e.y = (int)cd; list[t] = e; }
            }
            __syncwarp();
            uint32_t lo = 1, hi = k;   // f(k) >= k is known (feasible)
            while (lo < hi) {
                uint32_t mid = lo + (hi - lo) / 2;
                unsigned long long f = 0;
                for (uint32_t t0 = 0; t0 < m; t0 += kWarp) {
                    uint32_t t = t0 + lane;
                    uint32_t c = (t < m) ? (uint32_t)list[t].y : 0u;
                    f += warp_sum(c < mid ? c : mid);
                }
                if (f >= k) hi = mid; else lo = mid + 1;
            }
            const uint32_t R = lo;
            uint32_t base = 0;
            for (uint32_t r = 1; r <= R && base < k; ++r) {
                for (uint32_t t0 = 0; t0 < m && base < k; t0 += kWarp) {
                    uint32_t t = t0 + lane;
                    int2 e = (t < m) ? list[t] : make_int2(0, 0);
                    bool in = (t < m) && ((uint32_t)e.y >= r);
                    unsigned has = __ballot_sync(kFull, in);
                    uint32_t idx = base + __popc(has & ((1u << lane) - 1u));
                    if (in && idx < k) {
                        out[idx] = s.slot_node[g.sbase + e.x];
                        if (FIFO_MODE == 2 || (FIFO_MODE == 1 && r == 1)) {
                            longlong2* pp = s.pair + g.sbase + e.x;
                            longlong2 v = *pp;
                            v.x -= a.cpu.e; v.y -= a.mem.e;
                            *pp = v;
                            s.gpu[g.sbase + e.x] -= a.gpu.e;
                        }
                        if (FIFO_MODE != 0 && e.x == dslot) driver_hosts_executor = true;
                    }
                    base += __popc(has);
                }
            }
        }
    }

    // ---- FIFO: charge the driver (sparkpods.go:139-146 / exact) -----------------------------------
    if (FIFO_MODE != 0) {
        bool hosted = __any_sync(kFull, driver_hosts_executor);
        __syncwarp();   // executor charges (other lanes) are ordered before the driver charge
        if (lane == 0 && (FIFO_MODE == 2 || !hosted)) {
            longlong2* pp = s.pair + g.sbase + dslot;
            longlong2 v = *pp;
            v.x -= a.d_cpu; v.y -= a.d_mem;
            *pp = v;
            s.gpu[g.sbase + dslot] -= a.d_gpu;
        }
        __syncwarp();   // the next application of this queue sees the charged snapshot
    }
    return driver_node;
}

}  // namespace gp

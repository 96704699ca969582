// gangpack_minfrag.cuh -- minimal-fragmentation executor placement, one warp per application.
//
// Reference (all under /root/reference; LIB = vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg):
//   MinimalFragmentation / minimalFragmentation / internalMinimalFragmentation   LIB/binpack/minimal_fragmentation.go:27-137
//   GetNodeCapacities / GetNodeCapacity / FilterOutNodesWithoutCapacity          LIB/capacity/capacity.go:36-113
//   SparkBinPack driver loop                                                     LIB/binpack/binpack.go:60-87
//
// The reference sorts the (node, capacity) list of every driver candidate and then peels runs of equal
// capacity off its tail.  Here nothing is sorted.  With c(i) the UNCLAMPED capacity of the i-th executor
// candidate (nodes with c = 0 dropped, capacity.go:105-113):
//   * a driver candidate is feasible iff sum c >= k -- the same test as tightly-pack, so the driver loop is the
//     closed form of gangpack_kernels.cuh;
//   * M = max c.  If k < M the reference first tries the subset {c < target}, target = (k + M) / 2 in Go's wrapping
//     int arithmetic (:80-89); that attempt succeeds iff the subset's capacities add up to k.  Otherwise every node is
//     used (:93).  Because target >= k, the subset is either answered by the smallest c >= k or equals {c < k}: no
//     extra pass (see minfrag_emit);
//   * inside the chosen set: if some c >= k, all k executors go to the smallest such c (earliest node among equals,
//     :106-113).  Otherwise nodes are consumed whole in (c descending, priority order ascending) while the remainder is
//     >= c (:116-127): with F(v) = sum of c over {c >= v}, v* = max{v : F(v) > k} (binary search, one reduction pass per
//     probe), every node with c > v* is consumed, r* = k - F(v*+1), the first m = floor(r*/v*) nodes of the run c == v*
//     are consumed when v* < r*, and the rest r' goes to the smallest c >= r' among the unconsumed nodes;
//   * ExecutorNodes lists the consumed nodes in (c descending, order ascending): the offset of a consumed node is the sum
//     of the capacities ranked before it (enumeration over the <= k consumed entries), then r' copies of the last node.
// Capacities are recomputed from the L1/L2-resident snapshot in every pass (8-16 bytes and two multiply-highs per node);
// only the consumed entries (<= k) are written to scratch.
#pragma once

#include "gangpack_kernels.cuh"

namespace gp {

constexpr uint64_t kCapInf = 0x7fffffffffffffffull;   // math.MaxInt (capacity.go:42-45)

// unclamped cap_dim (see gangpack_kernels.cuh): a = avail - reserved
__device__ __forceinline__ uint64_t cap_dim_u(int64_t a, const DimDiv& p) {
    if (a < 0) return 0;
    if (p.kind == kDivInf) return kCapInf;
    const uint64_t xs = (uint64_t)a >> p.sh;
    if (p.kind == kDivShift) return xs;
    if (p.kind == kDivMagic && (xs >> 32) == 0) return __umul64hi(p.magic, xs);
    return udiv64((uint64_t)a, (uint64_t)p.e);
}

// ---- capacity providers: cap0(i) = unclamped capacity of executor candidate i with nothing reserved -------------
template <bool MUT>
struct MfCapGeneralT {           // any int64 request, gpu dimension on demand; MUT: the snapshot changes inside the launch
    typedef uint64_t T;
    DimDiv cpu, mem, gpu;
    const longlong2* pair;       // group base applied
    const int64_t* gpuv;
    bool ug;
    __device__ __forceinline__ T cap(int32_t i, int64_t rc, int64_t rm, int64_t rg) const {
        const longlong2 v = load_pair<MUT>(pair + i);
        T c = min(cap_dim_u(v.x - rc, cpu), cap_dim_u(v.y - rm, mem));
        if (ug) c = min(c, cap_dim_u(load_gpu<MUT>(gpuv + i) - rg, gpu));
        return c;
    }
    __device__ __forceinline__ T cap0(int32_t i) const { return cap(i, 0, 0, 0); }
};
typedef MfCapGeneralT<false> MfCapGeneral;
struct MfCapFast32 {             // fast class, gpu idle, compact 32-bit view: capacities < 2^32
    typedef uint32_t T;
    uint32_t mc_lo, mc_hi, mm_lo, mm_hi, shc, shm;
    const uint2* pair32;         // group base applied
    __device__ __forceinline__ T cap0(int32_t i) const {
        const uint2 v = __ldg(pair32 + i);
        const uint32_t xc = v.x >> shc, xm = v.y >> shm;
        const uint32_t qc = (uint32_t)(((uint64_t)mc_hi * xc + __umulhi(mc_lo, xc)) >> 32);
        const uint32_t qm = (uint32_t)(((uint64_t)mm_hi * xm + __umulhi(mm_lo, xm)) >> 32);
        return min(qc, qm);
    }
};

__device__ __forceinline__ uint64_t warp_max_t(uint64_t v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { uint64_t o = __shfl_xor_sync(kFull, v, d); v = o > v ? o : v; }
    return v;
}
__device__ __forceinline__ uint32_t warp_max_t(uint32_t v) { return __reduce_max_sync(kFull, v); }

// lexicographic minimum of (cap, pos) over the warp; "none" = pos < 0
template <class T>
__device__ __forceinline__ void warp_min_key(T& cap, int32_t& pos) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const T oc = __shfl_xor_sync(kFull, cap, d);
        const int32_t op = __shfl_xor_sync(kFull, pos, d);
        const bool take = op >= 0 && (pos < 0 || oc < cap || (oc == cap && op < pos));
        if (take) { cap = oc; pos = op; }
    }
}
template <class T>
__device__ __forceinline__ void note_min_key(T c, int32_t i, T& cap, int32_t& pos) {
    if (pos < 0 || c < cap) { cap = c; pos = i; }     // a lane visits its nodes in increasing order: ties keep the earlier one
}

// Executors of one application for driver slot `dslot` (group-local; >= ne when the driver's node is not an executor
// candidate) whose own capacity with the driver reserved is `cd`.  Feasibility (sum c >= k) is established by the
// caller.  list: k entries of scratch.
template <class CP>
__device__ __forceinline__ void minfrag_emit(const CP& cp, int32_t ne, int32_t dslot, typename CP::T cd, uint32_t k,
                                             const int32_t* __restrict__ slot_node, int32_t* __restrict__ out,
                                             int2* __restrict__ list, WarpStats& st, int lane) {
    typedef typename CP::T T;
    const T kT = (T)k;
    auto capd = [&](int32_t i) -> T { return i == dslot ? cd : cp.cap0(i); };

    // ---- pass 1 (the only full pass most applications need): M, the smallest c >= k, and sum / max of {c < k} ----
    T M = 0, max_lt = 0;
    unsigned long long sum_lt = 0;           // sum of c over {c < k}
    T best_c = 0; int32_t best_p = -1;       // smallest c >= k, earliest position among equals
    for (int32_t p0 = 0; p0 < ne; p0 += kWarp) {
        const int32_t i = p0 + lane;
        const T c = i < ne ? capd(i) : (T)0;
        M = c > M ? c : M;
        const T lt = c < kT ? c : (T)0;
        max_lt = lt > max_lt ? lt : max_lt;
        sum_lt += warp_sum((uint32_t)lt);
        if (c >= kT) note_min_key(c, i, best_c, best_p);
    }
    st.nodes += (unsigned long long)ne;
    M = warp_max_t(M);
    warp_min_key(best_c, best_p);
    // The subset {c < target}, target = (executorCount + maxCapacity) / 2 in Go's wrapping 64-bit int
    // (minimal_fragmentation.go:80-89), needs no pass of its own.  k < M makes target >= k when the sum does not wrap,
    // so with b = the smallest c >= k (it exists, M > k):  b < target -> the subset contains b, is feasible through b
    // alone and b is its answer as well;  b >= target -> no c of the subset reaches k and every c < k is below target,
    // i.e. the subset is exactly {c < k}.  A wrapped or tiny target (<= 1) empties the subset; the reference then uses
    // every node (:93), whose answer is b.
    T limit = ~(T)0;                         // the set is {0 < c <= limit}
    const unsigned long long total = sum_lt; // capacity of the set whenever the greedy part below runs (all its c < k)
    if (kT < M) {
        const int64_t target = (int64_t)((uint64_t)k + (uint64_t)M) / 2;
        if (target > 1 && (uint64_t)best_c >= (uint64_t)target && sum_lt >= k) { limit = kT - 1; best_p = -1; }
    }
    if (best_p >= 0) {                       // one node takes all k executors (:106-113)
        const int32_t node = slot_node[best_p];
        for (uint32_t j = lane; j < k; j += kWarp) out[j] = node;
        return;
    }
    const T U = warp_max_t(max_lt);          // largest c of the set

    // ---- every c of the set is < k: find v* = max{v : F(v) > k} -----------------------------------------------
    uint32_t vstar = 0;
    unsigned long long f_hi = k;             // F(v* + 1)
    if (total != k) {
        uint32_t lo = 1, hi = (uint32_t)U;   // invariant: F(lo) > k, F(hi + 1) <= k
        f_hi = 0;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo + 1) / 2;
            unsigned long long f = 0;
            for (int32_t p0 = 0; p0 < ne; p0 += kWarp) {
                const int32_t i = p0 + lane;
                const T c = i < ne ? capd(i) : (T)0;
                f += warp_sum((c >= (T)mid && c <= limit) ? (uint32_t)c : 0u);
            }
            st.nodes += (unsigned long long)ne;
            if (f > k) lo = mid; else { hi = mid - 1; f_hi = f; }
        }
        vstar = lo;
    }
    uint32_t r = k - (uint32_t)f_hi, m = 0;
    if (r > 0 && vstar < r) { m = r / vstar; r -= m * vstar; }

    // ---- collect the consumed nodes in priority order; pick the node of the remainder ------------------------------
    uint32_t L = 0, seen_star = 0;
    T fin_c = 0; int32_t fin_p = -1;
    for (int32_t p0 = 0; p0 < ne; p0 += kWarp) {
        const int32_t i = p0 + lane;
        const T c = i < ne ? capd(i) : (T)0;
        const bool inset = c > 0 && c <= limit;
        const bool star = inset && c == (T)vstar;
        const unsigned below = (1u << lane) - 1u;
        const unsigned sb = __ballot_sync(kFull, star);
        const bool consumed = inset && (c > (T)vstar || (star && seen_star + __popc(sb & below) < m));
        const unsigned cb = __ballot_sync(kFull, consumed);
        if (consumed) list[L + __popc(cb & below)] = make_int2(i, (int)c);
        else if (inset && r > 0 && c >= (T)r) note_min_key(c, i, fin_c, fin_p);
        L += __popc(cb);
        seen_star += __popc(sb);
    }
    st.nodes += (unsigned long long)ne;
    __syncwarp();

    // ---- emit: consumed nodes ranked by (c descending, order ascending), then the remainder ------------------------
    for (uint32_t t0 = 0; t0 < L; t0 += kWarp) {
        const uint32_t t = t0 + lane;
        const int2 mine = t < L ? list[t] : make_int2(0, 0);
        uint32_t off = 0;
        for (uint32_t u0 = 0; u0 < L; u0 += kWarp) {
            const uint32_t u = u0 + lane;
            const int cu = u < L ? list[u].y : 0;
            const uint32_t nu = min(L - u0, (uint32_t)kWarp);
            for (uint32_t w = 0; w < nu; ++w) {
                const int cw = __shfl_sync(kFull, cu, (int)w);
                if (cw > mine.y || (cw == mine.y && u0 + w < t)) off += (uint32_t)cw;
            }
        }
        const int32_t node = t < L ? slot_node[mine.x] : -1;
        const uint32_t nt = min(L - t0, (uint32_t)kWarp);
        for (uint32_t w = 0; w < nt; ++w) {
            const uint32_t ow = __shfl_sync(kFull, off, (int)w);
            const uint32_t cw = (uint32_t)__shfl_sync(kFull, mine.y, (int)w);
            const int32_t nw = __shfl_sync(kFull, node, (int)w);
            for (uint32_t j = lane; j < cw; j += kWarp) out[ow + j] = nw;
        }
    }
    if (r > 0) {
        warp_min_key(fin_c, fin_p);
        const int32_t node = slot_node[fin_p];
        for (uint32_t j = lane; j < r; j += kWarp) out[k - r + j] = node;
    }
}

// One application (GP_MODE_INDEPENDENT).  Returns the driver's node index or -1.
// MUT / grp_override / out_override: see pack_app_impl (gangpack_kernels.cuh)
template <bool FAST32, bool MUT = false>
__device__ __noinline__ int32_t pack_app_minfrag(const Snapshot& s, const PrepApp* __restrict__ pa,
                                                 int32_t* __restrict__ executor_nodes, int2* __restrict__ scratch,
                                                 WarpStats& st, int lane, int snap_flags,
                                                 int32_t grp_override = -1, int64_t out_override = -1) {
    static_assert(!(MUT && FAST32), "the compact view is read-only");
    Caps<false> a;
    a.init(pa, (pa->flags & kAppUsesGpu) || (snap_flags & kSnapGpuNegative));
    const bool ug = a.use_gpu;
    const GroupDesc g = s.groups[grp_override >= 0 ? grp_override : pa->group];
    const uint32_t k = a.k;
    const uint32_t lmax = (uint32_t)pa->lmax;
    const int32_t ne = g.ne;
    const int64_t out_off = out_override >= 0 ? out_override : pa->out_off;

    MfCapGeneralT<MUT> cg;
    cg.cpu = a.cpu; cg.mem = a.mem; cg.gpu = a.gpu; cg.ug = ug;
    cg.pair = s.pair + g.sbase; cg.gpuv = s.gpu + g.sbase;
    MfCapFast32 cf;
    if (FAST32) {
        cf.mc_lo = (uint32_t)a.cpu.magic; cf.mc_hi = (uint32_t)(a.cpu.magic >> 32);
        cf.mm_lo = (uint32_t)a.mem.magic; cf.mm_hi = (uint32_t)(a.mem.magic >> 32);
        cf.shc = a.cpu.sh - (uint32_t)s.meta->shift32[0];
        cf.shm = a.mem.sh - (uint32_t)s.meta->shift32[1];
        cf.pair32 = s.pair32 + g.sbase;
    }

    // ---- how many executors fit without a driver (same early exit as tightly-pack) ------------------------------
    unsigned long long P = 0;
    int32_t pos = 0;
    bool early = (k == 0);
    const unsigned long long need = (unsigned long long)k + lmax;
    while (!early && pos < ne) {
        const int32_t i = pos + lane;
        uint32_t c = 0;
        if (i < ne) {
            if (FAST32) { const uint32_t q = cf.cap0(i); c = q < k ? q : k; }
            else { const uint64_t q = cg.cap0(i); c = q < (uint64_t)k ? (uint32_t)q : k; }
        }
        P += warp_sum(c);
        pos += kWarp;
        early = P >= need;
    }
    st.nodes += (unsigned long long)(pos < ne ? pos : ne);
    const bool exact_total = !early;
    if (exact_total && P < k) return -1;

    // ---- first feasible driver candidate (binpack.go:67-85) ------------------------------------------------------------
    int32_t dslot = -1;
    for (int32_t j0 = 0; j0 < g.nd && dslot < 0; j0 += kWarp) {
        const int32_t j = j0 + lane;
        bool feasible = false;
        int32_t ls = -1;
        if (j < g.nd) {
            ls = s.drv_slot[g.dbase + j];
            feasible = driver_fits<MUT>(s, g.sbase + ls, a, ug);
            if (feasible && exact_total && ls < ne) {
                const uint32_t c0 = a.template cap0<MUT>(s, g.sbase + ls, ug);
                const uint32_t cdl = a.template cap<MUT>(s, g.sbase + ls, a.d_cpu, a.d_mem, a.d_gpu, ug);
                feasible = (P - c0 + cdl >= k);
            }
        }
        const unsigned vote = __ballot_sync(kFull, feasible);
        st.drivers += (unsigned long long)((g.nd - j0) < kWarp ? (g.nd - j0) : kWarp);
        if (vote) dslot = __shfl_sync(kFull, ls, __ffs(vote) - 1);
    }
    if (dslot < 0) return -1;
    const int32_t driver_node = s.slot_node[g.sbase + dslot];
    if (k == 0) return driver_node;

    const uint64_t cd = dslot < ne ? cg.cap(dslot, a.d_cpu, a.d_mem, a.d_gpu) : 0ull;
    int32_t* out = executor_nodes + out_off;
    int2* list = scratch + out_off;
    if (FAST32) minfrag_emit(cf, ne, dslot, (uint32_t)cd, k, s.slot_node + g.sbase, out, list, st, lane);
    else minfrag_emit(cg, ne, dslot, cd, k, s.slot_node + g.sbase, out, list, st, lane);
    return driver_node;
}

}  // namespace gp

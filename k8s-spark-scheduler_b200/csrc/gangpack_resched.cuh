// gangpack_resched.cuh -- node choice for ONE executor that has no usable reservation,
// batched: one warp per executor pod, every decision independent against the current snapshot.
//
// Reference (all under /root/reference; EXT = internal/extender, LIB = vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg):
//   rescheduleExecutor, first fit over the executor priority order       EXT/resource.go:657-662
//   rescheduleExecutorWithMinimalFragmentation                           EXT/resource.go:675-705
//   GetNodeCapacities (called with the overhead map as "reserved")       LIB/capacity/capacity.go:78-102, EXT/resource.go:682
//
// first fit:  the first node of the order with !executorResources.GreaterThan(available[node]).
// minimal fragmentation:  c(n) = capacity of node n for this executor shape with reserved(n) taken off; among c >= 1 the
//   reference keeps the candidate with the smallest key (hosts no executor of this application, c, position in the
//   order) -- its switch (:688-699) is exactly a running lexicographic minimum.  So: one scan for the minimum of
//   (c, position) over every node, one pass over the (short) list of nodes that already host executors of the
//   application; the second wins whenever it found a node.
#pragma once

#include "gangpack_kernels.cuh"
#include "gangpack_minfrag.cuh"

namespace gp {

// division parameters of one request dimension, computed in the kernel (one 64-bit division per dimension and executor)
__device__ __forceinline__ DimDiv make_dimdiv(int64_t e) {
    DimDiv dv;
    dv.e = e; dv.magic = 0; dv.sh = 0; dv.kind = kDivInf;
    if (e > 0) {
        const uint32_t sh = (uint32_t)(__ffsll((long long)e) - 1);
        const uint64_t odd = (uint64_t)e >> sh;
        dv.sh = sh;
        if (odd == 1) dv.kind = kDivShift;
        else if ((odd >> 32) == 0) { dv.kind = kDivMagic; dv.magic = udiv64(0xFFFFFFFFFFFFFFFFull, odd) + 1; }
        else { dv.kind = kDivSlow; dv.sh = 0; }
    }
    return dv;
}

struct ReschedIn {
    const int64_t* exe_cpu; const int64_t* exe_mem; const int64_t* exe_gpu;   // [n_execs]
    const int32_t* group;                                                     // [n_execs] or null
    const int64_t* res_cpu; const int64_t* res_mem; const int64_t* res_gpu;   // [n_nodes] "reserved" per node, or null
    const int64_t* host_off; const int32_t* host_nodes;                       // CSR: nodes hosting executors of the same app, or null
    const int32_t* node_slot;                                                 // [n_nodes] node -> slot (-1: not an executor candidate)
    int32_t n_execs;
};

template <bool MINFRAG>
__global__ void __launch_bounds__(256) gp_reschedule_kernel(Snapshot s, ReschedIn in, int32_t* __restrict__ node_out,
                                                      int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
    for (int32_t x = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); x < in.n_execs; x += warps) {
        const int64_t ec = in.exe_cpu[x], em = in.exe_mem[x], eg = in.exe_gpu ? in.exe_gpu[x] : 0;
        const int32_t grp = in.group ? in.group[x] : 0;
        if (grp < 0 || grp >= s.n_groups) {
            if (lane == 0) { atomicOr(err, kErrBadGroup); node_out[x] = -1; }
            continue;
        }
        const GroupDesc g = s.groups[grp];
        const longlong2* pair = s.pair + g.sbase;
        const int64_t* gpuv = s.gpu + g.sbase;
        const int32_t* slot_node = s.slot_node + g.sbase;
        int32_t chosen = -1;
        if (!MINFRAG) {
            for (int32_t p0 = 0; p0 < g.ne && chosen < 0; p0 += kWarp) {
                const int32_t i = p0 + lane;
                bool fits = false;
                if (i < g.ne) {
                    const longlong2 v = __ldg(pair + i);
                    fits = !(ec > v.x) && !(em > v.y) && !(eg > __ldg(gpuv + i));      // !GreaterThan, resources.go:239-241
                }
                const unsigned vote = __ballot_sync(kFull, fits);
                if (vote) chosen = p0 + __ffs(vote) - 1;
            }
        } else {
            const DimDiv dc = make_dimdiv(ec), dm = make_dimdiv(em), dg = make_dimdiv(eg);
            auto cap_at = [&](int32_t i) -> uint64_t {
                const longlong2 v = __ldg(pair + i);
                int64_t rc = 0, rm = 0, rg = 0;
                if (in.res_cpu) { const int32_t n = __ldg(slot_node + i); rc = in.res_cpu[n]; rm = in.res_mem[n]; rg = in.res_gpu ? in.res_gpu[n] : 0; }
                uint64_t c = min(cap_dim_u(v.x - rc, dc), cap_dim_u(v.y - rm, dm));
                return min(c, cap_dim_u(__ldg(gpuv + i) - rg, dg));
            };
            uint64_t best_c = 0; int32_t best_p = -1;
            for (int32_t p0 = 0; p0 < g.ne; p0 += kWarp) {
                const int32_t i = p0 + lane;
                if (i < g.ne) {
                    const uint64_t c = cap_at(i);
                    if (c >= 1) note_min_key(c, i, best_c, best_p);
                }
            }
            warp_min_key(best_c, best_p);
            if (in.host_off && best_p >= 0) {
                uint64_t hc = 0; int32_t hp = -1;
                for (int64_t t = in.host_off[x] + lane; t < in.host_off[x + 1]; t += kWarp) {
                    const int32_t n = in.host_nodes[t];
                    const int32_t local = in.node_slot[n] - g.sbase;
                    if (local >= 0 && local < g.ne) {              // an executor candidate of this group
                        const uint64_t c = cap_at(local);
                        if (c >= 1 && (hp < 0 || c < hc || (c == hc && local < hp))) { hc = c; hp = local; }
                    }
                }
                warp_min_key(hc, hp);
                if (hp >= 0) best_p = hp;
            }
            chosen = best_p;
        }
        if (lane == 0) node_out[x] = chosen >= 0 ? slot_node[chosen] : -1;
    }
}

}  // namespace gp

/*
 * oracle/gangpack_closed.c -- CLOSED-FORM CPU restatement (int64 SoA, index orders).
 * TEST INFRASTRUCTURE ONLY (see gangpack_oracle.h).  Parity status: partially pinned, see header.
 *
 * Same results as the literal restatement (tests/ cross-check them on randomized inputs), computed
 * from the per-node executor capacity -- the reference's own statement of which is
 * LIB/capacity/capacity.go:36-75 (GetNodeCapacity):
 *
 *   cap_dim(n | r) = 0                      if r_dim > avail_dim      (resources.go:239 on `reserved`)
 *                  = +INF                   if exe_dim == 0
 *                  = floor((avail-r)/exe)   otherwise
 *   cap(n | r)     = min over cpu, mem, gpu;   0 for a node that is not in the metadata
 *
 * tightly-pack  (LIB/binpack/pack_tightly.go:34-63): node n receives min(cap, remaining), node-major.
 * distribute-evenly (LIB/binpack/distribute_evenly.go:34-73): round r hands one executor to every
 *   node with cap >= r, in order, until `count` are placed.
 * driver loop (LIB/binpack/binpack.go:60-87): first d in driverOrder with !gt(drv, avail[d]) whose
 *   executor total  S0 - min(cap(d|0),k) + min(cap(d|drv),k)  reaches k.
 * minimal-fragmentation (LIB/binpack/minimal_fragmentation.go:59-137): see minfrag_emit().
 *
 * Precondition (checked by the product API too): exec_order / driver_order hold no duplicates.
 */
#include "gangpack_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define CAP_INF INT64_MAX

static inline int64_t cap_dim(int64_t avail, int64_t r, int64_t e) {
    if (r > avail) return 0;
    if (e == 0) return CAP_INF;
    return (avail - r) / e;
}
static inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

typedef struct {
    int32_t n_nodes;
    int64_t *cpu, *mem, *gpu;
    const int32_t* driver_order; int32_t n_driver;
    const int32_t* exec_order; int32_t n_exec;
    const uint8_t* in_exec; /* [n_nodes] */
} snap;

static inline int valid(const snap* s, int32_t n) { return n >= 0 && n < s->n_nodes; }

static inline int64_t node_cap(const snap* s, int32_t n, const orc_res* r, const orc_res* e) {
    if (!valid(s, n)) return 0;
    int64_t c = cap_dim(s->cpu[n], r->cpu, e->cpu);
    c = min64(c, cap_dim(s->mem[n], r->mem, e->mem));
    c = min64(c, cap_dim(s->gpu[n], r->gpu, e->gpu));
    return c;
}

static void minfrag_emit(const snap* s, int32_t d, int64_t cd, const orc_res* exe, int32_t k, int32_t* out);

/* returns driver node or -1; writes executor nodes (count entries) on success */
static int32_t pack_one(const snap* s, int algo, const orc_res* drv, const orc_res* exe, int32_t k,
                        int32_t* executor_nodes) {
    static const orc_res zero = {0, 0, 0};
    int64_t S0 = 0;
    for (int32_t i = 0; i < s->n_exec; ++i) {
        S0 += min64(node_cap(s, s->exec_order[i], &zero, exe), k);
    }
    int32_t d = -1;
    int64_t cd = 0;
    for (int32_t i = 0; i < s->n_driver; ++i) {
        int32_t c = s->driver_order[i];
        if (!valid(s, c)) continue;                                                       /* binpack.go:68-69 */
        if (drv->cpu > s->cpu[c] || drv->mem > s->mem[c] || drv->gpu > s->gpu[c]) continue; /* :69 */
        int64_t Sd = S0;
        int64_t capd = 0;
        if (s->in_exec[c]) {
            capd = node_cap(s, c, drv, exe);
            Sd = S0 - min64(node_cap(s, c, &zero, exe), k) + min64(capd, k);
        }
        if (Sd >= k) { d = c; cd = capd; break; }
    }
    if (d < 0) return -1;
    int32_t placed = 0;
    if (k == 0) return d;
    if (algo == ORC_MINIMAL_FRAGMENTATION) {
        minfrag_emit(s, s->in_exec[d] ? d : -1, cd, exe, k, executor_nodes);
    } else if (algo == ORC_TIGHTLY_PACK) {
        for (int32_t i = 0; i < s->n_exec && placed < k; ++i) {
            int32_t n = s->exec_order[i];
            int64_t c = (n == d) ? cd : node_cap(s, n, &zero, exe);
            int64_t take = min64(c, (int64_t)(k - placed));
            for (int64_t t = 0; t < take; ++t) executor_nodes[placed++] = n;
        }
    } else {
        for (int64_t r = 1; placed < k; ++r) {
            for (int32_t i = 0; i < s->n_exec && placed < k; ++i) {
                int32_t n = s->exec_order[i];
                int64_t c = (n == d) ? cd : node_cap(s, n, &zero, exe);
                if (c >= r) executor_nodes[placed++] = n;
            }
        }
    }
    return d;
}

/* minimal fragmentation (LIB/binpack/minimal_fragmentation.go:59-137) in closed form, for driver node d with
 * cap(d|drv) = cd.  c(i) = UNCLAMPED capacity of the i-th executor candidate, nodes with c = 0 are filtered out.
 *   M = max c.  If k < M: target = wrap64(k + M) / 2 and the subset S = {c < target} is tried first; it works iff
 *   sum_S c >= k; otherwise (and when k >= M) the set is every node.  Within the chosen set:
 *   - some c >= k: k executors on the node with the smallest such c (ties: earliest in the order);
 *   - else nodes are consumed whole in (c descending, order ascending) while the remainder r >= c.  With
 *     F(v) = sum of c over {c >= v}:  v* = max{v : F(v) > k}, every node with c > v* is consumed,
 *     r* = k - F(v*+1); if v* < r*, the first m = floor(r* / v*) nodes with c == v* are consumed too; the rest
 *     r' = r* - m v* (if any) goes to the node with the smallest c >= r' among the unconsumed ones (ties: earliest).
 * The feasibility of a driver candidate is the same as for tightly-pack (sum c >= k), so the driver loop above
 * is shared. */
static void minfrag_emit(const snap* s, int32_t d, int64_t cd, const orc_res* exe, int32_t k, int32_t* out) {
    static const orc_res zero = {0, 0, 0};
    const int32_t ne = s->n_exec;
    int64_t* c = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ne > 0 ? ne : 1));
    int64_t M = 0;
    for (int32_t i = 0; i < ne; ++i) {
        int32_t n = s->exec_order[i];
        c[i] = (n == d) ? cd : node_cap(s, n, &zero, exe);
        if (c[i] > M) M = c[i];
    }
    /* The subset {c < target} of minimal_fragmentation.go:80-89 needs no pass of its own.  k < M makes
     * target = (k + M) / 2 >= k when the sum does not wrap, so with b = the smallest c >= k (it exists, M > k):
     *   b < target  -> the subset contains b, is feasible through b alone, and b is its answer as well;
     *   b >= target -> no c of the subset reaches k and every c < k is below target: the subset is exactly {c < k}.
     * A wrapped or tiny target (<= 1) empties the subset; the reference then uses every node, whose answer is b. */
    int64_t limit = CAP_INF;      /* the set is {0 < c <= limit} */
    int32_t best = -1;
    int64_t sum_lt = 0;
    for (int32_t i = 0; i < ne; ++i) {
        if (c[i] >= k && (best < 0 || c[i] < c[best])) best = i;
        if (c[i] > 0 && c[i] < k) sum_lt += c[i];
    }
    if ((int64_t)k < M) {
        int64_t target = (int64_t)((uint64_t)k + (uint64_t)M) / 2;   /* Go int wraps; / truncates toward zero */
        if (target > 1 && c[best] >= target && sum_lt >= k) { limit = (int64_t)k - 1; best = -1; }
    }
    if (best >= 0) {
        for (int32_t t = 0; t < k; ++t) out[t] = s->exec_order[best];
        free(c);
        return;
    }
    /* every c in the set is < k now */
    int64_t U = 0, total = 0;
    for (int32_t i = 0; i < ne; ++i) if (c[i] > 0 && c[i] <= limit) { total += c[i]; if (c[i] > U) U = c[i]; }
    int64_t vstar = 0, Fhi = 0;   /* F(vstar + 1) */
    if (total == k) { vstar = 0; Fhi = k; }
    else {
        int64_t lo = 1, hi = U;   /* F(lo) > k, F(hi + 1) <= k */
        while (lo < hi) {
            int64_t mid = lo + (hi - lo + 1) / 2, f = 0;
            for (int32_t i = 0; i < ne; ++i) if (c[i] >= mid && c[i] <= limit) f += c[i];
            if (f > k) lo = mid; else { hi = mid - 1; Fhi = f; }
        }
        vstar = lo;
    }
    int64_t r = k - Fhi, m = 0;
    if (r > 0 && vstar < r) { m = r / vstar; r -= m * vstar; }
    /* consumed nodes: offset = sum of larger capacities + c * (#equal capacities earlier in the order) */
    int64_t seen_star = 0;
    int32_t fin = -1;
    for (int32_t i = 0; i < ne; ++i) {
        if (c[i] <= 0 || c[i] > limit) continue;
        int consumed = c[i] > vstar;
        if (c[i] == vstar) { consumed = seen_star < m; ++seen_star; }
        if (consumed) {
            int64_t off = 0;
            for (int32_t j = 0; j < ne; ++j) {
                if (c[j] <= 0 || c[j] > limit) continue;
                if (c[j] > c[i] || (c[j] == c[i] && j < i)) off += c[j];
            }
            for (int64_t t = 0; t < c[i]; ++t) out[off + t] = s->exec_order[i];
        } else if (r > 0 && c[i] >= r && (fin < 0 || c[i] < c[fin])) {
            fin = i;
        }
    }
    for (int64_t t = 0; t < r; ++t) out[k - r + t] = s->exec_order[fin];
    free(c);
}

typedef struct {
    const snap* s; int algo; int32_t lo, hi;
    const orc_res *drv, *exe; const int32_t* count;
    const int64_t* exec_off; int32_t* driver_node; int32_t* executor_nodes;
} job;

static void* worker(void* p) {
    job* j = (job*)p;
    for (int32_t i = j->lo; i < j->hi; ++i)
        j->driver_node[i] = pack_one(j->s, j->algo, &j->drv[i], &j->exe[i], j->count[i],
                                     j->executor_nodes + j->exec_off[i]);
    return NULL;
}

int32_t orc_closed_batch(int algo, int mode, int32_t n_nodes,
                         int64_t* avail_cpu, int64_t* avail_mem, int64_t* avail_gpu,
                         const int32_t* driver_order, int32_t n_driver,
                         const int32_t* exec_order, int32_t n_exec,
                         int32_t n_apps, const orc_res* drv, const orc_res* exe,
                         const int32_t* count, const uint8_t* young, int n_threads,
                         const int64_t* exec_off, int32_t* driver_node, int32_t* executor_nodes) {
    uint8_t* in_exec = (uint8_t*)calloc((size_t)(n_nodes > 0 ? n_nodes : 1), 1);
    for (int32_t i = 0; i < n_exec; ++i)
        if (exec_order[i] >= 0 && exec_order[i] < n_nodes) in_exec[exec_order[i]] = 1;
    snap s = {n_nodes, avail_cpu, avail_mem, avail_gpu, driver_order, n_driver, exec_order, n_exec, in_exec};
    int32_t blocked = -1;

    if (mode == 0) {
        if (n_threads < 1) n_threads = 1;
        if (n_threads > n_apps) n_threads = n_apps > 0 ? n_apps : 1;
        job* jobs = (job*)malloc(sizeof(job) * (size_t)n_threads);
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
        for (int t = 0; t < n_threads; ++t) {
            job j = {&s, algo, (int32_t)((int64_t)n_apps * t / n_threads), (int32_t)((int64_t)n_apps * (t + 1) / n_threads),
                     drv, exe, count, exec_off, driver_node, executor_nodes};
            jobs[t] = j;
            if (n_threads == 1) worker(&jobs[t]); else pthread_create(&th[t], NULL, worker, &jobs[t]);
        }
        if (n_threads > 1) for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(jobs); free(th);
    } else {
        /* fitEarlierDrivers, EXT/resource.go:224-262 */
        uint8_t* seen = (uint8_t*)calloc((size_t)(n_nodes > 0 ? n_nodes : 1), 1);
        int32_t i = 0;
        for (; i < n_apps; ++i) {
            int32_t* en = executor_nodes + exec_off[i];
            int32_t d = pack_one(&s, algo, &drv[i], &exe[i], count[i], en);
            driver_node[i] = d;
            if (d < 0) {
                if (young && young[i]) continue;      /* :245-249 */
                blocked = i; ++i; break;              /* :250-252 */
            }
            if (mode == ORC_FIFO_REFERENCE) {
                /* EXT/sparkpods.go:139-146: usage[driver]=drv, then usage[n]=exe for each executor
                 * node (assignment) -> each distinct executor node charged once; a driver node that
                 * also hosts an executor is charged the executor only. */
                int driver_hosts_executor = 0;
                for (int32_t k = 0; k < count[i]; ++k) {
                    int32_t n = en[k];
                    if (n == d) driver_hosts_executor = 1;
                    if (!seen[n]) {
                        seen[n] = 1;
                        avail_cpu[n] -= exe[i].cpu; avail_mem[n] -= exe[i].mem; avail_gpu[n] -= exe[i].gpu;
                    }
                }
                for (int32_t k = 0; k < count[i]; ++k) seen[en[k]] = 0;
                if (!driver_hosts_executor) {
                    avail_cpu[d] -= drv[i].cpu; avail_mem[d] -= drv[i].mem; avail_gpu[d] -= drv[i].gpu;
                }
            } else {
                avail_cpu[d] -= drv[i].cpu; avail_mem[d] -= drv[i].mem; avail_gpu[d] -= drv[i].gpu;
                for (int32_t k = 0; k < count[i]; ++k) {
                    int32_t n = en[k];
                    avail_cpu[n] -= exe[i].cpu; avail_mem[n] -= exe[i].mem; avail_gpu[n] -= exe[i].gpu;
                }
            }
        }
        for (; i < n_apps; ++i) driver_node[i] = -2;
        free(seen);
    }
    free(in_exec);
    return blocked;
}

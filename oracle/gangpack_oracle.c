/*
 * oracle/gangpack_oracle.c -- LITERAL CPU restatement of the reference placement hot path.
 * TEST INFRASTRUCTURE ONLY (see gangpack_oracle.h).  Parity status: partially pinned, see header.
 *
 * Written loop-for-loop from the cited reference lines, keeping the reference's data structures
 * (string-keyed hash maps, a fresh `reserved` map per driver candidate, an N-entry
 * `availableNodes` set per distribute-evenly call) so that it can also stand in for the cost
 * profile of the Go path when timed as the CPU baseline ("restated, not Go").
 */
#include "gangpack_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Allocation model.  The Go code allocates a fresh map per driver candidate / per call from the GC
 * heap (zeroed spans, no syscalls on the hot path).  glibc would serve these several-hundred-KB
 * blocks with mmap/munmap or trim the thread arenas after every free, so worker threads serialise
 * on the kernel's mm lock and the multi-threaded CPU baseline would look far worse than a Go
 * runtime.  Stand-in: a small thread-local cache of freed blocks; every "allocation" still pays
 * the zeroing a fresh Go map pays. */
#define POOL_SLOTS 24
typedef struct { void* p; size_t size; } pool_slot;
static __thread pool_slot g_pool[POOL_SLOTS];

static void* pool_alloc(size_t size, int zero) {
    for (int i = 0; i < POOL_SLOTS; ++i) {
        if (g_pool[i].p && g_pool[i].size == size) {
            void* p = g_pool[i].p;
            g_pool[i].p = NULL;
            if (zero) memset(p, 0, size);
            return p;
        }
    }
    return zero ? calloc(1, size) : malloc(size);
}
static void pool_free(void* p, size_t size) {
    if (!p) return;
    for (int i = 0; i < POOL_SLOTS; ++i) {
        if (!g_pool[i].p) { g_pool[i].p = p; g_pool[i].size = size; return; }
    }
    free(p);
}
static void tune_allocator(void) {}

/* ------------------------------------------------------------------ Resources ---- */
/* LIB/resources/resources.go:202-206 */
static inline void res_add(orc_res* r, const orc_res* o) { r->cpu += o->cpu; r->mem += o->mem; r->gpu += o->gpu; }
/* LIB/resources/resources.go:209-213 */
static inline void res_sub(orc_res* r, const orc_res* o) { r->cpu -= o->cpu; r->mem -= o->mem; r->gpu -= o->gpu; }
/* LIB/resources/resources.go:239-241: any dimension strictly greater */
static inline int res_greater_than(const orc_res* r, const orc_res* o) {
    return r->cpu > o->cpu || r->mem > o->mem || r->gpu > o->gpu;
}
/* LIB/resources/resources.go:244-246 */
static inline int res_eq(const orc_res* r, const orc_res* o) {
    return r->cpu == o->cpu && r->mem == o->mem && r->gpu == o->gpu;
}

/* ------------------------------------------------------------------ string-keyed map ---- */
/* Stand-in for Go's map[string]T: open addressing, FNV-1a, tombstones, value = int32 slot. */
typedef struct {
    const char** keys; /* NULL = empty, TOMB = deleted */
    int32_t* vals;
    uint32_t cap;      /* power of two */
    uint32_t len;
    uint32_t used;     /* len + tombstones */
} strmap;

static const char TOMB_OBJ = 0;
#define TOMB (&TOMB_OBJ)

static uint64_t fnv1a(const char* s) {
    uint64_t h = 1469598103934665603ull;
    for (; *s; ++s) { h ^= (unsigned char)*s; h *= 1099511628211ull; }
    return h;
}
static uint32_t pow2_at_least(uint32_t x) { uint32_t c = 8; while (c < x) c <<= 1; return c; }

static void strmap_init(strmap* m, uint32_t expected) {
    m->cap = pow2_at_least(expected * 2 + 2);
    m->keys = (const char**)pool_alloc(m->cap * sizeof(char*), 1);
    m->vals = (int32_t*)pool_alloc(m->cap * sizeof(int32_t), 0);
    m->len = 0; m->used = 0;
}
static void strmap_free(strmap* m) {
    pool_free((void*)m->keys, m->cap * sizeof(char*)); pool_free(m->vals, m->cap * sizeof(int32_t));
    m->keys = NULL; m->vals = NULL;
}

static int32_t* strmap_find(const strmap* m, const char* k) {
    uint32_t i = (uint32_t)fnv1a(k) & (m->cap - 1);
    for (;;) {
        const char* e = m->keys[i];
        if (e == NULL) return NULL;
        if (e != TOMB && strcmp(e, k) == 0) return &m->vals[i];
        i = (i + 1) & (m->cap - 1);
    }
}
static void strmap_grow(strmap* m);
static int32_t* strmap_insert(strmap* m, const char* k, int32_t v) { /* k assumed absent */
    if ((m->used + 1) * 4 > m->cap * 3) strmap_grow(m);
    uint32_t i = (uint32_t)fnv1a(k) & (m->cap - 1);
    while (m->keys[i] != NULL && m->keys[i] != TOMB) i = (i + 1) & (m->cap - 1);
    if (m->keys[i] == NULL) m->used++;
    m->keys[i] = k; m->vals[i] = v; m->len++;
    return &m->vals[i];
}
static void strmap_grow(strmap* m) {
    strmap n;
    n.cap = m->cap * 2;
    n.keys = (const char**)pool_alloc(n.cap * sizeof(char*), 1);
    n.vals = (int32_t*)pool_alloc(n.cap * sizeof(int32_t), 0);
    n.len = 0; n.used = 0;
    for (uint32_t i = 0; i < m->cap; ++i)
        if (m->keys[i] != NULL && m->keys[i] != TOMB) strmap_insert(&n, m->keys[i], m->vals[i]);
    strmap_free(m);
    *m = n;
}
static void strmap_delete(strmap* m, const char* k) {
    uint32_t i = (uint32_t)fnv1a(k) & (m->cap - 1);
    for (;;) {
        const char* e = m->keys[i];
        if (e == NULL) return;
        if (e != TOMB && strcmp(e, k) == 0) { m->keys[i] = TOMB; m->len--; return; }
        i = (i + 1) & (m->cap - 1);
    }
}

/* ------------------------------------------------------------------ cluster ---- */
typedef struct {
    orc_res available;    /* AvailableResources   */
    orc_res schedulable;  /* SchedulableResources */
    const char* zone;     /* ZoneLabel            */
    uint8_t unschedulable, ready;
} node_meta;

struct orc_cluster {
    int32_t n;
    char** names;     /* owned copies, insertion order */
    char** zones;     /* owned copies */
    node_meta* meta;
    strmap index;     /* name -> slot */
};

static char* dupstr(const char* s) {
    size_t l = strlen(s) + 1; char* d = (char*)malloc(l); memcpy(d, s, l); return d;
}

orc_cluster* orc_cluster_new(int32_t n, const char* const* names,
                             const int64_t* avail_cpu, const int64_t* avail_mem, const int64_t* avail_gpu,
                             const int64_t* sched_cpu, const int64_t* sched_mem, const int64_t* sched_gpu,
                             const char* const* zone, const uint8_t* unschedulable, const uint8_t* ready) {
    orc_cluster* c = (orc_cluster*)calloc(1, sizeof(*c));
    c->n = n;
    c->names = (char**)malloc(sizeof(char*) * (size_t)(n > 0 ? n : 1));
    c->zones = (char**)malloc(sizeof(char*) * (size_t)(n > 0 ? n : 1));
    c->meta = (node_meta*)calloc((size_t)(n > 0 ? n : 1), sizeof(node_meta));
    strmap_init(&c->index, (uint32_t)n);
    for (int32_t i = 0; i < n; ++i) {
        c->names[i] = dupstr(names[i]);
        /* LIB/resources/resources.go:78-81: missing zone label -> "default" */
        c->zones[i] = dupstr(zone ? zone[i] : "default");
        node_meta* m = &c->meta[i];
        m->available.cpu = avail_cpu[i]; m->available.mem = avail_mem[i]; m->available.gpu = avail_gpu ? avail_gpu[i] : 0;
        m->schedulable.cpu = sched_cpu ? sched_cpu[i] : m->available.cpu;
        m->schedulable.mem = sched_mem ? sched_mem[i] : m->available.mem;
        m->schedulable.gpu = sched_gpu ? sched_gpu[i] : m->available.gpu;
        m->zone = c->zones[i];
        m->unschedulable = unschedulable ? unschedulable[i] : 0;
        m->ready = ready ? ready[i] : 1;
        strmap_insert(&c->index, c->names[i], i);
    }
    return c;
}
void orc_cluster_free(orc_cluster* c) {
    if (!c) return;
    for (int32_t i = 0; i < c->n; ++i) { free(c->names[i]); free(c->zones[i]); }
    free(c->names); free(c->zones); free(c->meta); strmap_free(&c->index); free(c);
}
int32_t orc_cluster_size(const orc_cluster* c) { return c->n; }
int32_t orc_cluster_index(const orc_cluster* c, const char* name) {
    int32_t* s = strmap_find(&c->index, name); return s ? *s : -1;
}
void orc_cluster_get_available(const orc_cluster* c, int64_t* cpu, int64_t* mem, int64_t* gpu) {
    for (int32_t i = 0; i < c->n; ++i) {
        cpu[i] = c->meta[i].available.cpu; mem[i] = c->meta[i].available.mem; gpu[i] = c->meta[i].available.gpu;
    }
}
/* nodesSchedulingMetadata[name] with the ", ok" form */
static inline node_meta* meta_lookup(const orc_cluster* c, const char* name) {
    int32_t* s = strmap_find(&c->index, name); return s ? &c->meta[*s] : NULL;
}

/* ------------------------------------------------------------------ reserved map ---- */
/* resources.NodeGroupResources: map[string]*Resources (LIB/resources/resources.go:103) */
typedef struct { strmap idx; orc_res* vals; int32_t n, cap; } resmap;

static void resmap_init(resmap* r, int32_t expected) {
    strmap_init(&r->idx, (uint32_t)expected);
    r->cap = expected > 4 ? expected : 4;
    r->vals = (orc_res*)pool_alloc(sizeof(orc_res) * (size_t)r->cap, 0);
    r->n = 0;
}
static void resmap_free(resmap* r) { strmap_free(&r->idx); pool_free(r->vals, sizeof(orc_res) * (size_t)r->cap); }
static orc_res* resmap_get(resmap* r, const char* k) {
    int32_t* s = strmap_find(&r->idx, k); return s ? &r->vals[*s] : NULL;
}
static orc_res* resmap_put(resmap* r, const char* k, const orc_res* v) {
    orc_res* e = resmap_get(r, k);
    if (e) { *e = *v; return e; }
    if (r->n == r->cap) {
        orc_res* nv = (orc_res*)pool_alloc(sizeof(orc_res) * (size_t)r->cap * 2, 0);
        memcpy(nv, r->vals, sizeof(orc_res) * (size_t)r->n);
        pool_free(r->vals, sizeof(orc_res) * (size_t)r->cap);
        r->vals = nv; r->cap *= 2;
    }
    r->vals[r->n] = *v;
    strmap_insert(&r->idx, k, r->n);
    return &r->vals[r->n++];
}

/* ------------------------------------------------------------------ executor distributors ---- */
typedef int (*distribute_fn)(const orc_cluster*, const orc_res*, int32_t, const char* const*, int32_t,
                             resmap*, const char**);

/* tightlyPackExecutors, LIB/binpack/pack_tightly.go:34-63 */
static int tightly_pack_executors(const orc_cluster* c, const orc_res* exe, int32_t count,
                                  const char* const* order, int32_t n_order,
                                  resmap* reserved, const char** executor_nodes) {
    static const orc_res zero = {0, 0, 0};
    int32_t placed = 0;
    if (count == 0) return 1;                                       /* :42-44 */
    for (int32_t i = 0; i < n_order; ++i) {                         /* :45 */
        const char* n = order[i];
        orc_res* r = resmap_get(reserved, n);                       /* :46-48 */
        if (r == NULL) r = resmap_put(reserved, n, &zero);
        for (;;) {                                                  /* :49 */
            res_add(r, exe);                                        /* :50 */
            const node_meta* m = meta_lookup(c, n);                 /* :51 */
            if (m == NULL || res_greater_than(r, &m->available)) {  /* :52 */
                res_sub(r, exe);                                    /* :53 */
                break;                                              /* :54 */
            }
            executor_nodes[placed++] = n;                           /* :56 */
            if (placed == count) return 1;                          /* :57-59 */
        }
    }
    return 0;                                                       /* :62 */
}

/* distributeExecutorsEvenly, LIB/binpack/distribute_evenly.go:34-73 */
static int distribute_executors_evenly(const orc_cluster* c, const orc_res* exe, int32_t count,
                                       const char* const* order, int32_t n_order,
                                       resmap* reserved, const char** executor_nodes) {
    static const orc_res zero = {0, 0, 0};
    strmap available;                                               /* :41-44 */
    strmap_init(&available, (uint32_t)n_order);
    for (int32_t i = 0; i < n_order; ++i)
        if (strmap_find(&available, order[i]) == NULL) strmap_insert(&available, order[i], 1);
    int32_t placed = 0;
    if (count == 0) { strmap_free(&available); return 1; }          /* :46-48 */
    while (available.len > 0) {                                     /* :49 */
        for (int32_t i = 0; i < n_order; ++i) {                     /* :50 */
            const char* n = order[i];
            if (strmap_find(&available, n) == NULL) continue;       /* :51-53 */
            orc_res* r = resmap_get(reserved, n);                   /* :55-57 */
            if (r == NULL) r = resmap_put(reserved, n, &zero);
            res_add(r, exe);                                        /* :58 */
            const node_meta* m = meta_lookup(c, n);                 /* :59 */
            if (m == NULL || res_greater_than(r, &m->available)) {  /* :60 */
                strmap_delete(&available, n);                       /* :62 */
                res_sub(r, exe);                                    /* :63 */
            } else {
                executor_nodes[placed++] = n;                       /* :65 */
                if (placed == count) { strmap_free(&available); return 1; } /* :66-68 */
            }
        }
    }
    strmap_free(&available);
    return 0;                                                       /* :72 */
}

/* ---- minimal fragmentation ------------------------------------------------------ */
/* capacity.NodeAndExecutorCapacity, LIB/capacity/capacity.go:26-29 */
typedef struct { const char* name; int64_t capacity; } node_cap;

/* getCapacityAgainstSingleDimension, LIB/capacity/capacity.go:36-55 */
static int64_t capacity_single_dimension(int64_t available, int64_t reserved, int64_t required) {
    if (reserved > available) return 0;                                 /* :37-40 */
    if (required == 0) return INT64_MAX;                                /* :42-45 math.MaxInt */
    return (available - reserved) / required;                           /* :47-54 floor; numerator >= 0, required > 0 */
}
/* GetNodeCapacity + min, capacity.go:57-75,115-122 */
static int64_t node_capacity(const orc_res* available, const orc_res* reserved, const orc_res* single) {
    int64_t a = capacity_single_dimension(available->cpu, reserved->cpu, single->cpu);
    int64_t b = capacity_single_dimension(available->mem, reserved->mem, single->mem);
    int64_t g = capacity_single_dimension(available->gpu, reserved->gpu, single->gpu);
    if (a <= b && a <= g) return a;
    if (b <= g) return b;
    return g;
}
/* sort.SliceStable by Capacity ascending (minimal_fragmentation.go:76-78) */
static void msort_caps(node_cap* a, node_cap* tmp, int32_t n) {
    if (n < 2) return;
    int32_t h = n / 2;
    msort_caps(a, tmp, h); msort_caps(a + h, tmp, n - h);
    int32_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].capacity < a[i].capacity) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(node_cap) * (size_t)n);
}
/* sort.Search(len, func(i) caps[i].Capacity >= x) on the ascending slice */
static int32_t search_at_least(const node_cap* a, int32_t n, int64_t x) {
    int32_t lo = 0, hi = n;
    while (lo < hi) { int32_t mid = lo + (hi - lo) / 2; if (a[mid].capacity >= x) hi = mid; else lo = mid + 1; }
    return lo;
}
/* internalMinimalFragmentation, minimal_fragmentation.go:96-137.  `caps` is copied like the reference does. */
static int internal_minimal_fragmentation(int64_t count, const node_cap* caps_in, int32_t n,
                                          const char** executor_nodes) {
    node_cap* caps = (node_cap*)malloc(sizeof(node_cap) * (size_t)(n > 0 ? n : 1));      /* :99-100 */
    memcpy(caps, caps_in, sizeof(node_cap) * (size_t)n);
    int64_t placed = 0;
    while (n > 0) {                                                                       /* :104 */
        int32_t position = search_at_least(caps, n, count);                               /* :106-108 */
        if (position != n) {                                                              /* :110-113 */
            for (int64_t t = 0; t < count; ++t) executor_nodes[placed++] = caps[position].name;
            free(caps);
            return 1;
        }
        int64_t max_capacity = caps[n - 1].capacity;                                      /* :116 */
        int32_t first_max = search_at_least(caps, n, max_capacity);                       /* :117-119 */
        int32_t pos = first_max;                                                          /* :122 */
        for (; count >= max_capacity && pos < n; ++pos) {                                 /* :123 */
            for (int64_t t = 0; t < max_capacity; ++t) executor_nodes[placed++] = caps[pos].name; /* :125 */
            count -= max_capacity;                                                        /* :126 */
        }
        if (count == 0) { free(caps); return 1; }                                         /* :129-131 */
        memmove(caps + first_max, caps + pos, sizeof(node_cap) * (size_t)(n - pos));      /* :133 */
        n -= pos - first_max;
    }
    free(caps);
    return 0;                                                                             /* :136 */
}
/* minimalFragmentation, LIB/binpack/minimal_fragmentation.go:59-94.  Note: it never adds the executors to
 * `reserved` (only the driver's entry exists), so the efficiencies SparkBinPack computes afterwards
 * (binpack.go:77) see the driver only -- kept as is. */
static int minimal_fragmentation(const orc_cluster* c, const orc_res* exe, int32_t count,
                                 const char* const* order, int32_t n_order,
                                 resmap* reserved, const char** executor_nodes) {
    static const orc_res zero = {0, 0, 0};
    if (count == 0) return 1;                                                             /* :66-68 */
    node_cap* caps = (node_cap*)malloc(sizeof(node_cap) * (size_t)(n_order > 0 ? n_order : 1) * 2);
    node_cap* tmp = caps + (n_order > 0 ? n_order : 1);
    int32_t n = 0;
    for (int32_t i = 0; i < n_order; ++i) {                                               /* GetNodeCapacities, capacity.go:78-102 */
        const node_meta* m = meta_lookup(c, order[i]);
        if (m == NULL) continue;                                                          /* :87 */
        const orc_res* r = resmap_get(reserved, order[i]);                                /* :88-92 */
        if (r == NULL) r = &zero;
        int64_t cap = node_capacity(&m->available, r, exe);
        if (cap > 0) { caps[n].name = order[i]; caps[n].capacity = cap; ++n; }            /* FilterOutNodesWithoutCapacity, :105-113 */
    }
    if (n == 0) { free(caps); return 0; }                                                 /* :72-74 */
    msort_caps(caps, tmp, n);                                                             /* :76-78 */
    int64_t max_capacity = caps[n - 1].capacity;                                          /* :79 */
    if ((int64_t)count < max_capacity) {                                                  /* :80 */
        /* Go int arithmetic wraps: count + MaxInt is negative, and / truncates toward zero */
        int64_t target = (int64_t)((uint64_t)count + (uint64_t)max_capacity) / 2;         /* :81 */
        int32_t first = search_at_least(caps, n, target);                                 /* :82-84 */
        if (internal_minimal_fragmentation(count, caps, first, executor_nodes)) {         /* :87-89 */
            free(caps);
            return 1;
        }
    }
    int ok = internal_minimal_fragmentation(count, caps, n, executor_nodes);              /* :93 */
    free(caps);
    return ok;
}

/* ------------------------------------------------------------------ efficiencies ---- */
/* Quantity.Value() for a milli-scaled CPU (quantity.go:732-734 -> int64Amount.AsScaledInt64 ->
 * negativeScaleInt64, math.go:166-199): whole cores, inexact values rounded AWAY from zero.
 * Memory / GPU are scale 0. */
static inline int64_t cpu_value(int64_t milli) {
    int64_t q = milli / 1000, rem = milli % 1000;
    if (rem > 0) return q + 1;
    if (rem < 0) return q - 1;
    return q;
}
static inline int64_t normalize_resource(int64_t v) { return v == 0 ? 1 : v; } /* efficiency.go:104-109 */

typedef struct { double cpu, mem, gpu; } pack_eff;

/* computePackingEfficiency, LIB/binpack/efficiency.go:79-102 */
static pack_eff compute_packing_efficiency(const node_meta* m, const orc_res* reserved /* may be NULL */) {
    orc_res r = m->schedulable;               /* :84 */
    res_sub(&r, &m->available);               /* :85 */
    if (reserved) res_add(&r, reserved);      /* :86-88 */
    pack_eff e;
    e.gpu = 0.0;                              /* :92-95 */
    if (m->schedulable.gpu != 0) e.gpu = (double)r.gpu / (double)normalize_resource(m->schedulable.gpu);
    e.cpu = (double)cpu_value(r.cpu) / (double)normalize_resource(cpu_value(m->schedulable.cpu));
    e.mem = (double)r.mem / (double)normalize_resource(m->schedulable.mem);
    return e;
}

/* ComputePackingEfficiencies (efficiency.go:66-77) followed by
 * computeAvgPackingEfficiencyForResult (EXT/resource.go:372-381) -> ComputeAvgPackingEfficiency
 * (efficiency.go:114-156).  Iteration: cluster insertion order (Go: map order). */
static void avg_packing_efficiency(const orc_cluster* c, resmap* reserved, double* out4) {
    size_t effs_bytes = sizeof(pack_eff) * (size_t)(c->n > 0 ? c->n : 1);
    pack_eff* effs = (pack_eff*)pool_alloc(effs_bytes, 0); /* one per node, :72-74 */
    for (int32_t i = 0; i < c->n; ++i)
        effs[i] = compute_packing_efficiency(&c->meta[i], resmap_get(reserved, c->names[i]));
    double cpu = 0, mem = 0, gpu = 0, mx = 0; int32_t with_gpu = 0;
    for (int32_t i = 0; i < c->n; ++i) {
        cpu += effs[i].cpu; mem += effs[i].mem;
        if (c->meta[i].schedulable.gpu != 0) { gpu += effs[i].gpu; with_gpu++; }
        double m2 = effs[i].cpu > effs[i].mem ? effs[i].cpu : effs[i].mem;
        mx += effs[i].gpu > m2 ? effs[i].gpu : m2;
    }
    if (out4) {
        if (c->n == 0) { out4[0] = out4[1] = out4[2] = out4[3] = 0.0; }  /* WorstAvgPackingEfficiency */
        else {
            double len = (double)c->n;
            out4[0] = cpu / len; out4[1] = mem / len;
            out4[2] = with_gpu == 0 ? 1.0 : gpu / (double)with_gpu;
            out4[3] = mx / len;
        }
    }
    pool_free(effs, effs_bytes);
}

/* ComputeAvgPackingEfficiency (efficiency.go:114-156) over the node list [driver] + ExecutorNodes of one
 * result, exactly as chooseBestResult builds it (single_az.go:83-89; duplicates included, slice order). */
static void result_nodes_avg_efficiency(const orc_cluster* c, resmap* reserved, const char* driver,
                                        const char* const* execs, int32_t count, double* out4) {
    double cpu = 0, mem = 0, gpu = 0, mx = 0; int32_t with_gpu = 0;
    for (int32_t i = -1; i < count; ++i) {
        const char* n = i < 0 ? driver : execs[i];
        const node_meta* m = meta_lookup(c, n);
        pack_eff e = compute_packing_efficiency(m, resmap_get(reserved, n));
        cpu += e.cpu; mem += e.mem;
        if (m->schedulable.gpu != 0) { gpu += e.gpu; with_gpu++; }
        double m2 = e.cpu > e.mem ? e.cpu : e.mem;
        mx += e.gpu > m2 ? e.gpu : m2;
    }
    double len = (double)(count + 1);
    out4[0] = cpu / len; out4[1] = mem / len;
    out4[2] = with_gpu == 0 ? 1.0 : gpu / (double)with_gpu;
    out4[3] = mx / len;
}

/* ------------------------------------------------------------------ SparkBinPack ---- */
/* LIB/binpack/binpack.go:60-87.  executor_names: scratch of >= count entries. */
static int spark_bin_pack(const orc_cluster* c, const orc_res* drv, const orc_res* exe, int32_t count,
                          const char* const* driver_order, int32_t n_driver,
                          const char* const* exec_order, int32_t n_exec,
                          distribute_fn distribute, int with_efficiencies,
                          const char** driver_name, const char** executor_names, double* avg_eff,
                          double* choose_eff /* may be NULL: chooseBestResult's average */) {
    for (int32_t i = 0; i < n_driver; ++i) {                                   /* :67 */
        const char* d = driver_order[i];
        const node_meta* m = meta_lookup(c, d);                                /* :68 */
        if (m == NULL || res_greater_than(drv, &m->available)) continue;       /* :69-71 */
        resmap reserved;                                                       /* :72 make(map, len(metadata)) */
        resmap_init(&reserved, c->n);
        resmap_put(&reserved, d, drv);                                         /* :73 */
        int ok = distribute(c, exe, count, exec_order, n_exec, &reserved, executor_names); /* :74-75 */
        if (ok) {                                                              /* :76 */
            if (with_efficiencies) avg_packing_efficiency(c, &reserved, avg_eff); /* :77 */
            if (choose_eff) result_nodes_avg_efficiency(c, &reserved, d, executor_names, count, choose_eff);
            resmap_free(&reserved);
            *driver_name = d;                                                  /* :78-83 */
            return 1;
        }
        resmap_free(&reserved);
    }
    *driver_name = NULL;                                                       /* :86 EmptyPackingResult */
    if (avg_eff) { avg_eff[0] = avg_eff[1] = avg_eff[2] = avg_eff[3] = 0.0; }
    return 0;
}

/* internal/binpacker/binpack.go:43-58: name -> function; here by algo id. */
static distribute_fn select_distributor(int algo) {
    if (algo == ORC_MINIMAL_FRAGMENTATION) return minimal_fragmentation;       /* LIB/binpack/minimal_fragmentation.go:27-35 */
    return algo == ORC_DISTRIBUTE_EVENLY ? distribute_executors_evenly : tightly_pack_executors;
}

/* groupNodesByZone, LIB/binpack/single_az.go:57-73: zones in order of first appearance, names per zone in
 * order; names missing from the metadata are dropped. */
typedef struct { int32_t n_zones; const char** zone; const char*** names; int32_t* count; } zone_groups;

static void group_nodes_by_zone(const orc_cluster* c, const char* const* order, int32_t n, zone_groups* g) {
    g->n_zones = 0;
    g->zone = (const char**)malloc(sizeof(char*) * (size_t)(n > 0 ? n : 1));
    g->names = (const char***)malloc(sizeof(char**) * (size_t)(n > 0 ? n : 1));
    g->count = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; ++i) {
        const node_meta* m = meta_lookup(c, order[i]);
        if (m == NULL) continue;                                              /* :62-65 */
        int32_t z = -1;
        for (int32_t k = 0; k < g->n_zones; ++k) if (strcmp(g->zone[k], m->zone) == 0) { z = k; break; }
        if (z < 0) {                                                          /* :67-69 */
            z = g->n_zones++;
            g->zone[z] = m->zone;
            g->names[z] = (const char**)malloc(sizeof(char*) * (size_t)n);
            g->count[z] = 0;
        }
        g->names[z][g->count[z]++] = order[i];                                /* :70 */
    }
}
static void zone_groups_free(zone_groups* g) {
    for (int32_t k = 0; k < g->n_zones; ++k) free((void*)g->names[k]);
    free((void*)g->zone); free((void*)g->names); free(g->count);
}

/* getSingleAZSparkBinFunction + chooseBestResult, LIB/binpack/single_az.go:23-55,75-97 */
static int single_az_pack(const orc_cluster* c, const orc_res* drv, const orc_res* exe, int32_t count,
                          const char* const* driver_order, int32_t n_driver,
                          const char* const* exec_order, int32_t n_exec,
                          distribute_fn distribute, int with_efficiencies,
                          const char** driver_name, const char** executor_names, double* avg_eff) {
    zone_groups dz, ez;
    group_nodes_by_zone(c, driver_order, n_driver, &dz);                      /* :31 */
    group_nodes_by_zone(c, exec_order, n_exec, &ez);                          /* :32 */
    const char** trial = (const char**)malloc(sizeof(char*) * (size_t)(count > 0 ? count : 1));
    int have_best = 0;
    double best_max = 0.0;                                                    /* WorstAvgPackingEfficiency, :80 */
    double best_avg[4] = {0, 0, 0, 0};
    for (int32_t z = 0; z < dz.n_zones; ++z) {                                /* :36 */
        int32_t e = -1;
        for (int32_t k = 0; k < ez.n_zones; ++k) if (strcmp(ez.zone[k], dz.zone[z]) == 0) { e = k; break; }
        if (e < 0) continue;                                                  /* :38-41 */
        const char* dname = NULL; double all_eff[4], choose[4];
        int ok = spark_bin_pack(c, drv, exe, count, dz.names[z], dz.count[z], ez.names[e], ez.count[e], distribute,
                                with_efficiencies, &dname, trial, all_eff, choose);      /* :42 */
        if (!ok) continue;                                                    /* :44-46 */
        if (best_max < choose[3]) {                                           /* :91 LessThan compares Max only */
            have_best = 1; best_max = choose[3];
            *driver_name = dname;
            memcpy((void*)executor_names, trial, sizeof(char*) * (size_t)count);
            memcpy(best_avg, all_eff, sizeof(best_avg));
        }
    }
    free((void*)trial); zone_groups_free(&dz); zone_groups_free(&ez);
    if (!have_best) {                                                         /* :49-51 and the EmptyPackingResult start value of :79 */
        *driver_name = NULL;
        if (avg_eff) avg_eff[0] = avg_eff[1] = avg_eff[2] = avg_eff[3] = 0.0;
        return 0;
    }
    if (avg_eff && with_efficiencies) memcpy(avg_eff, best_avg, sizeof(best_avg));
    return 1;
}

/* the `binpack:` registry of internal/binpacker/binpack.go:43-49 restricted to the packers restated here */
static int pack_by_algo(const orc_cluster* c, int algo, const orc_res* drv, const orc_res* exe, int32_t count,
                        const char* const* driver_order, int32_t n_driver,
                        const char* const* exec_order, int32_t n_exec, int with_efficiencies,
                        const char** driver_name, const char** executor_names, double* avg_eff) {
    if (algo == ORC_SINGLE_AZ_TIGHTLY_PACK)                                   /* single_az_pack_tightly.go */
        return single_az_pack(c, drv, exe, count, driver_order, n_driver, exec_order, n_exec, tightly_pack_executors,
                              with_efficiencies, driver_name, executor_names, avg_eff);
    if (algo == ORC_SINGLE_AZ_MINIMAL_FRAGMENTATION)                          /* single_az_minimal_fragmentation.go:20 */
        return single_az_pack(c, drv, exe, count, driver_order, n_driver, exec_order, n_exec, minimal_fragmentation,
                              with_efficiencies, driver_name, executor_names, avg_eff);
    if (algo == ORC_AZ_AWARE_TIGHTLY_PACK) {                                  /* az_aware_pack_tightly.go:27-38 */
        if (single_az_pack(c, drv, exe, count, driver_order, n_driver, exec_order, n_exec, tightly_pack_executors,
                           with_efficiencies, driver_name, executor_names, avg_eff))
            return 1;
        return spark_bin_pack(c, drv, exe, count, driver_order, n_driver, exec_order, n_exec, tightly_pack_executors,
                              with_efficiencies, driver_name, executor_names, avg_eff, NULL);
    }
    return spark_bin_pack(c, drv, exe, count, driver_order, n_driver, exec_order, n_exec, select_distributor(algo),
                          with_efficiencies, driver_name, executor_names, avg_eff, NULL);
}

int orc_binpack(const orc_cluster* c, int algo, const orc_res* drv, const orc_res* exe, int32_t count,
                const char* const* driver_order, int32_t n_driver,
                const char* const* exec_order, int32_t n_exec,
                int with_efficiencies,
                int32_t* driver_node, int32_t* executor_nodes, double* avg_eff) {
    const char** names = (const char**)malloc(sizeof(char*) * (size_t)(count > 0 ? count : 1));
    const char* dname = NULL;
    int ok = pack_by_algo(c, algo, drv, exe, count, driver_order, n_driver, exec_order, n_exec,
                          with_efficiencies, &dname, names, avg_eff);
    *driver_node = ok ? orc_cluster_index(c, dname) : -1;
    if (ok) for (int32_t i = 0; i < count; ++i) executor_nodes[i] = orc_cluster_index(c, names[i]);
    free(names);
    return ok;
}

/* ------------------------------------------------------------------ independent batch ---- */
typedef struct {
    const orc_cluster* c; int algo; int32_t lo, hi;
    const orc_res *drv, *exe; const int32_t* count;
    const char* const* driver_order; int32_t n_driver;
    const char* const* exec_order; int32_t n_exec;
    int with_eff; const int64_t* exec_off; int32_t* driver_node; int32_t* executor_nodes;
} batch_job;

static void* batch_worker(void* p) {
    batch_job* j = (batch_job*)p;
    for (int32_t i = j->lo; i < j->hi; ++i) {
        double eff[4];
        orc_binpack(j->c, j->algo, &j->drv[i], &j->exe[i], j->count[i], j->driver_order, j->n_driver,
                    j->exec_order, j->n_exec, j->with_eff, &j->driver_node[i],
                    j->executor_nodes + j->exec_off[i], eff);
    }
    return NULL;
}

void orc_binpack_batch(const orc_cluster* c, int algo, int32_t n_apps,
                       const orc_res* drv, const orc_res* exe, const int32_t* count,
                       const char* const* driver_order, int32_t n_driver,
                       const char* const* exec_order, int32_t n_exec,
                       int with_efficiencies, int n_threads,
                       const int64_t* exec_off, int32_t* driver_node, int32_t* executor_nodes) {
    tune_allocator();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_apps) n_threads = n_apps > 0 ? n_apps : 1;
    batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * (size_t)n_threads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        batch_job j = {c, algo, (int32_t)((int64_t)n_apps * t / n_threads), (int32_t)((int64_t)n_apps * (t + 1) / n_threads),
                       drv, exe, count, driver_order, n_driver, exec_order, n_exec,
                       with_efficiencies, exec_off, driver_node, executor_nodes};
        jobs[t] = j;
        if (n_threads == 1) batch_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    if (n_threads > 1) for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th);
}

/* ------------------------------------------------------------------ FIFO loop ---- */
/* fitEarlierDrivers, EXT/resource.go:224-262 (+ the caller's own pack at :321 as the last app). */
int32_t orc_fifo(orc_cluster* c, int algo, int mode, int32_t n_apps,
                 const orc_res* drv, const orc_res* exe, const int32_t* count, const uint8_t* young,
                 const char* const* driver_order, int32_t n_driver,
                 const char* const* exec_order, int32_t n_exec,
                 int with_efficiencies,
                 const int64_t* exec_off, int32_t* driver_node, int32_t* executor_nodes) {
    int32_t blocked = -1;
    int32_t maxc = 1;
    for (int32_t i = 0; i < n_apps; ++i) if (count[i] > maxc) maxc = count[i];
    const char** names = (const char**)malloc(sizeof(char*) * (size_t)maxc);
    int32_t i = 0;
    for (; i < n_apps; ++i) {                                                    /* :230 */
        const char* dname = NULL; double eff[4];
        int ok = pack_by_algo(c, algo, &drv[i], &exe[i], count[i], driver_order, n_driver, exec_order, n_exec,
                              with_efficiencies, &dname, names, eff); /* :238-243 */
        if (!ok) {                                                               /* :244 */
            driver_node[i] = -1;
            if (young && young[i]) continue;                                     /* :245-249 shouldSkipDriverFifo */
            blocked = i;                                                         /* :250-252 */
            ++i;
            break;
        }
        driver_node[i] = orc_cluster_index(c, dname);
        for (int32_t k = 0; k < count[i]; ++k) executor_nodes[exec_off[i] + k] = orc_cluster_index(c, names[k]);

        if (mode == ORC_FIFO_REFERENCE) {
            /* sparkResourceUsage, EXT/sparkpods.go:139-146: map ASSIGNMENT, not accumulation */
            resmap usage; resmap_init(&usage, count[i] + 1);
            resmap_put(&usage, dname, &drv[i]);                                  /* :141 */
            for (int32_t k = 0; k < count[i]; ++k) resmap_put(&usage, names[k], &exe[i]); /* :142-144 */
            /* SubtractUsageIfExists, LIB/resources/resources.go:129-135 */
            for (uint32_t s = 0; s < usage.idx.cap; ++s) {
                const char* k = usage.idx.keys[s];
                if (k == NULL || k == TOMB) continue;
                node_meta* m = meta_lookup(c, k);
                if (m) res_sub(&m->available, &usage.vals[usage.idx.vals[s]]);
            }
            resmap_free(&usage);
        } else {
            /* exact accounting: every placed pod is charged */
            node_meta* m = meta_lookup(c, dname);
            if (m) res_sub(&m->available, &drv[i]);
            for (int32_t k = 0; k < count[i]; ++k) {
                node_meta* e = meta_lookup(c, names[k]);
                if (e) res_sub(&e->available, &exe[i]);
            }
        }
    }
    for (; i < n_apps; ++i) driver_node[i] = -2;  /* never evaluated: fitEarlierDrivers returned false */
    free(names);
    return blocked;
}

/* ------------------------------------------------------------------ executor reschedule ---- */
/* The node choice of rescheduleExecutor (EXT/resource.go:594-673) for one executor pod.
 * first fit (:657-662): `available` of the cluster plays availableResources (:643).
 * minimal fragmentation (rescheduleExecutorWithMinimalFragmentation, :675-705): `available` plays
 * availableNodesSchedulingMetadata (:640), reserved_* the overhead map handed to GetNodeCapacities as "reserved"
 * (:682), hosting_names the nodes of getNodesWithExecutorsBelongingToSameApp (:683). */
int32_t orc_reschedule_executor(const orc_cluster* c, int min_frag, const orc_res* exe,
                                const char* const* exec_order, int32_t n_exec,
                                const char* const* reserved_names, const orc_res* reserved, int32_t n_reserved,
                                const char* const* hosting_names, int32_t n_hosting) {
    static const orc_res zero = {0, 0, 0};
    if (!min_frag) {
        for (int32_t i = 0; i < n_exec; ++i) {                                            /* :658 */
            const node_meta* m = meta_lookup(c, exec_order[i]);
            if (m == NULL) continue;       /* PotentialNodes only yields known names; a missing one cannot be chosen */
            if (!res_greater_than(exe, &m->available)) return orc_cluster_index(c, exec_order[i]);   /* :659-661 */
        }
        return -1;                                                                        /* :672 */
    }
    resmap over; resmap_init(&over, n_reserved > 0 ? n_reserved : 4);
    for (int32_t i = 0; i < n_reserved; ++i) resmap_put(&over, reserved_names[i], &reserved[i]);
    strmap hosting; strmap_init(&hosting, (uint32_t)(n_hosting > 0 ? n_hosting : 4));    /* map[string]bool, :683 */
    for (int32_t i = 0; i < n_hosting; ++i)
        if (strmap_find(&hosting, hosting_names[i]) == NULL) strmap_insert(&hosting, hosting_names[i], 1);
    const char* best = NULL; int64_t best_cap = 0; int best_hosts = 0;                    /* var best, :685 */
    for (int32_t i = 0; i < n_exec; ++i) {                      /* GetNodeCapacities (capacity.go:78-102) fused with the loop :686 */
        const node_meta* m = meta_lookup(c, exec_order[i]);
        if (m == NULL) continue;                                                          /* capacity.go:87 */
        const orc_res* r = resmap_get(&over, exec_order[i]);                              /* capacity.go:88-92 */
        if (r == NULL) r = &zero;
        int64_t cap = node_capacity(&m->available, r, exe);
        if (cap < 1) continue;                                                            /* :687 */
        int hosts = strmap_find(&hosting, exec_order[i]) != NULL;
        if (best == NULL                                                                  /* :689-691 */
            || (hosts && !best_hosts)                                                     /* :692-694 */
            || (hosts == best_hosts && cap < best_cap)) {                                 /* :695-697 */
            best = exec_order[i]; best_cap = cap; best_hosts = hosts;
        }
    }
    resmap_free(&over); strmap_free(&hosting);
    return best ? orc_cluster_index(c, best) : -1;                                        /* :702 */
}

/* ------------------------------------------------------------------ snapshot build ---- */
/* UsageForNodes (LIB/resources/resources.go:31-43) over hard reservations + UsedSoftReservationResources
 * (internal/cache/softreservations.go:155-170), summed like GetReservedResources
 * (EXT/resourcereservations.go:258-263); then NodeSchedulingMetadataForNodes (resources.go:61-100):
 * available = allocatable - (usage + overhead), schedulable = allocatable - overhead, per listed node. */
void orc_node_scheduling_metadata(int32_t n_nodes, const char* const* names,
                                  const int64_t* alloc_cpu, const int64_t* alloc_mem, const int64_t* alloc_gpu,
                                  const int64_t* over_cpu, const int64_t* over_mem, const int64_t* over_gpu,
                                  int64_t n_res, const char* const* res_node_name,
                                  const int64_t* res_cpu, const int64_t* res_mem, const int64_t* res_gpu,
                                  int64_t* avail_cpu, int64_t* avail_mem, int64_t* avail_gpu,
                                  int64_t* sched_cpu, int64_t* sched_mem, int64_t* sched_gpu) {
    static const orc_res zero = {0, 0, 0};
    resmap usage;                                              /* resources.go:32 */
    resmap_init(&usage, n_nodes);
    for (int64_t r = 0; r < n_res; ++r) {                      /* :33-41 (and softreservations.go:160-168) */
        orc_res* u = resmap_get(&usage, res_node_name[r]);
        if (u == NULL) u = resmap_put(&usage, res_node_name[r], &zero);
        orc_res add = {res_cpu[r], res_mem[r], res_gpu ? res_gpu[r] : 0};
        res_add(u, &add);                                      /* AddFromReservation, :186-190 */
    }
    for (int32_t i = 0; i < n_nodes; ++i) {                    /* resources.go:67 */
        orc_res overhead = {over_cpu ? over_cpu[i] : 0, over_mem ? over_mem[i] : 0, over_gpu ? over_gpu[i] : 0};   /* :68-71 */
        orc_res* u = resmap_get(&usage, names[i]);             /* :72-75 */
        orc_res cur = u ? *u : zero;
        res_add(&cur, &overhead);                              /* :76 */
        orc_res alloc = {alloc_cpu[i], alloc_mem[i], alloc_gpu ? alloc_gpu[i] : 0};
        orc_res a = alloc, sc = alloc;
        res_sub(&a, &cur);                                     /* :89 subtractFromResourceList(allocatable, usage+overhead) */
        res_sub(&sc, &overhead);                               /* :90 */
        avail_cpu[i] = a.cpu; avail_mem[i] = a.mem; avail_gpu[i] = a.gpu;
        if (sched_cpu) { sched_cpu[i] = sc.cpu; sched_mem[i] = sc.mem; sched_gpu[i] = sc.gpu; }
    }
    resmap_free(&usage);
}

/* availableResources of rescheduleExecutor's first-fit branch (EXT/resource.go:638-643), statement by statement:
 *   usage := GetReservedResources()                                             :638
 *   availableNodesSchedulingMetadata := NodeSchedulingMetadataForNodes(availableNodes, usage, overhead)   :640
 *        -- which ADDS the overhead into usage's existing entries in place (resources.go:72-76: currentUsageForNode is the
 *           map's own *Resources when the node has an entry, a fresh Zero() otherwise)
 *   usage.Add(overhead)                                                         :642  (creates missing entries)
 *   availableResources := AvailableForNodes(availableNodes, usage)              :643
 * => nodes that carry at least one reservation lose their overhead TWICE. */
void orc_reschedule_available(int32_t n_nodes, const char* const* names,
                              const int64_t* alloc_cpu, const int64_t* alloc_mem, const int64_t* alloc_gpu,
                              const int64_t* over_cpu, const int64_t* over_mem, const int64_t* over_gpu,
                              int64_t n_res, const char* const* res_node_name,
                              const int64_t* res_cpu, const int64_t* res_mem, const int64_t* res_gpu,
                              int64_t* avail_cpu, int64_t* avail_mem, int64_t* avail_gpu) {
    static const orc_res zero = {0, 0, 0};
    resmap usage;
    resmap_init(&usage, n_nodes);
    for (int64_t r = 0; r < n_res; ++r) {                      /* GetReservedResources: UsageForNodes + soft reservations */
        orc_res* u = resmap_get(&usage, res_node_name[r]);
        if (u == NULL) u = resmap_put(&usage, res_node_name[r], &zero);
        orc_res add = {res_cpu[r], res_mem[r], res_gpu ? res_gpu[r] : 0};
        res_add(u, &add);
    }
    for (int32_t i = 0; i < n_nodes; ++i) {                    /* NodeSchedulingMetadataForNodes, resources.go:67-76 */
        orc_res overhead = {over_cpu ? over_cpu[i] : 0, over_mem ? over_mem[i] : 0, over_gpu ? over_gpu[i] : 0};
        orc_res* u = resmap_get(&usage, names[i]);
        if (u != NULL) res_add(u, &overhead);                  /* :76 mutates the map's entry; a missing entry is a temporary */
    }
    for (int32_t i = 0; i < n_nodes; ++i) {                    /* usage.Add(overhead), resources.go:109-116 */
        orc_res overhead = {over_cpu ? over_cpu[i] : 0, over_mem ? over_mem[i] : 0, over_gpu ? over_gpu[i] : 0};
        orc_res* u = resmap_get(&usage, names[i]);
        if (u == NULL) u = resmap_put(&usage, names[i], &zero);
        res_add(u, &overhead);
    }
    for (int32_t i = 0; i < n_nodes; ++i) {                    /* AvailableForNodes, resources.go:46-56 */
        const orc_res* u = resmap_get(&usage, names[i]);
        orc_res a = {alloc_cpu[i], alloc_mem[i], alloc_gpu ? alloc_gpu[i] : 0};
        if (u) res_sub(&a, u);
        avail_cpu[i] = a.cpu; avail_mem[i] = a.mem; avail_gpu[i] = a.gpu;
    }
    resmap_free(&usage);
}

/* ------------------------------------------------------------------ node sorting ---- */
/* resourcesLessThan, internal/sort/nodesorting.go:74-80 */
static int resources_less_than(const orc_res* l, const orc_res* r) {
    if (l->mem != r->mem) return l->mem < r->mem;
    return l->cpu < r->cpu;
}
typedef struct { int32_t az_priority; const orc_res* res; const char* name; int32_t idx; } sched_ctx;
/* scheduleContextLessThan, internal/sort/nodesorting.go:83-93 */
static int schedule_context_less_than(const sched_ctx* l, const sched_ctx* r) {
    if (l->az_priority != r->az_priority) return l->az_priority < r->az_priority;
    if (!res_eq(l->res, r->res)) return resources_less_than(l->res, r->res);
    return strcmp(l->name, r->name) < 0;
}

/* stable merge sort on sched_ctx / generic index arrays */
static void msort_ctx(sched_ctx* a, sched_ctx* tmp, int32_t n) {
    if (n < 2) return;
    int32_t h = n / 2;
    msort_ctx(a, tmp, h); msort_ctx(a + h, tmp, n - h);
    int32_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = schedule_context_less_than(&a[j], &a[i]) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(sched_ctx) * (size_t)n);
}

/* createLabelLessThanFunction, internal/sort/nodesorting.go:161-180, on precomputed ranks:
 * rank1 unknown -> false; rank2 unknown -> true; else rank1 < rank2. */
static int label_less_than(int32_t r1, int32_t r2) {
    if (r1 < 0) return 0;
    if (r2 < 0) return 1;
    return r1 < r2;
}
/* sort.SliceStable(nodeNames, lessThan)  (:191-200) */
static void stable_sort_by_label(int32_t* a, int32_t* tmp, int32_t n, const int32_t* rank) {
    if (n < 2) return;
    int32_t h = n / 2;
    stable_sort_by_label(a, tmp, h, rank); stable_sort_by_label(a + h, tmp, n - h, rank);
    int32_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = label_less_than(rank[a[j]], rank[a[i]]) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(int32_t) * (size_t)n);
}

void orc_potential_nodes(const orc_cluster* c, const char* const* candidate_names, int32_t n_candidates,
                         const int32_t* driver_label_rank, const int32_t* exec_label_rank,
                         int32_t* driver_out, int32_t* n_driver_out,
                         int32_t* exec_out, int32_t* n_exec_out) {
    int32_t n = c->n;
    /* getNodeNamesInPriorityOrder, :95-122 */
    /* groupNodeNamesByAZ :136-143 + getAvailableResourcesByAZ :124-134 */
    strmap az_index; strmap_init(&az_index, 16);
    int32_t n_az = 0, az_cap = 16;
    const char** az_label = (const char**)malloc(sizeof(char*) * (size_t)az_cap);
    orc_res* az_res = (orc_res*)malloc(sizeof(orc_res) * (size_t)az_cap);
    int32_t* node_az = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; ++i) {
        int32_t* s = strmap_find(&az_index, c->meta[i].zone);
        int32_t a;
        if (s) a = *s;
        else {
            if (n_az == az_cap) {
                az_cap *= 2;
                az_label = (const char**)realloc(az_label, sizeof(char*) * (size_t)az_cap);
                az_res = (orc_res*)realloc(az_res, sizeof(orc_res) * (size_t)az_cap);
            }
            a = n_az++;
            az_label[a] = c->meta[i].zone;
            az_res[a].cpu = az_res[a].mem = az_res[a].gpu = 0;
            strmap_insert(&az_index, c->meta[i].zone, a);
        }
        node_az[i] = a;
        res_add(&az_res[a], &c->meta[i].available);
    }
    /* sort AZ labels by resourcesLessThan (:102-104).  Unstable in the reference; here: insertion
     * sort (stable) over first-seen order. */
    int32_t* az_order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_az > 0 ? n_az : 1));
    for (int32_t a = 0; a < n_az; ++a) az_order[a] = a;
    for (int32_t a = 1; a < n_az; ++a) {
        int32_t v = az_order[a], b = a - 1;
        while (b >= 0 && resources_less_than(&az_res[v], &az_res[az_order[b]])) { az_order[b + 1] = az_order[b]; --b; }
        az_order[b + 1] = v;
    }
    int32_t* az_priority = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_az > 0 ? n_az : 1));
    for (int32_t p = 0; p < n_az; ++p) az_priority[az_order[p]] = p;   /* :106-115 */

    sched_ctx* ctx = (sched_ctx*)malloc(sizeof(sched_ctx) * (size_t)(n > 0 ? n : 1));
    sched_ctx* tmp = (sched_ctx*)malloc(sizeof(sched_ctx) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; ++i) {
        ctx[i].az_priority = az_priority[node_az[i]];
        ctx[i].res = &c->meta[i].available;
        ctx[i].name = c->names[i];
        ctx[i].idx = i;
    }
    msort_ctx(ctx, tmp, n);                                              /* :117-119 */

    /* PotentialNodes, :41-64 */
    strmap cand; strmap_init(&cand, (uint32_t)n_candidates);             /* :46-49 */
    for (int32_t i = 0; i < n_candidates; ++i)
        if (strmap_find(&cand, candidate_names[i]) == NULL) strmap_insert(&cand, candidate_names[i], 1);
    int32_t nd = 0, ne = 0;
    for (int32_t i = 0; i < n; ++i) {                                    /* :51 */
        int32_t node = ctx[i].idx;
        if (strmap_find(&cand, c->names[node]) != NULL) driver_out[nd++] = node;         /* :52-54 */
        if (!c->meta[node].unschedulable && c->meta[node].ready) exec_out[ne++] = node;  /* :55-57 */
    }
    int32_t* itmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    if (driver_label_rank) stable_sort_by_label(driver_out, itmp, nd, driver_label_rank); /* :61 */
    if (exec_label_rank) stable_sort_by_label(exec_out, itmp, ne, exec_label_rank);       /* :62 */
    *n_driver_out = nd; *n_exec_out = ne;

    free(itmp); strmap_free(&cand); free(ctx); free(tmp); free(az_priority); free(az_order);
    free(node_az); free(az_res); free(az_label); strmap_free(&az_index);
}

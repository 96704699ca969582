/*
 * gangpack.h -- C ABI of libgangpack.so, the B200-native gang-scheduling bin-packer.
 *
 * This is the drop-in boundary for ONE hot path of palantir/k8s-spark-scheduler: what a cgo shim
 * registered in internal/binpacker.binpackFunctions (internal/binpacker/binpack.go:43-49) binds in
 * place of binpack.TightlyPack / binpack.DistributeEvenly
 * (vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg/binpack/pack_tightly.go:25-32,
 *  distribute_evenly.go:25-32; contract binpack.SparkBinPackFunction, binpack.go:43-48) and what a
 * batched fitEarlierDrivers (internal/extender/resource.go:224-262) calls once per Predicate.
 * The reference-side binding is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, plain pointers and sizes; no exceptions cross the boundary; every call returns a
 *    gp_status (0 = ok) and gp_last_error() describes the last failure on that context.
 *  - Quantities are exact int64: CPU in millicores (Quantity.MilliValue), memory in bytes, GPU in
 *    units -- the int64 model of resource.Quantity.  Inputs that are not
 *    representable that way must be routed by the caller to the original Go packer.
 *  - Nodes are addressed by index into the caller's node table; names never cross the boundary.
 *  - A gp_ctx is NOT thread-safe (one per calling goroutine role, or a mutex), owns one CUDA
 *    stream, and there is NO CPU fallback: if no CUDA device is usable gp_create fails.
 */
#ifndef GANGPACK_H
#define GANGPACK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 1

typedef struct gp_ctx gp_ctx;

typedef enum {
    GP_OK = 0,
    GP_ERR_INVALID = 1,       /* bad argument / precondition (negative request, index out of range, duplicate in an order ...) */
    GP_ERR_CUDA = 2,          /* CUDA runtime error; message in gp_last_error */
    GP_ERR_NO_DEVICE = 3,     /* no usable sm_100 device: the library has no CPU path */
    GP_ERR_NO_SNAPSHOT = 4,   /* gp_pack_* before gp_set_snapshot */
    GP_ERR_CAPACITY = 5,      /* result buffer too small */
    GP_ERR_UNREPRESENTABLE = 6 /* quantity outside the exact-int64 domain (|v| >= 2^61) */
} gp_status;

/* binpack algorithm: the `binpack:` config values of internal/binpacker/binpack.go:21-24 */
typedef enum {
    GP_TIGHTLY_PACK = 0,      /* "tightly-pack"      -> pack_tightly.go:34-63, ExecutorNodes node-major  */
    GP_DISTRIBUTE_EVENLY = 1, /* "distribute-evenly" -> distribute_evenly.go:34-73, round-major (also the
                                 fallback for unknown names, binpack.go:52-57) */
    GP_MINIMAL_FRAGMENTATION = 2 /* binpack.MinimalFragmentation -> LIB/binpack/minimal_fragmentation.go:27-137 over
                                 LIB/capacity/capacity.go:36-113: as few nodes as possible, fullest-fitting node first;
                                 ExecutorNodes in (capacity descending, priority order) then the remainder's node.
                                 GP_MODE_INDEPENDENT only: the building block of "single-az-minimal-fragmentation"
                                 (internal/binpacker/binpack.go:48), which packs one application per zone (zone =
                                 instance group) and chooses the zone on the host */
} gp_algo;

typedef enum {
    /* every app is packed against the same unmodified snapshot: the shape of
       internal/extender/unschedulablepods.go:132-166 and of each Predicate's own pack (resource.go:321) */
    GP_MODE_INDEPENDENT = 0,
    /* fitEarlierDrivers (resource.go:224-262): apps in queue order per instance group against a
       MUTATING snapshot, usage charged like sparkResourceUsage (sparkpods.go:139-146: each distinct
       executor node one executor; a driver node that hosts an executor is charged the executor only) */
    GP_MODE_FIFO_REFERENCE = 1,
    /* same loop with exact accounting (driver + every executor charged), i.e. what separate
       Predicate calls converge to through UsageForNodes (resources.go:31-43) */
    GP_MODE_FIFO_EXACT = 2
} gp_mode;

/* gp_config.flags */
#define GP_CFG_ASYNC_SNAPSHOT 1  /* gp_set_snapshot with page-locked (gp_alloc_pinned / gp_register_host) inputs returns while the
                                    device is still reading them: the buffers must stay unchanged until the next gp_pack_* /
                                    gp_get_snapshot / gp_synchronize call on the context has returned.  The upload then overlaps the
                                    input copies of that call (one Predicate = snapshot + pack, back to back) */
typedef struct {
    int32_t device;          /* CUDA device ordinal; -1 = current device */
    int32_t flags;           /* GP_CFG_* */
    int32_t reserved[6];
} gp_config;

/* The node snapshot (NodeGroupSchedulingMetadata.AvailableResources, resources.go:61-100) plus the
 * two priority orders of NodeSorter.PotentialNodes (internal/sort/nodesorting.go:41-64), per
 * instance group.  Group g owns exec_order[exec_off[g]..exec_off[g+1]) and
 * drv_order[drv_off[g]..drv_off[g+1]); entries are indices into the node table, without
 * duplicates inside a group.  A driver candidate need not be an executor candidate.
 * Names absent from the metadata map are simply omitted by the caller (binpack.go:68-69,
 * pack_tightly.go:51-52: they can neither host a driver nor an executor). */
typedef struct {
    int32_t n_nodes;
    const int64_t* avail_cpu_milli;  /* [n_nodes] may be negative (over-committed node) */
    const int64_t* avail_mem_bytes;  /* [n_nodes] */
    const int64_t* avail_gpu;        /* [n_nodes] or NULL (= all zero) */
    int32_t n_groups;                /* >= 1 */
    const int32_t* exec_off;         /* [n_groups+1] */
    const int32_t* exec_order;       /* [exec_off[n_groups]] */
    const int32_t* drv_off;          /* [n_groups+1] */
    const int32_t* drv_order;        /* [drv_off[n_groups]] */
} gp_nodes;

/* The pending-application queue (types.SparkApplicationResources, internal/types/types.go:22-27, as
 * produced by sparkResources, internal/extender/sparkpods.go:73-137): SoA, queue order = index
 * order (creation-time ascending, sparkpods.go:67-69).  All requests must be >= 0. */
typedef struct {
    int32_t n_apps;
    const int64_t* drv_cpu_milli;    /* [n_apps] driver resources */
    const int64_t* drv_mem_bytes;
    const int64_t* drv_gpu;          /* or NULL (= 0) */
    const int64_t* exe_cpu_milli;    /* [n_apps] per-executor resources */
    const int64_t* exe_mem_bytes;
    const int64_t* exe_gpu;          /* or NULL (= 0) */
    const int32_t* exe_count;        /* [n_apps] MinExecutorCount (resource.go:242,325); >= 0, <= 2^24 (<= 2^20 in the
                                        FIFO modes); larger -> GP_ERR_UNREPRESENTABLE */
    const int32_t* group;            /* [n_apps] instance group, or NULL (= 0) */
    const uint8_t* skip_if_no_fit;   /* [n_apps] FIFO only: shouldSkipDriverFifo (resource.go:264-270), or NULL */
    const int64_t* exec_out_off;     /* [n_apps+1] exclusive prefix sum of exe_count, or NULL (computed) */
} gp_apps;

/* binpack.PackingResult (binpack.go:25-30) for a batch.
 *   driver_node[i] >= 0 : HasCapacity, DriverNode = that node index
 *                  == -1: HasCapacity=false (EmptyPackingResult)
 *                  == -2: not evaluated -- an earlier app of its group blocked the FIFO queue
 *   executor_nodes[exec_out_off[i] .. +exe_count[i]) : ExecutorNodes, in the reference's order
 *     (reservation names executor-1..k follow it, resourcereservations.go:501-510); its contents are
 *     UNDEFINED unless driver_node[i] >= 0 (the FIFO kernels emit optimistically before feasibility is known).
 * PackingEfficiencies are not produced on the device (metrics by-product). */
typedef struct {
    int32_t* driver_node;            /* [n_apps] */
    int32_t* executor_nodes;         /* [executor_nodes_cap] */
    int64_t executor_nodes_cap;      /* entries available; must be >= sum(exe_count) */
} gp_results;

/* ---- lifecycle ------------------------------------------------------------------------- */
int gp_abi_version(void);
gp_status gp_create(gp_ctx** out, const gp_config* cfg /* may be NULL */);
void gp_destroy(gp_ctx* ctx);
const char* gp_last_error(const gp_ctx* ctx /* NULL = creation errors */);
/* 1 = CUDA (the only backend).  Present so a shim can assert it is not on a silent fallback. */
int gp_backend(const gp_ctx* ctx);

/* Pinned host memory for the SoA buffers that cross the boundary (cudaHostAlloc); a Go caller
 * wraps it with unsafe.Slice.  Pageable memory is accepted everywhere, just slower. */
gp_status gp_alloc_pinned(gp_ctx* ctx, size_t bytes, void** out);
gp_status gp_free_pinned(gp_ctx* ctx, void* p);

/* Page-locks caller-owned host memory (cudaHostRegister, portable + mapped) so that gp_pack_batch can DMA results
 * straight into it / read inputs in place -- e.g. a POSIX shared-memory segment that the scheduler process and the
 * per-GPU worker processes of one box share: every worker copies its block of the placements over its OWN PCIe
 * link into the scheduler's buffer (no gather through one GPU).  The range must stay mapped until unregistered. */
gp_status gp_register_host(gp_ctx* ctx, void* p, size_t bytes);
gp_status gp_unregister_host(gp_ctx* ctx, void* p);

/* ---- snapshot ------------------------------------------------------------------------------ */
/* Validates, copies to the device and lays the snapshot out in executor-priority order. */
gp_status gp_set_snapshot(gp_ctx* ctx, const gp_nodes* nodes);
/* Current device snapshot back in node-table order (after FIFO modes: with the usage subtracted,
 * i.e. metadata after SubtractUsageIfExists, resources.go:129-135).  Any pointer may be NULL. */
gp_status gp_get_snapshot(gp_ctx* ctx, int64_t* avail_cpu_milli, int64_t* avail_mem_bytes, int64_t* avail_gpu);

/* ---- availability snapshot from reservations ------------------------------ */
/* GetReservedResources (internal/extender/resourcereservations.go:258-263: UsageForNodes over the hard
 * reservations, LIB/resources/resources.go:31-43, plus the soft reservations, internal/cache/
 * softreservations.go:155-170) followed by NodeSchedulingMetadataForNodes (resources.go:61-100):
 *   available[n]   = allocatable[n] - (sum of reservations on n + overhead[n])
 *   schedulable[n] = allocatable[n] - overhead[n]
 * One entry of res_* per reservation (hard and soft alike); res_node = index into the node table, or -1 for a
 * reservation whose node is not in the table (ignored, like the map lookup at resources.go:72-75). */
typedef struct {
    int32_t n_nodes;
    const int64_t* alloc_cpu_milli;      /* [n_nodes] node.Status.Allocatable */
    const int64_t* alloc_mem_bytes;
    const int64_t* alloc_gpu;            /* or NULL (= 0) */
    const int64_t* overhead_cpu_milli;   /* [n_nodes] OverheadComputer.GetOverhead, or NULL (= 0) */
    const int64_t* overhead_mem_bytes;   /* or NULL */
    const int64_t* overhead_gpu;         /* or NULL */
    int64_t n_reservations;
    const int32_t* res_node;             /* [n_reservations] */
    const int64_t* res_cpu_milli;
    const int64_t* res_mem_bytes;
    const int64_t* res_gpu;              /* or NULL (= 0) */
} gp_usage_input;
/* Outputs are [n_nodes] each; any of them may be NULL. */
gp_status gp_build_availability(gp_ctx* ctx, const gp_usage_input* in,
                                int64_t* avail_cpu_milli, int64_t* avail_mem_bytes, int64_t* avail_gpu,
                                int64_t* sched_cpu_milli, int64_t* sched_mem_bytes, int64_t* sched_gpu);

/* availableResources of rescheduleExecutor's FIRST-FIT branch (internal/extender/resource.go:638-643), bug for bug: the
 * reference calls NodeSchedulingMetadataForNodes(availableNodes, usage, overhead) -- which adds the overhead into the usage
 * map's EXISTING entries in place (resources.go:72-76) -- and then usage.Add(overhead) once more, so
 *   available[n] = allocatable[n] - (sum of reservations on n + overhead[n] * (n carries a reservation ? 2 : 1))
 * (SURVEY App. B7).  This is the availability to upload with gp_set_snapshot before gp_reschedule_executors(min_frag = 0);
 * the minimal-fragmentation branch uses the plain gp_build_availability output (:640, :682). */
gp_status gp_build_reschedule_availability(gp_ctx* ctx, const gp_usage_input* in,
                                           int64_t* avail_cpu_milli, int64_t* avail_mem_bytes, int64_t* avail_gpu);

/* ---- node priority order (the step before the hot path) ------------------- */
/* NodeSorter.PotentialNodes (internal/sort/nodesorting.go:41-64) on the device: nodes in ascending
 * (AZ priority, available memory, available CPU, name) order (:83-122), split into the driver candidates
 * (restricted to kube-scheduler's NodeNames, :52-54) and the executor candidates (schedulable && ready,
 * :55-57), each optionally re-sorted stably by the rank of a configured label value (:61-62,161-200).
 * The outputs are node indices and can be passed straight to gp_nodes.drv_order / exec_order.
 * Where the reference's comparators leave the order undefined (equal zone totals; equal (memory, cpu) with
 * different gpu) zones are ordered by id and nodes by name. */
typedef struct {
    int32_t n_nodes;
    const int64_t* avail_cpu_milli;      /* [n_nodes] */
    const int64_t* avail_mem_bytes;      /* [n_nodes] */
    int32_t n_zones;                     /* >= 1 */
    const int32_t* zone_id;              /* [n_nodes] dense ids 0..n_zones-1 (ZoneLabel interned in first-seen order), or NULL (one zone) */
    const int32_t* name_rank;            /* [n_nodes] rank of the node name in ascending byte order (unique), or NULL (= index order) */
    const uint8_t* is_driver_candidate;  /* [n_nodes] member of ExtenderArgs.NodeNames, or NULL (= all) */
    const uint8_t* unschedulable;        /* [n_nodes] node.Spec.Unschedulable, or NULL (= none) */
    const uint8_t* ready;                /* [n_nodes] NodeReady == True, or NULL (= all) */
    const int32_t* driver_label_rank;    /* [n_nodes] rank of the node's value of driver-prioritized-node-label, -1 unknown; NULL = not configured */
    const int32_t* executor_label_rank;  /* same for executor-prioritized-node-label */
    const int64_t* avail_gpu;            /* [n_nodes] or NULL: only used to DETECT the ties below (gp_prepare_cluster: taken from the
                                            availability it computes) */
    int32_t* undefined_ties;             /* out, or NULL: number of adjacent pairs of the node order whose relative order the
                                            reference's comparator leaves undefined -- same zone priority, memory and cpu but
                                            different gpu (scheduleContextLessThan, nodesorting.go:83-93, is then "not less" both
                                            ways and sort.Slice is unstable).  Non-zero: this library ordered them
                                            by name; a shim that wants the reference's own (unspecified) choice falls back to Go */
} gp_sort_input;
gp_status gp_potential_nodes(gp_ctx* ctx, const gp_sort_input* in,
                             int32_t* driver_order /* [n_nodes] */, int32_t* n_driver,
                             int32_t* executor_order /* [n_nodes] */, int32_t* n_executor);

/* ---- everything before the hot path in one call ---------------------------------------------- */
/* gp_build_availability -> gp_potential_nodes -> gp_set_snapshot chained on the device: the availability computed
 * from the reservations never returns to the host, the two priority orders are consumed where they are produced
 * and the context ends up with a one-group snapshot ready for gp_pack_batch (what selectDriverNode does at
 * internal/extender/resource.go:300-304 before the FIFO loop).  sort->avail_* are ignored.  Node indices in later
 * results refer to this node table. */
gp_status gp_prepare_cluster(gp_ctx* ctx, const gp_usage_input* usage, const gp_sort_input* sort,
                             int32_t* n_driver /* may be NULL */, int32_t* n_executor /* may be NULL */);

/* ---- executors without a usable reservation (the step after the hot path) --- */
/* rescheduleExecutor's node choice (internal/extender/resource.go:594-673) for a BATCH of executor pods, every
 * decision independent against the current snapshot (its executor priority orders and availabilities):
 *   min_frag == 0: the first node of the order the executor fits on (:657-662);
 *   min_frag != 0: rescheduleExecutorWithMinimalFragmentation (:675-705) -- capacities from
 *                  capacity.GetNodeCapacities (LIB/capacity/capacity.go:78-102) with `reserved_*` taken off each
 *                  node (the reference passes its overhead map there, :682); a node already hosting executors of
 *                  the same application wins, then the smallest capacity >= 1, then the order.
 * The caller uploads with gp_set_snapshot the availability the reference would use for that branch
 * (availableResources of :643 for first fit, availableNodesSchedulingMetadata of :640 for min_frag). */
typedef struct {
    int32_t n_execs;
    const int64_t* exe_cpu_milli;        /* [n_execs] */
    const int64_t* exe_mem_bytes;        /* [n_execs] */
    const int64_t* exe_gpu;              /* [n_execs] or NULL (= 0) */
    const int32_t* group;                /* [n_execs] instance group of the executor's application, or NULL (= 0) */
    int32_t min_frag;
    /* min_frag only: */
    const int64_t* reserved_cpu_milli;   /* [n_nodes] or NULL (= 0) */
    const int64_t* reserved_mem_bytes;   /* [n_nodes] (required when reserved_cpu_milli is given) */
    const int64_t* reserved_gpu;         /* [n_nodes] or NULL (= 0) */
    const int64_t* host_off;             /* [n_execs + 1] CSR offsets into host_nodes, or NULL (no application hosts anything yet) */
    const int32_t* host_nodes;           /* node indices already hosting executors of the same application
                                            (getNodesWithExecutorsBelongingToSameApp, :683) */
} gp_reschedule;
/* node_out[i] = node index, or -1 ("not enough capacity to reschedule the executor", :672) */
gp_status gp_reschedule_executors(gp_ctx* ctx, const gp_reschedule* in, int32_t* node_out /* [n_execs] */);

/* ---- packing ------------------------------------------------------------------------------- */
/* One batch through the hot path with HOST buffers: H2D of the app SoA, kernels, D2H of the
 * results; returns when the results are in host memory.  FIFO modes mutate the device snapshot. */
gp_status gp_pack_batch(gp_ctx* ctx, const gp_apps* apps, gp_algo algo, gp_mode mode, gp_results* out);

/* ---- compact wire formats for the PCIe-bound host path ------------------------------------------------------------
 * The reference hands a packer six resource.Quantity values and a count per application (types.SparkApplicationResources,
 * internal/types/types.go:22-27) and reads ExecutorNodes back as strings (binpack.go:25-30); what crosses PCIe here is
 * chosen per batch by the shim:
 *   inputs   quantity_bits 64: int64 columns exactly as in gp_apps (60 B per application with offsets);
 *            quantity_bits 32: int32 columns -- millicores, (bytes >> mem_shift), gpu units -- 28 B per application.  Only
 *            when every value is exactly representable (memory a whole multiple of 2^mem_shift bytes, everything < 2^31);
 *            the library rebuilds the exact int64 quantities, so results are bit-identical to the 64-bit layout.
 *   offsets  exec_out_off NULL: the exclusive prefix sum of exe_count is derived on the device (the caller recomputes
 *            it with a running sum while it walks the results).
 *   results  node_bits 16: ExecutorNodes as uint16 node indices -- allowed when the node table has <= 65 535 entries;
 *            GP_MODE_INDEPENDENT with tightly-pack / distribute-evenly only (GP_ERR_INVALID otherwise). */
typedef struct {
    int32_t n_apps;
    int32_t quantity_bits;           /* 64 or 32 */
    int32_t mem_shift;               /* 32-bit layout only: 0..40 */
    int32_t reserved;
    const void* drv_cpu;             /* [n_apps] int64 or int32, as quantity_bits says */
    const void* drv_mem;
    const void* drv_gpu;             /* or NULL (= 0) */
    const void* exe_cpu;
    const void* exe_mem;
    const void* exe_gpu;             /* or NULL (= 0) */
    const int32_t* exe_count;        /* as in gp_apps */
    const int32_t* group;            /* or NULL */
    const uint8_t* skip_if_no_fit;   /* or NULL */
    const int64_t* exec_out_off;     /* [n_apps+1] or NULL (derived on the device) */
} gp_apps_wire;
typedef struct {
    int32_t* driver_node;            /* [n_apps] as in gp_results */
    void* executor_nodes;            /* int32[executor_nodes_cap] (node_bits 32) or uint16[executor_nodes_cap] (node_bits 16) */
    int64_t executor_nodes_cap;
    int32_t node_bits;               /* 32 or 16 */
    int32_t reserved;
} gp_results_wire;
gp_status gp_pack_batch_wire(gp_ctx* ctx, const gp_apps_wire* apps, gp_algo algo, gp_mode mode, gp_results_wire* out);

/* binpack.SparkBinPackFunction for one application (group 0 of the snapshot).
 * Returns GP_OK and *has_capacity; executor_nodes must hold exe_count entries. */
gp_status gp_pack_one(gp_ctx* ctx, gp_algo algo,
                      int64_t drv_cpu_milli, int64_t drv_mem_bytes, int64_t drv_gpu,
                      int64_t exe_cpu_milli, int64_t exe_mem_bytes, int64_t exe_gpu,
                      int32_t exe_count, int32_t* has_capacity, int32_t* driver_node, int32_t* executor_nodes);

/* ---- device-resident entry points (multi-GPU plumbing, kernel-only timing) ----------------- */
/* Same as gp_set_snapshot / gp_pack_batch but every array pointer inside gp_nodes / gp_apps /
 * gp_results is a DEVICE pointer on ctx's device (group, skip_if_no_fit and -- for independent tightly-pack /
 * distribute-evenly -- exec_out_off may be NULL) and nothing is copied or validated on the host.  Work is
 * enqueued on `stream` (a cudaStream_t; NULL = the context's stream) and NOT synchronised (the two
 * order lengths are passed by value so that no device->host read is needed).
 * gp_set_snapshot_device keeps no reference to the caller's arrays after it returns. */
gp_status gp_set_snapshot_device(gp_ctx* ctx, const gp_nodes* dev_nodes,
                                 int32_t n_exec_total /* = exec_off[n_groups] */,
                                 int32_t n_drv_total /* = drv_off[n_groups] */, void* stream);
gp_status gp_pack_batch_device(gp_ctx* ctx, const gp_apps* dev_apps, gp_algo algo, gp_mode mode,
                               gp_results* dev_out, void* stream);
/* the context's own stream (cudaStream_t) so callers can order their work / record events on it */
void* gp_stream(gp_ctx* ctx);
gp_status gp_synchronize(gp_ctx* ctx);

/* ---- single-AZ packers on the device -----------------------------------------------------------
 * getSingleAZSparkBinFunction + chooseBestResult (LIB/binpack/single_az.go:23-97) for a BATCH of applications.  The
 * snapshot's instance groups are the candidate zones in driverZonesInOrder order (groupNodesByZone, :57-73; zones without
 * executor candidates are left out by the caller like :38-41).  Every application is packed in every zone
 * (algo GP_TIGHTLY_PACK = "single-az-tightly-pack", GP_MINIMAL_FRAGMENTATION = "single-az-minimal-fragmentation"), the
 * float64 packing efficiencies (LIB/binpack/efficiency.go:66-156) are evaluated in the reference's operation order --
 * [driver] + ExecutorNodes, duplicates kept, sequential sums, cpu through Quantity.Value() -- and the zone with the highest
 * AvgPackingEfficiency.Max wins: the first among equals, and none when every average is 0 (:80, :91).  For
 * minimal-fragmentation the efficiencies see the driver only (minimalFragmentation never adds its executors to
 * SparkBinPack's reserved map) -- kept.  "az-aware-tightly-pack" (az_aware_pack_tightly.go:27-38) = this call, then plain
 * gp_pack_batch(GP_TIGHTLY_PACK) over the undivided orders for the applications that got zone -1.
 * gp_set_schedulable: NodeSchedulingMetadata.SchedulableResources of the current snapshot's nodes (resources.go:61-100). */
gp_status gp_set_schedulable(gp_ctx* ctx, const int64_t* sched_cpu_milli, const int64_t* sched_mem_bytes, const int64_t* sched_gpu /* or NULL */);
typedef struct {
    int32_t* zone;                   /* [n_apps] chosen instance group (zone), -1 = EmptyPackingResult */
    int32_t* driver_node;            /* [n_apps] */
    int32_t* executor_nodes;         /* CSR by exec_out_off (or the prefix sum of exe_count); defined when zone >= 0 */
    int64_t executor_nodes_cap;
    double* avg_efficiency;          /* [n_apps][4] = AvgPackingEfficiency{CPU, Memory, GPU, Max} of the winner, or NULL */
} gp_zone_results;
gp_status gp_pack_batch_zones(gp_ctx* ctx, const gp_apps* apps /* group, skip_if_no_fit ignored */, gp_algo algo, gp_zone_results* out);
/* The FIFO loop with a single-AZ packer in ONE launch: fitEarlierDrivers (internal/extender/resource.go:224-262) when
 * binpacker.BinpackFunc is SingleAZTightlyPack / SingleAZMinimalFragmentation.  The applications are the queue in order;
 * each one is packed in every zone against the availability its predecessors left, the zone is chosen like above, the
 * winner's usage is subtracted (mode GP_MODE_FIFO_REFERENCE: sparkResourceUsage's map assignment, sparkpods.go:139-146;
 * GP_MODE_FIFO_EXACT: every pod) and the next application sees it.  An application that fits in no zone gets zone -1,
 * driver_node -1 and -- unless skip_if_no_fit[i] -- blocks the queue: everything behind it reports driver_node -2
 * (never evaluated, resource.go:244-253).  The availability the device holds afterwards is what the loop left (read it
 * with gp_get_snapshot).  At most 64 zones; exec_out_off must be NULL or the prefix sum of exe_count.
 * ("az-aware-tightly-pack" needs the undivided orders for its fallback and stays a per-driver call sequence.) */
gp_status gp_pack_fifo_zones(gp_ctx* ctx, const gp_apps* apps /* group ignored */, gp_algo algo, gp_mode mode, gp_zone_results* out);

/* ---- reservation table of a batch + a snapshot that stays on the device ---------------
 * gp_reserve_placements: newResourceReservation (internal/extender/resourcereservations.go:491-528) for every application
 * of a packed batch whose driver_node >= 0, as a flat SoA table the shim turns into ResourceReservation objects: one row
 * per reservation slot -- slot 0 = "driver", slot i = "executor-i" (executorReservationName, :530-533) -> node index and
 * the slot's resources; rows of one application are contiguous, applications in queue order.  With subtract_from_snapshot
 * != 0 the same kernel takes every reserved pod off the availability the device holds (exact accounting, what
 * UsageForNodes, LIB/resources/resources.go:31-43, adds up from these reservations on the next Predicate), so the next
 * gp_pack_* call needs no gp_set_snapshot.
 * gp_apply_usage_delta: reservations that appeared (sign = +1: availability goes down) or went away (sign = -1) since
 * the snapshot was laid out -- the incremental form of GetReservedResources (resourcereservations.go:258-263). */
typedef struct {
    int64_t rows_cap;                /* rows the arrays below can hold; sum over placed apps of 1 + exe_count */
    int32_t* app;                    /* [rows_cap] index of the application in the batch */
    int32_t* slot;                   /* 0 = "driver", i = "executor-i" */
    int32_t* node;                   /* Reservation.Node as node index */
    int64_t* cpu_milli;              /* Reservation.Resources */
    int64_t* mem_bytes;
    int64_t* gpu;                    /* or NULL (not returned) */
    int64_t n_rows;                  /* out */
} gp_reservation_table;
gp_status gp_reserve_placements(gp_ctx* ctx, const gp_apps* apps, const gp_results* placed /* host arrays of a previous gp_pack_batch */,
                                int32_t subtract_from_snapshot, gp_reservation_table* out /* may be NULL with subtract */);
gp_status gp_apply_usage_delta(gp_ctx* ctx, int64_t n_rows, const int32_t* node, const int64_t* cpu_milli, const int64_t* mem_bytes,
                               const int64_t* gpu /* or NULL */, int32_t sign /* +1 reserve, -1 release */);

/* ---- several GPUs of one box behind ONE host process -----------------------------------------------
 * The reference is a single process with a serial Predicate (internal/extender/resource.go:194-205); this handle lets
 * that process use every GPU: one gp_ctx + one worker thread per entry of `devices` (an ordinal may repeat).
 *   gp_multi_set_snapshot   every device receives the full snapshot by a direct H2D copy over its own PCIe link.
 *   gp_multi_pack_batch     GP_MODE_INDEPENDENT: the queue is cut into one contiguous block per device and every device
 *                           copies ITS placements straight into the caller's result buffers (no gather through one GPU,
 *                           no collective: the path has no exchange step).  FIFO modes: whole instance groups per device
 *                           (queues of different groups are independent: internal/extender/sparkpods.go:61,
 *                           resource.go:292-295; longest-processing-time first on applications x nodes), results are
 *                           scattered back into queue order; exec_out_off is mandatory and node_bits must be 32 there.
 *                           Results are bit-identical to gp_pack_batch_wire on one context.
 *   gp_multi_get_snapshot   after a FIFO batch: every node from the device that owned its instance group.
 *   gp_multi_group_owner    which device ran each instance group's queue in the last FIFO batch. */
typedef struct gp_multi gp_multi;
gp_status gp_multi_create(gp_multi** out, const int32_t* devices /* NULL = 0..n-1 */, int32_t n_devices);
void gp_multi_destroy(gp_multi* m);
const char* gp_multi_last_error(const gp_multi* m);
int32_t gp_multi_size(const gp_multi* m);
gp_ctx* gp_multi_ctx(gp_multi* m, int32_t i);      /* the i-th context, e.g. for gp_alloc_pinned / gp_last_stats */
gp_status gp_multi_set_snapshot(gp_multi* m, const gp_nodes* nodes);
gp_status gp_multi_pack_batch(gp_multi* m, const gp_apps_wire* apps, gp_algo algo, gp_mode mode, gp_results_wire* out);
gp_status gp_multi_get_snapshot(gp_multi* m, int64_t* avail_cpu_milli, int64_t* avail_mem_bytes, int64_t* avail_gpu);
gp_status gp_multi_group_owner(gp_multi* m, int32_t* owner /* [n_groups] */);

/* Statistics of the last gp_pack_batch* on this context (host call, synchronises the stream):
 * nodes_scanned = executor-order entries visited, drivers_tried = driver-order entries visited,
 * summed over apps -- the N_e / N_d of the algorithmic-bytes formula. */
typedef struct {
    int64_t nodes_scanned;
    int64_t drivers_tried;
    int64_t kernel_launches;   /* launches of this library's kernels in that call */
    int64_t pack_kernel_ns;    /* device time of the pack kernel (CUDA events on the launch stream) */
    int64_t prep_kernel_ns;    /* device time of the kernels before it (preparation / shape classification + capacity tables) */
    int64_t scan_path_apps;    /* independent tightly/evenly: applications decided by the node-order scan instead of the
                                  per-shape capacity tables (all of them with GANGPACK_TABLES=0) */
    int64_t scan_path_nodes;   /* executor-order entries those scans visited (nodes_scanned also counts table probes) */
    int64_t reserved[1];
} gp_stats;
gp_status gp_last_stats(gp_ctx* ctx, gp_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* GANGPACK_H */

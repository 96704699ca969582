/*
 * oracle/gangpack_oracle.h -- CPU restatement of the reference gang-placement hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under k8s-spark-scheduler_b200/ (the product) may
 * include, link or execute this code; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS: partially pinned.  Pinned by the reference's own tests / worked examples (tests/golden/, asserted in
 * tests/test_oracle_golden.py): fit / no-fit for the harness app shapes and min-executor semantics
 * (EXT/resource_test.go, unschedulablepods_test.go), node-priority orders and label priorities
 * (internal/sort/nodesorting_test.go), the node chosen by rescheduleExecutorWithMinimalFragmentation
 * (resource_test.go:73-165), ExecutorNodes of minimalFragmentation for the four doc-comment examples that agree with its
 * code (LIB/binpack/minimal_fragmentation.go:43-55), annotation -> tuple parsing and FIFO queue membership/order
 * (EXT/sparkpods_test.go).  NOT pinned by any test or golden vector in /root/reference (the lib's *_test.go files are
 * not vendored) and Go cannot run here: the exact contents/order of ExecutorNodes for tightly-pack / distribute-evenly
 * and the FIFO subtraction -> for those outputs: "parity unpinned", authority is the cited source lines; three
 * independent restatements (literal C, closed-form C, pure Python) must agree on them.  tests/golden/ also holds
 * hand-derived vectors, re-derived by the pure-Python restatement.
 *
 * Two restatements live here:
 *   literal  (gangpack_oracle.c) -- loop-for-loop, string-keyed maps like the Go code.
 *   closed   (gangpack_closed.c) -- int64 SoA closed form (per-node capacity), cross-checked
 *                                   against the literal one in tests/.
 *
 * Path abbreviations: LIB = vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg,
 *                     EXT = internal/extender   (all under /root/reference).
 *
 * Quantity model: CPU in millicores, memory in bytes, GPU in units, all int64.
 */
#ifndef GANGPACK_ORACLE_H
#define GANGPACK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* LIB/resources/resources.go:151-155 */
typedef struct { int64_t cpu, mem, gpu; } orc_res;

/* 2, 3: LIB/binpack/single_az_pack_tightly.go + single_az.go:23-97, az_aware_pack_tightly.go:27-38 (literal oracle only) */
/* 4, 5: LIB/binpack/minimal_fragmentation.go:27-137 + LIB/capacity/capacity.go, single_az_minimal_fragmentation.go:20
 *       (4 = MinimalFragmentation is also restated in closed form; 5 literal only) */
enum { ORC_TIGHTLY_PACK = 0, ORC_DISTRIBUTE_EVENLY = 1, ORC_SINGLE_AZ_TIGHTLY_PACK = 2, ORC_AZ_AWARE_TIGHTLY_PACK = 3,
       ORC_MINIMAL_FRAGMENTATION = 4, ORC_SINGLE_AZ_MINIMAL_FRAGMENTATION = 5 };
/* FIFO accounting: 1 = the reference's sparkResourceUsage overwrite (EXT/sparkpods.go:139-146),
 *                  2 = exact sum (what separate Predicate calls converge to via UsageForNodes). */
enum { ORC_FIFO_REFERENCE = 1, ORC_FIFO_EXACT = 2 };

/* ------------------------------------------------------------------ literal ---- */

/* NodeGroupSchedulingMetadata: map[string]*NodeSchedulingMetadata (LIB/resources/resources.go:158-166,121) */
typedef struct orc_cluster orc_cluster;

/* sched_* may be NULL (then schedulable := available and efficiencies are meaningless).
 * zone / unschedulable / ready / may be NULL (single zone "default", schedulable, ready). */
orc_cluster* orc_cluster_new(int32_t n, const char* const* names,
                             const int64_t* avail_cpu, const int64_t* avail_mem, const int64_t* avail_gpu,
                             const int64_t* sched_cpu, const int64_t* sched_mem, const int64_t* sched_gpu,
                             const char* const* zone, const uint8_t* unschedulable, const uint8_t* ready);
void orc_cluster_free(orc_cluster*);
int32_t orc_cluster_size(const orc_cluster*);
/* node index (insertion order) of a name, -1 if absent */
int32_t orc_cluster_index(const orc_cluster*, const char* name);
/* copy out AvailableResources in insertion order */
void orc_cluster_get_available(const orc_cluster*, int64_t* cpu, int64_t* mem, int64_t* gpu);

/* binpack.SparkBinPackFunction (LIB/binpack/binpack.go:43-48) for TightlyPack / DistributeEvenly.
 * Orders are node NAMES (may name nodes absent from the cluster, may be empty).
 * Returns HasCapacity.  driver_node: cluster index or -1.  executor_nodes[count]: cluster
 * indices in ExecutorNodes order (untouched when !HasCapacity).
 * avg_eff (may be NULL): {CPU, Memory, GPU, Max} of computeAvgPackingEfficiencyForResult
 * (EXT/resource.go:372-381) summed in cluster insertion order (Go: map order).
 * with_efficiencies: run ComputePackingEfficiencies (LIB/binpack/efficiency.go:66-110) on success
 * like the reference does (costs O(N) per successful pack). */
int orc_binpack(const orc_cluster*, int algo, const orc_res* drv, const orc_res* exe, int32_t count,
                const char* const* driver_order, int32_t n_driver,
                const char* const* exec_order, int32_t n_exec,
                int with_efficiencies,
                int32_t* driver_node, int32_t* executor_nodes, double* avg_eff);

/* Independent batch: app i packed against the same (unmodified) cluster.
 * exec_off[n_apps+1] = exclusive prefix sum of count[]; executor_nodes sized exec_off[n_apps].
 * driver_node[i] = index or -1.  n_threads > 1 partitions apps over pthreads. */
void orc_binpack_batch(const orc_cluster*, int algo, int32_t n_apps,
                       const orc_res* drv, const orc_res* exe, const int32_t* count,
                       const char* const* driver_order, int32_t n_driver,
                       const char* const* exec_order, int32_t n_exec,
                       int with_efficiencies, int n_threads,
                       const int64_t* exec_off, int32_t* driver_node, int32_t* executor_nodes);

/* fitEarlierDrivers (EXT/resource.go:224-262) over apps[0..n_apps) in queue order against the
 * MUTATING cluster: pack with count[i]; on fit subtract sparkResourceUsage
 * (EXT/sparkpods.go:139-146) via SubtractUsageIfExists (LIB/resources/resources.go:129-135);
 * on no-fit: young[i] ? skip : stop.  driver_node[i] = index / -1 (no fit) / -2 (not evaluated).
 * Returns index of the blocking app or -1.  (The caller's own final pack, EXT/resource.go:321,
 * is simply the last app of the queue.) */
int32_t orc_fifo(orc_cluster*, int algo, int mode, int32_t n_apps,
                 const orc_res* drv, const orc_res* exe, const int32_t* count, const uint8_t* young,
                 const char* const* driver_order, int32_t n_driver,
                 const char* const* exec_order, int32_t n_exec,
                 int with_efficiencies,
                 const int64_t* exec_off, int32_t* driver_node, int32_t* executor_nodes);

/* NodeSorter.PotentialNodes (internal/sort/nodesorting.go:41-64) incl. label-priority re-sort
 * (:161-200).  candidate_names = kube-scheduler's NodeNames.  *_label_rank[i]: rank of node i's
 * value for the configured label, or -1 if the label/value is unknown; NULL = no label config.
 * Ties the reference leaves to an unstable sort are broken stably here.
 * Outputs cluster indices; returns counts through n_driver_out / n_exec_out. */
void orc_potential_nodes(const orc_cluster*, const char* const* candidate_names, int32_t n_candidates,
                         const int32_t* driver_label_rank, const int32_t* exec_label_rank,
                         int32_t* driver_out, int32_t* n_driver_out,
                         int32_t* exec_out, int32_t* n_exec_out);

/* rescheduleExecutor's node choice (EXT/resource.go:594-705) for one executor: first fit over the
 * executor order (min_frag == 0) or rescheduleExecutorWithMinimalFragmentation (min_frag != 0; reserved_* = the
 * overhead map it passes to GetNodeCapacities, hosting_names = nodes already hosting executors of the same
 * application).  Returns the cluster index of the chosen node or -1. */
int32_t orc_reschedule_executor(const orc_cluster*, int min_frag, const orc_res* exe,
                                const char* const* exec_order, int32_t n_exec,
                                const char* const* reserved_names, const orc_res* reserved, int32_t n_reserved,
                                const char* const* hosting_names, int32_t n_hosting);

/* Snapshot build: GetReservedResources (EXT/resourcereservations.go:258-263: UsageForNodes over
 * the hard reservations, LIB/resources/resources.go:31-43, plus the soft reservations,
 * internal/cache/softreservations.go:155-170) and NodeSchedulingMetadataForNodes (resources.go:61-100).
 * Reservations name their node; names outside the node list are ignored.  sched_* may be NULL. */
void orc_node_scheduling_metadata(int32_t n_nodes, const char* const* names,
                                  const int64_t* alloc_cpu, const int64_t* alloc_mem, const int64_t* alloc_gpu,
                                  const int64_t* over_cpu, const int64_t* over_mem, const int64_t* over_gpu,
                                  int64_t n_res, const char* const* res_node_name,
                                  const int64_t* res_cpu, const int64_t* res_mem, const int64_t* res_gpu,
                                  int64_t* avail_cpu, int64_t* avail_mem, int64_t* avail_gpu,
                                  int64_t* sched_cpu, int64_t* sched_mem, int64_t* sched_gpu);

/* availableResources of rescheduleExecutor's first-fit branch (EXT/resource.go:638-643) with the overhead double count of
 * SURVEY App. B7 (nodes that carry a reservation lose their overhead twice). */
void orc_reschedule_available(int32_t n_nodes, const char* const* names,
                              const int64_t* alloc_cpu, const int64_t* alloc_mem, const int64_t* alloc_gpu,
                              const int64_t* over_cpu, const int64_t* over_mem, const int64_t* over_gpu,
                              int64_t n_res, const char* const* res_node_name,
                              const int64_t* res_cpu, const int64_t* res_mem, const int64_t* res_gpu,
                              int64_t* avail_cpu, int64_t* avail_mem, int64_t* avail_gpu);

/* ------------------------------------------------------------------ closed form ---- */

/* Same semantics on index arrays: node table [n_nodes] (int64 SoA), exec_order / driver_order are
 * indices into it; an index < 0 or >= n_nodes models "name not in metadata".
 * Mutates avail_* in FIFO modes (mode 0 = independent). */
int32_t orc_closed_batch(int algo, int mode, int32_t n_nodes,
                         int64_t* avail_cpu, int64_t* avail_mem, int64_t* avail_gpu,
                         const int32_t* driver_order, int32_t n_driver,
                         const int32_t* exec_order, int32_t n_exec,
                         int32_t n_apps, const orc_res* drv, const orc_res* exe,
                         const int32_t* count, const uint8_t* young, int n_threads,
                         const int64_t* exec_off, int32_t* driver_node, int32_t* executor_nodes);

#ifdef __cplusplus
}
#endif
#endif

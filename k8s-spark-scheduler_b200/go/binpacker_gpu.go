//go:build cgo

// Package binpackergpu is the cgo shim a maintainer of palantir/k8s-spark-scheduler adds to route the
// placement hot path through libgangpack.so (include/gangpack.h).
//
// NOT COMPILED OR TESTED HERE: this environment has no Go toolchain.  The same marshalling, written in
// C++ and exercised on a B200, is k8s-spark-scheduler_b200/host/gangpack_host.hpp; this file is kept
// deliberately thin so that the two can be compared line by line.  See INTEGRATION.md.
//
// Build: CGO_ENABLED=1 (the reference builds with CGO_ENABLED=0, godel/config/dist-plugin.yml:7,26),
//
//	CGO_CFLAGS=-I<repo>/include  CGO_LDFLAGS="-L<repo>/k8s-spark-scheduler_b200 -lgangpack"
package binpackergpu

/*
#include <stdlib.h>
#include "gangpack.h"
*/
import "C"

import (
	"context"
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"github.com/palantir/k8s-spark-scheduler-lib/pkg/binpack"
	"github.com/palantir/k8s-spark-scheduler-lib/pkg/resources"
	"k8s.io/apimachinery/pkg/api/resource"
)

// Names under which the GPU packers are registered next to the reference's own entries in
// internal/binpacker.binpackFunctions (internal/binpacker/binpack.go:43-49).  Registering them under
// the existing names "tightly-pack" / "distribute-evenly" instead makes cmd/server.go:145 pick them
// with an unchanged install.yml.
const (
	TightlyPackGPU      = "tightly-pack-gpu"
	DistributeEvenlyGPU = "distribute-evenly-gpu"
)

// device owns one gp_ctx; Predicate is serial (internal/extender/resource.go:194-205) but the
// unschedulable-pod marker calls BinpackFunc from its own goroutine (cmd/server.go:230), hence the mutex.
type device struct {
	mu  sync.Mutex
	ctx *C.gp_ctx
}

var (
	dev     *device
	devOnce sync.Once
	devErr  error
)

func getDevice() (*device, error) {
	devOnce.Do(func() {
		var ctx *C.gp_ctx
		if st := C.gp_create(&ctx, nil); st != C.GP_OK {
			devErr = fmt.Errorf("gp_create: %s", C.GoString(C.gp_last_error(nil)))
			return
		}
		dev = &device{ctx: ctx}
	})
	return dev, devErr
}

// milli converts a Quantity to the exact-int64 model; ok=false routes the call to the Go packer
// (Quantity falls back to inf.Dec beyond int64 / for sub-milli scales, quantity.go:556-591).
func milli(q resource.Quantity) (int64, bool) {
	if q.Cmp(*resource.NewMilliQuantity(q.MilliValue(), q.Format)) != 0 {
		return 0, false
	}
	return q.MilliValue(), true
}
func whole(q resource.Quantity) (int64, bool) {
	v, ok := q.AsInt64()
	return v, ok
}

func toTriple(r *resources.Resources) (cpu, mem, gpu int64, ok bool) {
	var a, b, c bool
	cpu, a = milli(r.CPU)
	mem, b = whole(r.Memory)
	gpu, c = whole(r.NvidiaGPU)
	return cpu, mem, gpu, a && b && c
}

// snapshot marshals the metadata map and the two priority orders into SoA buffers.  Names that are
// not in the metadata are dropped: they can host neither a driver (binpack.go:68-69) nor an executor
// (pack_tightly.go:51-52).
type snapshot struct {
	names            []string
	cpu, mem, gpu    []int64
	execIdx, drvIdx  []int32
}

func marshal(md resources.NodeGroupSchedulingMetadata, driverOrder, executorOrder []string) (*snapshot, bool) {
	s := &snapshot{}
	index := make(map[string]int32, len(md))
	intern := func(n string) (int32, bool, bool) {
		if i, ok := index[n]; ok {
			return i, true, true
		}
		m, ok := md[n]
		if !ok {
			return -1, false, true
		}
		c, mm, g, exact := toTriple(m.AvailableResources)
		if !exact {
			return -1, false, false
		}
		i := int32(len(s.names))
		index[n] = i
		s.names = append(s.names, n)
		s.cpu, s.mem, s.gpu = append(s.cpu, c), append(s.mem, mm), append(s.gpu, g)
		return i, true, true
	}
	for _, n := range executorOrder {
		i, present, exact := intern(n)
		if !exact {
			return nil, false
		}
		if present {
			s.execIdx = append(s.execIdx, i)
		}
	}
	for _, n := range driverOrder {
		i, present, exact := intern(n)
		if !exact {
			return nil, false
		}
		if present {
			s.drvIdx = append(s.drvIdx, i)
		}
	}
	return s, true
}

func ptr64(v []int64) *C.int64_t {
	if len(v) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&v[0]))
}
func ptr32(v []int32) *C.int32_t {
	if len(v) == 0 {
		return nil
	}
	return (*C.int32_t)(unsafe.Pointer(&v[0]))
}

func (d *device) setSnapshot(s *snapshot) error {
	eoff := []int32{0, int32(len(s.execIdx))}
	doff := []int32{0, int32(len(s.drvIdx))}
	var pin runtime.Pinner // Go >= 1.21: the struct holds Go pointers for the duration of the call
	defer pin.Unpin()
	for _, p := range []any{ptrOrNil(s.cpu), ptrOrNil(s.mem), ptrOrNil(s.gpu), ptrOrNil32(s.execIdx), ptrOrNil32(s.drvIdx), &eoff[0], &doff[0]} {
		if p != nil {
			pin.Pin(p)
		}
	}
	n := C.gp_nodes{
		n_nodes: C.int32_t(len(s.names)), avail_cpu_milli: ptr64(s.cpu), avail_mem_bytes: ptr64(s.mem), avail_gpu: ptr64(s.gpu),
		n_groups: 1, exec_off: ptr32(eoff), exec_order: ptr32(s.execIdx), drv_off: ptr32(doff), drv_order: ptr32(s.drvIdx),
	}
	if st := C.gp_set_snapshot(d.ctx, &n); st != C.GP_OK {
		return fmt.Errorf("gp_set_snapshot: %s", C.GoString(C.gp_last_error(d.ctx)))
	}
	return nil
}

func ptrOrNil(v []int64) any {
	if len(v) == 0 {
		return nil
	}
	return &v[0]
}
func ptrOrNil32(v []int32) any {
	if len(v) == 0 {
		return nil
	}
	return &v[0]
}
func ptrOrNil8(v []uint8) any {
	if len(v) == 0 {
		return nil
	}
	return &v[0]
}
func ptr8(v []uint8) *C.uint8_t {
	if len(v) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&v[0]))
}

// gpuPacker returns a binpack.SparkBinPackFunction (LIB/binpack/binpack.go:43-48) backed by the device;
// any failure (no device, CUDA error, unrepresentable quantity) runs `fallback`, the original Go packer,
// so Predicate never fails because of the accelerator.
func gpuPacker(algo C.gp_algo, fallback binpack.SparkBinPackFunction) binpack.SparkBinPackFunction {
	return func(ctx context.Context, driverResources, executorResources *resources.Resources, executorCount int,
		driverNodePriorityOrder, executorNodePriorityOrder []string,
		nodesSchedulingMetadata resources.NodeGroupSchedulingMetadata) *binpack.PackingResult {
		goPath := func() *binpack.PackingResult {
			return fallback(ctx, driverResources, executorResources, executorCount, driverNodePriorityOrder,
				executorNodePriorityOrder, nodesSchedulingMetadata)
		}
		d, err := getDevice()
		if err != nil {
			return goPath()
		}
		dc, dm, dg, ok1 := toTriple(driverResources)
		ec, em, eg, ok2 := toTriple(executorResources)
		s, ok3 := marshal(nodesSchedulingMetadata, driverNodePriorityOrder, executorNodePriorityOrder)
		if !ok1 || !ok2 || !ok3 || executorCount < 0 {
			return goPath()
		}
		runtime.LockOSThread()
		defer runtime.UnlockOSThread()
		d.mu.Lock()
		defer d.mu.Unlock()
		if err := d.setSnapshot(s); err != nil {
			return goPath()
		}
		nodes := make([]int32, max(executorCount, 1))
		var has, driver C.int32_t
		st := C.gp_pack_one(d.ctx, algo, C.int64_t(dc), C.int64_t(dm), C.int64_t(dg), C.int64_t(ec), C.int64_t(em), C.int64_t(eg),
			C.int32_t(executorCount), &has, &driver, ptr32(nodes))
		if st != C.GP_OK {
			return goPath()
		}
		if has == 0 {
			return binpack.EmptyPackingResult()
		}
		res := &binpack.PackingResult{DriverNode: s.names[driver], HasCapacity: true,
			ExecutorNodes: make([]string, executorCount)}
		for i := 0; i < executorCount; i++ {
			res.ExecutorNodes[i] = s.names[nodes[i]]
		}
		// PackingEfficiencies (metrics/debug log only for these two packers, EXT/resource.go:329-352) are
		// computed on the host from the result when a caller asks for them:
		res.PackingEfficiencies = efficienciesFor(nodesSchedulingMetadata, driverResources, executorResources, res,
			algo != C.GP_MINIMAL_FRAGMENTATION)
		return res
	}
}

// efficienciesFor rebuilds the reserved map of SparkBinPack (binpack.go:72-77) from the placement and
// calls the reference's own ComputePackingEfficiencies.  tightlyPackExecutors / distributeExecutorsEvenly add every
// executor to that map; minimalFragmentation never touches it (withExecutors=false), which is what the reference's
// chooseBestResult then sees -- kept as is.
func efficienciesFor(md resources.NodeGroupSchedulingMetadata, drv, exe *resources.Resources, r *binpack.PackingResult,
	withExecutors bool) map[string]*binpack.PackingEfficiency {
	reserved := resources.NodeGroupResources{r.DriverNode: drv.Copy()}
	if withExecutors {
		for _, n := range r.ExecutorNodes {
			if reserved[n] == nil {
				reserved[n] = resources.Zero()
			}
			reserved[n].Add(exe)
		}
	}
	return binpack.ComputePackingEfficiencies(md, reserved)
}

// TightlyPack / DistributeEvenly / MinimalFragmentation are drop-in values for binpack.TightlyPack /
// binpack.DistributeEvenly / binpack.MinimalFragmentation.
var (
	TightlyPack          = gpuPacker(C.GP_TIGHTLY_PACK, binpack.TightlyPack)
	DistributeEvenly     = gpuPacker(C.GP_DISTRIBUTE_EVENLY, binpack.DistributeEvenly)
	MinimalFragmentation = gpuPacker(C.GP_MINIMAL_FRAGMENTATION, binpack.MinimalFragmentation)
)

// singleAZ is getSingleAZSparkBinFunction + chooseBestResult (LIB/binpack/single_az.go:23-97) over a device packer
// (both are unexported in the lib, hence restated here with its exported efficiency helpers).  The C++ mirror packs all
// zones in ONE device batch (zone = instance group, host/gangpack_host.hpp SingleAZPackImpl); this thin version calls
// the per-zone packer once per zone.
func singleAZ(perZone binpack.SparkBinPackFunction) binpack.SparkBinPackFunction {
	return func(ctx context.Context, driverResources, executorResources *resources.Resources, executorCount int,
		driverNodePriorityOrder, executorNodePriorityOrder []string,
		md resources.NodeGroupSchedulingMetadata) *binpack.PackingResult {
		group := func(names []string) ([]string, map[string][]string) { // groupNodesByZone, :57-73
			order, byZone := []string{}, map[string][]string{}
			for _, n := range names {
				m, ok := md[n]
				if !ok {
					continue
				}
				if _, seen := byZone[m.ZoneLabel]; !seen {
					order = append(order, m.ZoneLabel)
				}
				byZone[m.ZoneLabel] = append(byZone[m.ZoneLabel], n)
			}
			return order, byZone
		}
		zones, dz := group(driverNodePriorityOrder)
		_, ez := group(executorNodePriorityOrder)
		best, bestAvg := binpack.EmptyPackingResult(), binpack.WorstAvgPackingEfficiency() // :79-80
		for _, z := range zones {
			eo, ok := ez[z]
			if !ok {
				continue // :38-41
			}
			r := perZone(ctx, driverResources, executorResources, executorCount, dz[z], eo, md)
			if !r.HasCapacity {
				continue // :44-46
			}
			effs := make([]*binpack.PackingEfficiency, 0, 1+len(r.ExecutorNodes)) // :83-89
			for _, n := range append([]string{r.DriverNode}, r.ExecutorNodes...) {
				effs = append(effs, r.PackingEfficiencies[n])
			}
			if avg := binpack.ComputeAvgPackingEfficiency(md, effs); bestAvg.LessThan(avg) { // :90-94
				best, bestAvg = r, avg
			}
		}
		return best
	}
}

// Drop-in values for binpack.SingleAZTightlyPack / binpack.SingleAZMinimalFragmentation / binpack.AzAwareTightlyPack.
var (
	SingleAZTightlyPack          = singleAZ(TightlyPack)
	SingleAZMinimalFragmentation = singleAZ(MinimalFragmentation)
	AzAwareTightlyPack           = binpack.SparkBinPackFunction(func(ctx context.Context, d, e *resources.Resources, k int,
		dord, eord []string, md resources.NodeGroupSchedulingMetadata) *binpack.PackingResult {
		if r := SingleAZTightlyPack(ctx, d, e, k, dord, eord, md); r.HasCapacity { // az_aware_pack_tightly.go:33-37
			return r
		}
		return TightlyPack(ctx, d, e, k, dord, eord, md)
	})
)

// RescheduleExecutorNode is the node choice of rescheduleExecutor (EXT/resource.go:652-662, 675-705) on the device:
// minFrag selects rescheduleExecutorWithMinimalFragmentation (`md` = availableNodesSchedulingMetadata of :640, `overhead`
// the map it hands to GetNodeCapacities, `hosting` = getNodesWithExecutorsBelongingToSameApp); otherwise the first
// node of the order that fits (`md` then carries availableResources of :643).  ok=false: run the Go code instead.
func RescheduleExecutorNode(minFrag bool, executorResources *resources.Resources, executorNodeNames []string,
	md resources.NodeGroupSchedulingMetadata, overhead resources.NodeGroupResources, hosting map[string]bool) (node string, found, ok bool) {
	d, err := getDevice()
	if err != nil {
		return "", false, false
	}
	ec, em, eg, ok1 := toTriple(executorResources)
	s, ok2 := marshal(md, nil, executorNodeNames)
	if !ok1 || !ok2 || len(s.names) == 0 {
		return "", false, ok1 && ok2
	}
	index := make(map[string]int32, len(s.names))
	for i, n := range s.names {
		index[n] = int32(i)
	}
	rc, rm, rg := make([]int64, len(s.names)), make([]int64, len(s.names)), make([]int64, len(s.names))
	hostNodes, hostOff := []int32{}, []int64{0, 0}
	in := C.gp_reschedule{n_execs: 1, exe_cpu_milli: (*C.int64_t)(unsafe.Pointer(&ec)), exe_mem_bytes: (*C.int64_t)(unsafe.Pointer(&em)),
		exe_gpu: (*C.int64_t)(unsafe.Pointer(&eg))}
	if minFrag {
		for n, r := range overhead {
			if i, known := index[n]; known {
				var exact bool
				if rc[i], rm[i], rg[i], exact = toTriple(r); !exact {
					return "", false, false
				}
			}
		}
		for n, yes := range hosting {
			if i, known := index[n]; known && yes {
				hostNodes = append(hostNodes, i)
			}
		}
		hostOff[1] = int64(len(hostNodes))
		in.min_frag = 1
		in.reserved_cpu_milli, in.reserved_mem_bytes, in.reserved_gpu = ptr64(rc), ptr64(rm), ptr64(rg)
		in.host_off, in.host_nodes = ptr64(hostOff), ptr32(hostNodes)
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	d.mu.Lock()
	defer d.mu.Unlock()
	if err := d.setSnapshot(s); err != nil {
		return "", false, false
	}
	var out C.int32_t = -1
	var pin runtime.Pinner
	defer pin.Unpin()
	for _, p := range []any{&ec, &em, &eg, ptrOrNil(rc), ptrOrNil(rm), ptrOrNil(rg), &hostOff[0], ptrOrNil32(hostNodes)} {
		if p != nil {
			pin.Pin(p)
		}
	}
	if st := C.gp_reschedule_executors(d.ctx, &in, &out); st != C.GP_OK {
		return "", false, false
	}
	if out < 0 {
		return "", false, true // failureFit, :672
	}
	return s.names[out], true, true
}

// QueuedApp is what fitEarlierDrivers reads from one earlier driver pod (EXT/resource.go:230-243).
type QueuedApp struct {
	Driver, Executor *resources.Resources
	MinExecutorCount int
	SkipIfNoFit      bool // shouldSkipDriverFifo (EXT/resource.go:264-270)
}

// FitEarlierDriversBatch replaces the body of fitEarlierDrivers (EXT/resource.go:224-262): ONE cgo call
// packs the whole queue in order on the device against the mutating snapshot (GP_MODE_FIFO_REFERENCE keeps
// the sparkResourceUsage accounting, EXT/sparkpods.go:139-146) and writes the charged availability back
// into metadata, as SubtractUsageIfExists (LIB/resources/resources.go:129-135) would have.
// ok=false: the caller must run the original Go loop instead.
func FitEarlierDriversBatch(algo C.gp_algo, apps []QueuedApp, nodeNames, executorNodeNames []string,
	metadata resources.NodeGroupSchedulingMetadata) (fits bool, ok bool) {
	d, err := getDevice()
	if err != nil {
		return false, false
	}
	if len(apps) == 0 {
		return true, true // no earlier drivers: nothing to fit, nothing to subtract (EXT/resource.go:230 loop body never runs)
	}
	s, exact := marshal(metadata, nodeNames, executorNodeNames)
	if !exact {
		return false, false
	}
	q := len(apps)
	dc, dm, dg := make([]int64, q), make([]int64, q), make([]int64, q)
	ec, em, eg := make([]int64, q), make([]int64, q), make([]int64, q)
	cnt, skip, off := make([]int32, q), make([]uint8, q), make([]int64, q+1)
	for i, a := range apps {
		var o1, o2 bool
		dc[i], dm[i], dg[i], o1 = toTriple(a.Driver)
		ec[i], em[i], eg[i], o2 = toTriple(a.Executor)
		if !o1 || !o2 || a.MinExecutorCount < 0 {
			return false, false
		}
		cnt[i] = int32(a.MinExecutorCount)
		if a.SkipIfNoFit {
			skip[i] = 1
		}
		off[i+1] = off[i] + int64(cnt[i])
	}
	driver := make([]int32, q)
	exec := make([]int32, max(int(off[q]), 1))
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	d.mu.Lock()
	defer d.mu.Unlock()
	if err := d.setSnapshot(s); err != nil {
		return false, false
	}
	// gp_apps / gp_results hold Go pointers: every slice base is pinned for the duration of the call
	// (cgocheck: "Go pointer to unpinned Go pointer"), exactly as setSnapshot does for the node table.
	var pin runtime.Pinner
	defer pin.Unpin()
	for _, p := range []any{ptrOrNil(dc), ptrOrNil(dm), ptrOrNil(dg), ptrOrNil(ec), ptrOrNil(em), ptrOrNil(eg),
		ptrOrNil32(cnt), ptrOrNil8(skip), ptrOrNil(off), ptrOrNil32(driver), ptrOrNil32(exec)} {
		if p != nil {
			pin.Pin(p)
		}
	}
	ga := C.gp_apps{n_apps: C.int32_t(q), drv_cpu_milli: ptr64(dc), drv_mem_bytes: ptr64(dm), drv_gpu: ptr64(dg),
		exe_cpu_milli: ptr64(ec), exe_mem_bytes: ptr64(em), exe_gpu: ptr64(eg), exe_count: ptr32(cnt),
		skip_if_no_fit: ptr8(skip), exec_out_off: ptr64(off)}
	gr := C.gp_results{driver_node: ptr32(driver), executor_nodes: ptr32(exec), executor_nodes_cap: C.int64_t(len(exec))}
	if st := C.gp_pack_batch(d.ctx, &ga, algo, C.GP_MODE_FIFO_REFERENCE, &gr); st != C.GP_OK {
		return false, false
	}
	cpu, mem, gpu := make([]int64, len(s.names)), make([]int64, len(s.names)), make([]int64, len(s.names))
	if st := C.gp_get_snapshot(d.ctx, ptr64(cpu), ptr64(mem), ptr64(gpu)); st != C.GP_OK {
		return false, false
	}
	for i, n := range s.names {
		a := metadata[n].AvailableResources
		a.CPU = *resource.NewMilliQuantity(cpu[i], resource.DecimalSI)
		a.Memory = *resource.NewQuantity(mem[i], resource.BinarySI)
		a.NvidiaGPU = *resource.NewQuantity(gpu[i], resource.DecimalSI)
	}
	fits = true
	for i := range apps {
		if driver[i] == -2 || (driver[i] == -1 && skip[i] == 0) {
			fits = false
		}
	}
	return fits, true
}

// FitEarlierDriversSingleAZ is FitEarlierDriversBatch for `binpack: single-az-tightly-pack` (algo GP_TIGHTLY_PACK) and
// `single-az-minimal-fragmentation` (algo GP_MINIMAL_FRAGMENTATION): fitEarlierDrivers (EXT/resource.go:224-262) when
// binpacker.BinpackFunc packs every zone and keeps the best result (LIB/binpack/single_az.go:23-97).  The zones of
// groupNodesByZone (:57-73) become the snapshot's instance groups, SchedulableResources go up with gp_set_schedulable, and
// ONE gp_pack_fifo_zones call runs the whole queue on the device -- the zone chosen for driver i feeds driver i+1 there.
// ok=false: the caller must run the original Go loop instead.
func FitEarlierDriversSingleAZ(algo C.gp_algo, apps []QueuedApp, nodeNames, executorNodeNames []string,
	metadata resources.NodeGroupSchedulingMetadata) (fits bool, ok bool) {
	d, err := getDevice()
	if err != nil {
		return false, false
	}
	if len(apps) == 0 {
		return true, true
	}
	// groupNodesByZone for both orders; zones without executor candidates are left out (single_az.go:38-41)
	group := func(names []string) ([]string, map[string][]string) {
		order, byZone := []string{}, map[string][]string{}
		for _, n := range names {
			m, present := metadata[n]
			if !present {
				continue
			}
			if _, seen := byZone[m.ZoneLabel]; !seen {
				order = append(order, m.ZoneLabel)
			}
			byZone[m.ZoneLabel] = append(byZone[m.ZoneLabel], n)
		}
		return order, byZone
	}
	zones, dz := group(nodeNames)
	_, ez := group(executorNodeNames)
	s := &snapshot{}
	index := map[string]int32{}
	var sc, sm, sg []int64
	intern := func(n string) (int32, bool) {
		if i, seen := index[n]; seen {
			return i, true
		}
		m := metadata[n]
		c, mm, g, exact := toTriple(m.AvailableResources)
		c2, m2, g2, exact2 := toTriple(m.SchedulableResources)
		if !exact || !exact2 {
			return -1, false
		}
		i := int32(len(s.names))
		index[n] = i
		s.names = append(s.names, n)
		s.cpu, s.mem, s.gpu = append(s.cpu, c), append(s.mem, mm), append(s.gpu, g)
		sc, sm, sg = append(sc, c2), append(sm, m2), append(sg, g2)
		return i, true
	}
	eoff, doff := []int32{0}, []int32{0}
	for _, z := range zones {
		eo, has := ez[z]
		if !has {
			continue
		}
		for _, n := range eo {
			i, exact := intern(n)
			if !exact {
				return false, false
			}
			s.execIdx = append(s.execIdx, i)
		}
		for _, n := range dz[z] {
			i, exact := intern(n)
			if !exact {
				return false, false
			}
			s.drvIdx = append(s.drvIdx, i)
		}
		eoff, doff = append(eoff, int32(len(s.execIdx))), append(doff, int32(len(s.drvIdx)))
	}
	nGroups := len(eoff) - 1
	if nGroups == 0 || nGroups > 64 {
		return false, false // no candidate zone (every BinpackFunc call is an EmptyPackingResult) or too many: the Go loop decides
	}
	q := len(apps)
	dc, dm, dg := make([]int64, q), make([]int64, q), make([]int64, q)
	ec, em, eg := make([]int64, q), make([]int64, q), make([]int64, q)
	cnt, skip := make([]int32, q), make([]uint8, q)
	total := 0
	for i, a := range apps {
		var o1, o2 bool
		dc[i], dm[i], dg[i], o1 = toTriple(a.Driver)
		ec[i], em[i], eg[i], o2 = toTriple(a.Executor)
		if !o1 || !o2 || a.MinExecutorCount < 0 {
			return false, false
		}
		cnt[i] = int32(a.MinExecutorCount)
		if a.SkipIfNoFit {
			skip[i] = 1
		}
		total += a.MinExecutorCount
	}
	zone, driver := make([]int32, q), make([]int32, q)
	exec := make([]int32, max(total, 1))
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	d.mu.Lock()
	defer d.mu.Unlock()
	var pin runtime.Pinner
	defer pin.Unpin()
	for _, p := range []any{ptrOrNil(s.cpu), ptrOrNil(s.mem), ptrOrNil(s.gpu), ptrOrNil32(s.execIdx), ptrOrNil32(s.drvIdx),
		ptrOrNil32(eoff), ptrOrNil32(doff), ptrOrNil(sc), ptrOrNil(sm), ptrOrNil(sg),
		ptrOrNil(dc), ptrOrNil(dm), ptrOrNil(dg), ptrOrNil(ec), ptrOrNil(em), ptrOrNil(eg),
		ptrOrNil32(cnt), ptrOrNil8(skip), ptrOrNil32(zone), ptrOrNil32(driver), ptrOrNil32(exec)} {
		if p != nil {
			pin.Pin(p)
		}
	}
	gn := C.gp_nodes{
		n_nodes: C.int32_t(len(s.names)), avail_cpu_milli: ptr64(s.cpu), avail_mem_bytes: ptr64(s.mem), avail_gpu: ptr64(s.gpu),
		n_groups: C.int32_t(nGroups), exec_off: ptr32(eoff), exec_order: ptr32(s.execIdx), drv_off: ptr32(doff), drv_order: ptr32(s.drvIdx),
	}
	if st := C.gp_set_snapshot(d.ctx, &gn); st != C.GP_OK {
		return false, false
	}
	if st := C.gp_set_schedulable(d.ctx, ptr64(sc), ptr64(sm), ptr64(sg)); st != C.GP_OK {
		return false, false
	}
	ga := C.gp_apps{n_apps: C.int32_t(q), drv_cpu_milli: ptr64(dc), drv_mem_bytes: ptr64(dm), drv_gpu: ptr64(dg),
		exe_cpu_milli: ptr64(ec), exe_mem_bytes: ptr64(em), exe_gpu: ptr64(eg), exe_count: ptr32(cnt), skip_if_no_fit: ptr8(skip)}
	gz := C.gp_zone_results{zone: ptr32(zone), driver_node: ptr32(driver), executor_nodes: ptr32(exec), executor_nodes_cap: C.int64_t(len(exec))}
	if st := C.gp_pack_fifo_zones(d.ctx, &ga, algo, C.GP_MODE_FIFO_REFERENCE, &gz); st != C.GP_OK {
		return false, false
	}
	cpu, mem, gpu := make([]int64, len(s.names)), make([]int64, len(s.names)), make([]int64, len(s.names))
	if st := C.gp_get_snapshot(d.ctx, ptr64(cpu), ptr64(mem), ptr64(gpu)); st != C.GP_OK {
		return false, false
	}
	for i, n := range s.names { // SubtractUsageIfExists (LIB/resources/resources.go:129-135), as the device left it
		a := metadata[n].AvailableResources
		a.CPU = *resource.NewMilliQuantity(cpu[i], resource.DecimalSI)
		a.Memory = *resource.NewQuantity(mem[i], resource.BinarySI)
		a.NvidiaGPU = *resource.NewQuantity(gpu[i], resource.DecimalSI)
	}
	fits = true
	for i := range apps {
		if driver[i] == -2 || (driver[i] == -1 && skip[i] == 0) { // EXT/resource.go:250-252
			fits = false
		}
	}
	return fits, true
}

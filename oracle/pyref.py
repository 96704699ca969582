"""Pure-Python literal restatement of the reference hot path (dicts keyed by node name).

TEST INFRASTRUCTURE ONLY; small cases only (pure-Python loops).  It exists as an independent third
statement of the algorithm: tests/gen_golden.py uses it to derive tests/golden/*.json and the tests
check the two C oracles against it.  Parity status: partially pinned (see gangpack_oracle.h).

LIB = /root/reference/vendor/github.com/palantir/k8s-spark-scheduler-lib/pkg, EXT = internal/extender.
Resources are (cpu_milli, mem_bytes, gpu) tuples.
"""
from __future__ import annotations

import functools


def gt(a, b):
    """Resources.GreaterThan, LIB/resources/resources.go:239-241."""
    return a[0] > b[0] or a[1] > b[1] or a[2] > b[2]


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


ZERO = (0, 0, 0)


def tightly_pack_executors(exe, count, order, meta, reserved):
    """LIB/binpack/pack_tightly.go:34-63.  meta: name -> available Resources."""
    out = []
    if count == 0:
        return out, True
    for n in order:
        if n not in reserved:
            reserved[n] = ZERO
        while True:
            reserved[n] = add(reserved[n], exe)
            if n not in meta or gt(reserved[n], meta[n]):
                reserved[n] = sub(reserved[n], exe)
                break
            out.append(n)
            if len(out) == count:
                return out, True
    return None, False


def distribute_executors_evenly(exe, count, order, meta, reserved):
    """LIB/binpack/distribute_evenly.go:34-73."""
    available = {n: True for n in order}
    out = []
    if count == 0:
        return out, True
    while len(available) > 0:
        for n in order:
            if n not in available:
                continue
            if n not in reserved:
                reserved[n] = ZERO
            reserved[n] = add(reserved[n], exe)
            if n not in meta or gt(reserved[n], meta[n]):
                del available[n]
                reserved[n] = sub(reserved[n], exe)
            else:
                out.append(n)
                if len(out) == count:
                    return out, True
    return None, False


MAX_INT = (1 << 63) - 1


def _go_int(v):
    """wrap to Go's 64-bit int"""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _capacity_single_dimension(available, reserved, required):
    """getCapacityAgainstSingleDimension, LIB/capacity/capacity.go:36-55."""
    if reserved > available:
        return 0
    if required == 0:
        return MAX_INT
    return (available - reserved) // required


def node_capacities(order, meta, reserved, exe):
    """GetNodeCapacities, LIB/capacity/capacity.go:78-102 -> [(name, capacity)] in order."""
    out = []
    for n in order:
        if n in meta:
            r = reserved.get(n, ZERO)
            out.append((n, min(_capacity_single_dimension(meta[n][d], r[d], exe[d]) for d in range(3))))
    return out


def _search(caps, pred):
    """sort.Search: smallest index with pred true (pred monotone), len(caps) if none."""
    lo, hi = 0, len(caps)
    while lo < hi:
        mid = (lo + hi) // 2
        if pred(caps[mid]):
            hi = mid
        else:
            lo = mid + 1
    return lo


def _internal_minimal_fragmentation(count, caps):
    """internalMinimalFragmentation, LIB/binpack/minimal_fragmentation.go:96-137."""
    caps = list(caps)
    out = []
    while caps:
        position = _search(caps, lambda c: c[1] >= count)
        if position != len(caps):
            return out + [caps[position][0]] * count, True
        max_capacity = caps[-1][1]
        first_max = _search(caps, lambda c: c[1] >= max_capacity)
        pos = first_max
        while count >= max_capacity and pos < len(caps):
            out += [caps[pos][0]] * max_capacity
            count -= max_capacity
            pos += 1
        if count == 0:
            return out, True
        caps = caps[:first_max] + caps[pos:]
    return None, False


def minimal_fragmentation(exe, count, order, meta, reserved):
    """minimalFragmentation, LIB/binpack/minimal_fragmentation.go:59-94 (does not touch `reserved`)."""
    if count == 0:
        return [], True
    caps = [c for c in node_capacities(order, meta, reserved, exe) if c[1] > 0]
    if not caps:
        return None, False
    caps.sort(key=lambda c: c[1])  # list.sort is stable, like sort.SliceStable
    max_capacity = caps[-1][1]
    if count < max_capacity:
        wrapped = _go_int(count + max_capacity)
        target = -((-wrapped) // 2) if wrapped < 0 else wrapped // 2  # Go integer division truncates toward zero
        first = _search(caps, lambda c: c[1] >= target)
        nodes, ok = _internal_minimal_fragmentation(count, caps[:first])
        if ok:
            return nodes, True
    return _internal_minimal_fragmentation(count, caps)


DISTRIBUTORS = {"tightly-pack": tightly_pack_executors, "distribute-evenly": distribute_executors_evenly,
                "minimal-fragmentation": minimal_fragmentation}


def spark_bin_pack(drv, exe, count, driver_order, exec_order, meta, distribute):
    """LIB/binpack/binpack.go:60-87 -> (driver|None, executor_nodes, has_capacity)."""
    for d in driver_order:
        if d not in meta or gt(drv, meta[d]):
            continue
        reserved = {d: drv}
        nodes, ok = distribute(exe, count, exec_order, meta, reserved)
        if ok:
            return d, nodes, True
    return None, [], False


def spark_resource_usage(drv, exe, driver_node, executor_nodes):
    """EXT/sparkpods.go:139-146 (assignment, not accumulation)."""
    res = {driver_node: drv}
    for n in executor_nodes:
        res[n] = exe
    return res


def fit_earlier_drivers(apps, driver_order, exec_order, meta, algo, mode="reference"):
    """EXT/resource.go:224-262.  apps: dicts {drv, exe, count, young}.  Mutates meta.
    Returns (blocked_index|-1, [(driver|None|'unevaluated', nodes)])."""
    distribute = DISTRIBUTORS[algo]
    results = []
    blocked = -1
    for i, a in enumerate(apps):
        if blocked >= 0:
            results.append(("unevaluated", []))
            continue
        d, nodes, ok = spark_bin_pack(a["drv"], a["exe"], a["count"], driver_order, exec_order, meta, distribute)
        if not ok:
            results.append((None, []))
            if a.get("young"):
                continue
            blocked = i
            continue
        results.append((d, nodes))
        if mode == "reference":
            usage = spark_resource_usage(a["drv"], a["exe"], d, nodes)
            for n, u in usage.items():  # SubtractUsageIfExists, LIB/resources/resources.go:129-135
                if n in meta:
                    meta[n] = sub(meta[n], u)
        else:
            meta[d] = sub(meta[d], a["drv"])
            for n in nodes:
                meta[n] = sub(meta[n], a["exe"])
    return blocked, results


def resources_less_than(l, r):
    """internal/sort/nodesorting.go:74-80."""
    if l[1] != r[1]:
        return l[1] < r[1]
    return l[0] < r[0]


def node_names_in_priority_order(meta, zones):
    """getNodeNamesInPriorityOrder, internal/sort/nodesorting.go:95-122.  meta: name -> avail,
    zones: name -> zone label.  Python's sort is stable; the reference's is not."""
    by_az = {}
    for n in meta:
        by_az.setdefault(zones.get(n, "default"), []).append(n)
    az_res = {}
    for az, ns in by_az.items():
        tot = ZERO
        for n in ns:
            tot = add(tot, meta[n])
        az_res[az] = tot

    def az_cmp(a, b):
        if resources_less_than(az_res[a], az_res[b]):
            return -1
        if resources_less_than(az_res[b], az_res[a]):
            return 1
        return 0

    azs = sorted(by_az.keys(), key=functools.cmp_to_key(az_cmp))
    prio = {az: i for i, az in enumerate(azs)}

    def less(a, b):  # scheduleContextLessThan :83-93
        pa, pb = prio[zones.get(a, "default")], prio[zones.get(b, "default")]
        if pa != pb:
            return pa < pb
        # the documented deterministic order (include/gangpack.h, gp_potential_nodes): when resourcesLessThan is
        # false BOTH ways -- equal (memory, cpu), whatever the gpu column says (:88-92 never looks at it) -- the
        # name decides.  The reference leaves that case to an unstable sort.
        if resources_less_than(meta[a], meta[b]):
            return True
        if resources_less_than(meta[b], meta[a]):
            return False
        return a < b

    def cmp(a, b):
        if less(a, b):
            return -1
        if less(b, a):
            return 1
        return 0

    return sorted(meta.keys(), key=functools.cmp_to_key(cmp))


def potential_nodes(meta, zones, candidate_names, unschedulable=(), not_ready=(),
                    driver_label_rank=None, exec_label_rank=None):
    """PotentialNodes, internal/sort/nodesorting.go:41-64.  *_label_rank: name -> rank (absent = unknown)."""
    order = node_names_in_priority_order(meta, zones)
    cand = set(candidate_names)
    drivers = [n for n in order if n in cand]
    execs = [n for n in order if n not in unschedulable and n not in not_ready]

    def label_sort(names, rank):
        if rank is None:
            return names

        def less(a, b):  # createLabelLessThanFunction :161-180
            if a not in rank:
                return False
            if b not in rank:
                return True
            return rank[a] < rank[b]

        def cmp(a, b):
            if less(a, b):
                return -1
            if less(b, a):
                return 1
            return 0

        return sorted(names, key=functools.cmp_to_key(cmp))

    return label_sort(drivers, driver_label_rank), label_sort(execs, exec_label_rank)


# ---------------------------------------------------------------------------------------------
# single-AZ / AZ-aware tightly-pack with the float64 packing efficiencies
# ---------------------------------------------------------------------------------------------
def cpu_value(milli):
    """Quantity.Value() of a milli-scaled CPU: whole cores, inexact values rounded away from zero
    (quantity.go:732-734 -> math.go:166-199)."""
    q, rem = abs(milli) // 1000, abs(milli) % 1000
    v = q + (1 if rem else 0)
    return v if milli >= 0 else -v


def compute_packing_efficiency(avail, sched, reserved):
    """computePackingEfficiency, LIB/binpack/efficiency.go:79-102 -> (cpu, mem, gpu)."""
    r = sub(sched, avail)
    if reserved is not None:
        r = add(r, reserved)
    norm = lambda v: 1 if v == 0 else v
    gpu = 0.0
    if sched[2] != 0:
        gpu = float(r[2]) / float(norm(sched[2]))
    return (float(cpu_value(r[0])) / float(norm(cpu_value(sched[0]))), float(r[1]) / float(norm(sched[1])), gpu)


def avg_packing_efficiency(effs_with_sched_gpu):
    """ComputeAvgPackingEfficiency, efficiency.go:114-156 over [(cpu, mem, gpu, sched_gpu)] -> (cpu, mem, gpu, max)."""
    if not effs_with_sched_gpu:
        return (0.0, 0.0, 0.0, 0.0)
    cpu = mem = gpu = mx = 0.0
    with_gpu = 0
    for c, m, g, sg in effs_with_sched_gpu:
        cpu += c
        mem += m
        if sg != 0:
            gpu += g
            with_gpu += 1
        mx += max(g, max(c, m))
    n = max(float(len(effs_with_sched_gpu)), 1.0)
    return (cpu / n, mem / n, 1.0 if with_gpu == 0 else gpu / float(with_gpu), mx / n)


def _reserved_of(drv, exe, driver, nodes):
    reserved = {driver: drv}
    for n in nodes:
        reserved[n] = add(reserved.get(n, ZERO), exe)
    return reserved


def group_nodes_by_zone(order, meta, zones):
    """groupNodesByZone, LIB/binpack/single_az.go:57-73."""
    in_order, by_zone = [], {}
    for n in order:
        if n not in meta:
            continue
        z = zones.get(n, "default")
        if z not in by_zone:
            in_order.append(z)
            by_zone[z] = []
        by_zone[z].append(n)
    return in_order, by_zone


def single_az_tightly_pack(drv, exe, count, driver_order, exec_order, meta, sched, zones, distribute=None):
    """getSingleAZSparkBinFunction(fn) + chooseBestResult, single_az.go:23-55,75-97 (fn: tightlyPackExecutors unless given)."""
    distribute = distribute or tightly_pack_executors
    dz_order, dz = group_nodes_by_zone(driver_order, meta, zones)
    _, ez = group_nodes_by_zone(exec_order, meta, zones)
    best, best_max = (None, [], False), 0.0
    for z in dz_order:
        if z not in ez:
            continue
        d, nodes, ok = spark_bin_pack(drv, exe, count, dz[z], ez[z], meta, distribute)
        if not ok:
            continue
        # the efficiencies come from SparkBinPack's `reserved` map (binpack.go:72-77): tightly-pack adds every
        # executor to it, minimalFragmentation never does -> driver only
        reserved = {d: drv} if distribute is minimal_fragmentation else _reserved_of(drv, exe, d, nodes)
        effs = [compute_packing_efficiency(meta[n], sched[n], reserved.get(n)) + (sched[n][2],) for n in [d] + nodes]
        avg = avg_packing_efficiency(effs)
        if best_max < avg[3]:
            best, best_max = (d, nodes, True), avg[3]
    return best


def single_az_minimal_fragmentation(drv, exe, count, driver_order, exec_order, meta, sched, zones):
    """SingleAZMinimalFragmentation, LIB/binpack/single_az_minimal_fragmentation.go:20."""
    return single_az_tightly_pack(drv, exe, count, driver_order, exec_order, meta, sched, zones, minimal_fragmentation)


def az_aware_tightly_pack(drv, exe, count, driver_order, exec_order, meta, sched, zones):
    """AzAwareTightlyPack, LIB/binpack/az_aware_pack_tightly.go:27-38."""
    r = single_az_tightly_pack(drv, exe, count, driver_order, exec_order, meta, sched, zones)
    if r[2]:
        return r
    return spark_bin_pack(drv, exe, count, driver_order, exec_order, meta, tightly_pack_executors)


# ---------------------------------------------------------------------------------------------
# executor reschedule
# ---------------------------------------------------------------------------------------------
def reschedule_first_fit(exe, exec_order, available):
    """rescheduleExecutor, EXT/resource.go:657-662: first node of the order the executor fits on, else None."""
    for n in exec_order:
        if n in available and not gt(exe, available[n]):
            return n
    return None


def reschedule_minimal_fragmentation(exe, exec_order, meta, overhead, hosting):
    """rescheduleExecutorWithMinimalFragmentation, EXT/resource.go:675-705 (overhead is what it passes to
    GetNodeCapacities as `reserved`; hosting = getNodesWithExecutorsBelongingToSameApp)."""
    best = None
    for n, cap in node_capacities(exec_order, meta, overhead, exe):
        if cap >= 1:
            if best is None:
                best = (n, cap)
            elif (n in hosting) and (best[0] not in hosting):
                best = (n, cap)
            elif ((n in hosting) == (best[0] in hosting)) and cap < best[1]:
                best = (n, cap)
    return best[0] if best else None


# ---------------------------------------------------------------------------------------------
# snapshot build
# ---------------------------------------------------------------------------------------------
def node_scheduling_metadata(alloc, overhead, reservations):
    """GetReservedResources (EXT/resourcereservations.go:258-263) + NodeSchedulingMetadataForNodes
    (LIB/resources/resources.go:61-100).  alloc / overhead: name -> Resources; reservations: [(node, Resources)]
    (hard and soft alike).  -> (available, schedulable) dicts over the nodes of `alloc`."""
    usage = {}
    for node, r in reservations:                      # UsageForNodes :31-43, softreservations.go:155-170
        usage[node] = add(usage.get(node, ZERO), r)
    available, schedulable = {}, {}
    for n, a in alloc.items():
        o = overhead.get(n, ZERO)
        u = add(usage.get(n, ZERO), o)                # :76
        available[n] = sub(a, u)                      # :89
        schedulable[n] = sub(a, o)                    # :90
    return available, schedulable


# ---------------------------------------------------------------------------------------------
# inputs of the path: annotations -> application tuple, the FIFO queue
# ---------------------------------------------------------------------------------------------
_DEC_SUFFIX = {"n": -9, "u": -6, "m": -3, "": 0, "k": 3, "M": 6, "G": 9, "T": 12, "P": 15, "E": 18}   # suffix.go:121-132
_BIN_SUFFIX = {"Ki": 10, "Mi": 20, "Gi": 30, "Ti": 40, "Pi": 50, "Ei": 60}                            # suffix.go:113-118


def parse_quantity(s):
    """resource.ParseQuantity (vendor/k8s.io/apimachinery/pkg/api/resource/quantity.go:147-300) -> (status, Fraction).
    status: 'ok' | 'ErrFormatWrong' | 'ErrSuffix'.  The value is exact (no rounding to nano, no int64 cap: the callers
    here decide what their integer model can hold)."""
    from fractions import Fraction
    import re
    if s == "":
        return "ErrFormatWrong", None
    m = re.fullmatch(r"([+-]?)([0-9]*)(?:\.([0-9]*))?([eEinumkKMGTP]*[+-]?[0-9]*)", s)
    if not m:
        return "ErrFormatWrong", None
    sign, num, denom, suffix = m.group(1), m.group(2), m.group(3) or "", m.group(4)
    if suffix in _DEC_SUFFIX:
        mult = Fraction(10) ** _DEC_SUFFIX[suffix]
    elif suffix in _BIN_SUFFIX:
        mult = Fraction(2) ** _BIN_SUFFIX[suffix]
    elif len(suffix) > 1 and suffix[0] in "eE" and re.fullmatch(r"[+-]?[0-9]+", suffix[1:]):
        mult = Fraction(10) ** int(suffix[1:])
    else:
        return "ErrSuffix", None
    digits = (num or "0") + denom
    v = Fraction(int(digits), 10 ** len(denom)) * mult
    return "ok", (-v if sign == "-" else v)


def quantity_scaled(s, scale, round_up_fraction=False):
    """-> (status, int): the quantity in units of 10^-scale if that is an exact integer below 2^61 ('unrepresentable'
    otherwise); round_up_fraction reproduces Quantity.Value() (away from zero, quantity.go:732-734)."""
    st, v = parse_quantity(s)
    if st != "ok":
        return st, None
    v = v * 10 ** scale
    if v.denominator != 1:
        if not round_up_fraction:
            return "unrepresentable", None
        v = abs(v)
        v = (v.numerator // v.denominator + 1) * (1 if parse_quantity(s)[1] > 0 else -1)
    v = int(v)
    if abs(v) >= 1 << 61:
        return "unrepresentable", None
    return "ok", v


def spark_resources(annotations):
    """sparkResources, EXT/sparkpods.go:73-137 -> (error_text | None, dict).  Quantities in millicores / bytes / units;
    'exact': False when one of them is outside the int64 model."""
    da = False
    if "spark-dynamic-allocation-enabled" in annotations:
        v = annotations["spark-dynamic-allocation-enabled"]
        if v in ("1", "t", "T", "TRUE", "true", "True"):
            da = True
        elif v in ("0", "f", "F", "FALSE", "false", "False"):
            da = False
        else:
            return "annotation DynamicAllocationEnabled could not be parsed as a boolean", None
    fields = [("spark-driver-cpu", 3), ("spark-driver-mem", 0), ("spark-driver-nvidia.com/gpu", 0), ("spark-executor-cpu", 3),
              ("spark-executor-mem", 0), ("spark-executor-nvidia.com/gpu", 0), ("spark-executor-count", 0),
              ("spark-dynamic-allocation-min-executor-count", 0), ("spark-dynamic-allocation-max-executor-count", 0)]
    counts = {"spark-executor-count", "spark-dynamic-allocation-min-executor-count", "spark-dynamic-allocation-max-executor-count"}
    parsed, exact = {}, True
    for a, scale in fields:
        if a not in annotations:
            if a in ("spark-driver-nvidia.com/gpu", "spark-executor-nvidia.com/gpu"):
                continue
            if not da and a == "spark-executor-count":
                return "annotation ExecutorCount is required when DynamicAllocationEnabled is false", None
            if da and a in ("spark-dynamic-allocation-min-executor-count", "spark-dynamic-allocation-max-executor-count"):
                return "annotation %s is required when DynamicAllocationEnabled is true" % a, None
            if a in counts:
                continue
            return "annotation %s is missing from driver" % a, None
        st, v = quantity_scaled(annotations[a], scale, round_up_fraction=a in counts)
        if st == "unrepresentable":
            exact = False
            continue
        if st != "ok":
            return "annotation %s does not have a parseable value %s" % (a, annotations[a]), None
        parsed[a] = v
    g = lambda k: parsed.get(k, 0)
    lo, hi = ((g("spark-dynamic-allocation-min-executor-count"), g("spark-dynamic-allocation-max-executor-count")) if da
              else (g("spark-executor-count"), g("spark-executor-count")))
    return None, {"drv": [g("spark-driver-cpu"), g("spark-driver-mem"), g("spark-driver-nvidia.com/gpu")],
                  "exe": [g("spark-executor-cpu"), g("spark-executor-mem"), g("spark-executor-nvidia.com/gpu")],
                  "min": lo, "max": hi, "exact": exact}


def filter_to_earliest_and_sort(driver, all_drivers):
    """filterToEarliestAndSort, EXT/sparkpods.go:54-74.  Pods are dicts {uid, created, node, scheduler, group, deleting}
    (group None = no instance group found, internal/podspec.go:22-35).  Equal timestamps keep their input order."""
    earlier = [p for p in all_drivers
               if not p.get("node") and p.get("scheduler", "") == driver.get("scheduler", "")
               and p.get("group") is not None and p.get("group") == driver.get("group")
               and p["created"] < driver["created"] and not p.get("deleting")]
    return sorted(earlier, key=lambda p: p["created"])

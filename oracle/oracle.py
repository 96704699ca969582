"""ctypes binding of the CPU oracle (oracle/gangpack_oracle.{h,c}, gangpack_closed.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity status: partially pinned (see gangpack_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgangpack_oracle.so")

TIGHTLY_PACK = 0
DISTRIBUTE_EVENLY = 1
MODE_INDEPENDENT = 0
MODE_FIFO_REFERENCE = 1
MODE_FIFO_EXACT = 2


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("gangpack_oracle.c", "gangpack_closed.c", "gangpack_oracle.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgangpack_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_cluster_new.restype = C.c_void_p
        L.orc_cluster_new.argtypes = [C.c_int32, C.POINTER(C.c_char_p)] + [C.c_void_p] * 6 + [
            C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p]
        L.orc_cluster_free.argtypes = [C.c_void_p]
        L.orc_cluster_index.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_cluster_index.restype = C.c_int32
        L.orc_cluster_get_available.argtypes = [C.c_void_p] * 4
        L.orc_binpack.restype = C.c_int
        L.orc_binpack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_char_p), C.c_int32,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_binpack_batch.restype = None
        L.orc_binpack_batch.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_char_p), C.c_int32,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fifo.restype = C.c_int32
        L.orc_fifo.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_char_p), C.c_int32,
                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_potential_nodes.restype = None
        L.orc_potential_nodes.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]
        L.orc_node_scheduling_metadata.restype = None
        L.orc_node_scheduling_metadata.argtypes = [C.c_int32, C.POINTER(C.c_char_p)] + [C.c_void_p] * 6 + [
            C.c_int64, C.POINTER(C.c_char_p)] + [C.c_void_p] * 9
        L.orc_reschedule_available.restype = None
        L.orc_reschedule_available.argtypes = [C.c_int32, C.POINTER(C.c_char_p)] + [C.c_void_p] * 6 + [
            C.c_int64, C.POINTER(C.c_char_p)] + [C.c_void_p] * 6
        L.orc_reschedule_executor.restype = C.c_int32
        L.orc_reschedule_executor.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_char_p), C.c_int32,
                                              C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.c_int32]
        L.orc_closed_batch.restype = C.c_int32
        L.orc_closed_batch.argtypes = [C.c_int, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _names(names):
    arr = (C.c_char_p * max(len(names), 1))()
    for i, n in enumerate(names):
        arr[i] = n.encode() if isinstance(n, str) else n
    return arr


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def res_array(cpu, mem, gpu):
    """[n] x orc_res as an (n,3) int64 array."""
    return np.ascontiguousarray(np.stack([_i64(cpu), _i64(mem), _i64(gpu)], axis=1))


def exec_offsets(count):
    off = np.zeros(len(count) + 1, dtype=np.int64)
    np.cumsum(np.asarray(count, dtype=np.int64), out=off[1:])
    return off


class Cluster:
    """NodeGroupSchedulingMetadata of the literal oracle (string-keyed)."""

    def __init__(self, names, avail_cpu, avail_mem, avail_gpu=None, sched=None, zone=None,
                 unschedulable=None, ready=None):
        self.names = [n if isinstance(n, str) else n.decode() for n in names]
        n = len(self.names)
        self._keep = []
        ac, am = _i64(avail_cpu), _i64(avail_mem)
        ag = _i64(avail_gpu) if avail_gpu is not None else np.zeros(n, dtype=np.int64)
        sc = sm = sg = None
        if sched is not None:
            sc, sm, sg = (_i64(x) for x in sched)
        z = _names(zone) if zone is not None else None
        u = np.ascontiguousarray(unschedulable, dtype=np.uint8) if unschedulable is not None else None
        r = np.ascontiguousarray(ready, dtype=np.uint8) if ready is not None else None
        self._h = lib().orc_cluster_new(n, _names(self.names), _ptr(ac), _ptr(am), _ptr(ag),
                                        _ptr(sc), _ptr(sm), _ptr(sg), z, _ptr(u), _ptr(r))
        self.n = n

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_cluster_free(self._h)
            self._h = None

    def available(self):
        cpu = np.empty(self.n, np.int64); mem = np.empty(self.n, np.int64); gpu = np.empty(self.n, np.int64)
        lib().orc_cluster_get_available(self._h, _ptr(cpu), _ptr(mem), _ptr(gpu))
        return cpu, mem, gpu

    def binpack(self, algo, drv, exe, count, driver_order, exec_order, with_efficiencies=False):
        """SparkBinPackFunction.  drv/exe = (cpu, mem, gpu); orders = node names.
        Returns (has_capacity, driver_name|None, [executor names], avg_eff[4])."""
        d = np.array(drv, dtype=np.int64); e = np.array(exe, dtype=np.int64)
        dn = C.c_int32(-1)
        en = np.full(max(count, 1), -1, dtype=np.int32)
        eff = np.zeros(4, dtype=np.float64)
        ok = lib().orc_binpack(self._h, algo, _ptr(d), _ptr(e), count,
                               _names(driver_order), len(driver_order), _names(exec_order), len(exec_order),
                               int(with_efficiencies), C.byref(dn), _ptr(en), _ptr(eff))
        if not ok:
            return False, None, [], eff
        return True, self.names[dn.value], [self.names[i] for i in en[:count]], eff

    def binpack_batch(self, algo, drv, exe, count, driver_order, exec_order, with_efficiencies=False, n_threads=1):
        count = np.ascontiguousarray(count, dtype=np.int32)
        off = exec_offsets(count)
        q = len(count)
        driver_node = np.full(q, -9, dtype=np.int32)
        executor_nodes = np.full(max(int(off[-1]), 1), -1, dtype=np.int32)
        lib().orc_binpack_batch(self._h, algo, q, _ptr(drv), _ptr(exe), _ptr(count),
                                _names(driver_order), len(driver_order), _names(exec_order), len(exec_order),
                                int(with_efficiencies), n_threads, _ptr(off), _ptr(driver_node), _ptr(executor_nodes))
        return driver_node, executor_nodes, off

    def fifo(self, algo, mode, drv, exe, count, young, driver_order, exec_order, with_efficiencies=False):
        count = np.ascontiguousarray(count, dtype=np.int32)
        off = exec_offsets(count)
        q = len(count)
        y = np.ascontiguousarray(young, dtype=np.uint8) if young is not None else None
        driver_node = np.full(q, -9, dtype=np.int32)
        executor_nodes = np.full(max(int(off[-1]), 1), -1, dtype=np.int32)
        blocked = lib().orc_fifo(self._h, algo, mode, q, _ptr(drv), _ptr(exe), _ptr(count), _ptr(y),
                                 _names(driver_order), len(driver_order), _names(exec_order), len(exec_order),
                                 int(with_efficiencies), _ptr(off), _ptr(driver_node), _ptr(executor_nodes))
        return blocked, driver_node, executor_nodes, off

    def reschedule_executor(self, min_frag, exe, exec_order, reserved=None, hosting=()):
        """rescheduleExecutor's node choice -> node name or None.  reserved: {name: (cpu, mem, gpu)}."""
        e = np.array(exe, dtype=np.int64)
        reserved = reserved or {}
        rn = list(reserved)
        rv = np.ascontiguousarray([reserved[k] for k in rn] or [(0, 0, 0)], dtype=np.int64)
        hosting = list(hosting)
        i = lib().orc_reschedule_executor(self._h, int(bool(min_frag)), _ptr(e), _names(exec_order), len(exec_order),
                                          _names(rn), _ptr(rv), len(rn), _names(hosting), len(hosting))
        return self.names[i] if i >= 0 else None

    def potential_nodes(self, candidate_names, driver_label_rank=None, exec_label_rank=None):
        d = np.empty(max(self.n, 1), np.int32); e = np.empty(max(self.n, 1), np.int32)
        nd = C.c_int32(0); ne = C.c_int32(0)
        dr = np.ascontiguousarray(driver_label_rank, dtype=np.int32) if driver_label_rank is not None else None
        er = np.ascontiguousarray(exec_label_rank, dtype=np.int32) if exec_label_rank is not None else None
        lib().orc_potential_nodes(self._h, _names(candidate_names), len(candidate_names), _ptr(dr), _ptr(er),
                                  _ptr(d), C.byref(nd), _ptr(e), C.byref(ne))
        return ([self.names[i] for i in d[:nd.value]], [self.names[i] for i in e[:ne.value]])


def node_scheduling_metadata(names, alloc, overhead, res_node_names, res):
    """(alloc, overhead, res) = triples of int64 arrays (cpu, mem, gpu); -> (avail triple, sched triple)."""
    n = len(names)
    al = [_i64(x) for x in alloc]
    ov = [_i64(x) for x in overhead] if overhead is not None else [None] * 3
    rs = [_i64(x) for x in res]
    av = [np.empty(n, np.int64) for _ in range(3)]
    sc = [np.empty(n, np.int64) for _ in range(3)]
    lib().orc_node_scheduling_metadata(n, _names(names), _ptr(al[0]), _ptr(al[1]), _ptr(al[2]), _ptr(ov[0]), _ptr(ov[1]), _ptr(ov[2]),
                                       len(res_node_names), _names(res_node_names), _ptr(rs[0]), _ptr(rs[1]), _ptr(rs[2]),
                                       _ptr(av[0]), _ptr(av[1]), _ptr(av[2]), _ptr(sc[0]), _ptr(sc[1]), _ptr(sc[2]))
    return av, sc


def reschedule_available(names, alloc, overhead, res_node_names, res):
    """availableResources of rescheduleExecutor's first-fit branch (overhead counted twice on nodes with reservations)."""
    n = len(names)
    al = [_i64(x) for x in alloc]
    ov = [_i64(x) for x in overhead] if overhead is not None else [None] * 3
    rs = [_i64(x) for x in res]
    av = [np.empty(n, np.int64) for _ in range(3)]
    lib().orc_reschedule_available(n, _names(names), _ptr(al[0]), _ptr(al[1]), _ptr(al[2]), _ptr(ov[0]), _ptr(ov[1]), _ptr(ov[2]),
                                   len(res_node_names), _names(res_node_names), _ptr(rs[0]), _ptr(rs[1]), _ptr(rs[2]),
                                   _ptr(av[0]), _ptr(av[1]), _ptr(av[2]))
    return av


def closed_batch(algo, mode, avail_cpu, avail_mem, avail_gpu, driver_order, exec_order,
                 drv, exe, count, young=None, n_threads=1):
    """Closed-form oracle on index arrays.  Returns (blocked, driver_node, executor_nodes, off,
    (cpu, mem, gpu) after the call)."""
    cpu, mem, gpu = _i64(avail_cpu).copy(), _i64(avail_mem).copy(), _i64(avail_gpu).copy()
    dord = np.ascontiguousarray(driver_order, dtype=np.int32)
    eord = np.ascontiguousarray(exec_order, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.int32)
    off = exec_offsets(count)
    q = len(count)
    y = np.ascontiguousarray(young, dtype=np.uint8) if young is not None else None
    driver_node = np.full(q, -9, dtype=np.int32)
    executor_nodes = np.full(max(int(off[-1]), 1), -1, dtype=np.int32)
    blocked = lib().orc_closed_batch(algo, mode, len(cpu), _ptr(cpu), _ptr(mem), _ptr(gpu),
                                     _ptr(dord), len(dord), _ptr(eord), len(eord),
                                     q, _ptr(drv), _ptr(exe), _ptr(count), _ptr(y), n_threads,
                                     _ptr(off), _ptr(driver_node), _ptr(executor_nodes))
    return blocked, driver_node, executor_nodes, off, (cpu, mem, gpu)

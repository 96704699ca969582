#!/usr/bin/env python
"""bench.py -- gang placements/sec of the B200 bin-packer on the BASELINE workload.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                   (CPU restatement of the reference path)

A "step" is one pass of the hot path over one batch of synthetic input: the 10 000-node snapshot is
laid out on the device (gp_set_snapshot_device) and 100 000 pending applications are packed
(tightly-pack, independent decisions against that snapshot) by prep + pack kernels.  With N>1 every
rank packs its own 100 000 applications (weak scaling).

value : decisions/s with inputs resident in HBM (every rank: snapshot + its block of the queue), no collective
        in the data path, CUDA events on the launch stream, L2 flushed between steps (outside the per-step
        event pairs), max over ranks.
e2e   : host buffers in, host results out.  N=1: gp_set_snapshot + gp_pack_batch from pinned memory through
        the C ABI.  N>1: rank 0 uploads the snapshot, ONE NCCL broadcast distributes it, every rank packs its
        host-resident block, the placements are all-gathered over NVLink (overlapped chunk by chunk with the
        packing) and rank 0 copies all of them to host memory.  Wall clock, max over ranks.
roofline / cpu_baseline: see DESIGN.md section 6.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nodes, apps per GPU, algo, mode, groups)
    "tightly-100k": dict(nodes=10000, apps=100000, algo=0, mode=0, groups=1,
                         desc="10k nodes x 100k pending apps, tightly-pack, independent decisions vs one snapshot "
                              "(BASELINE.json metric / north_star size; configs[1] shape)"),
    "tightly-10k": dict(nodes=10000, apps=10000, algo=0, mode=0, groups=1,
                        desc="10k nodes x 10k pending apps, tightly-pack, independent (BASELINE configs[1])"),
    "evenly-100k": dict(nodes=10000, apps=100000, algo=1, mode=0, groups=1,
                        desc="10k nodes x 100k pending apps, distribute-evenly, independent (BASELINE configs[2])"),
    "minfrag-100k": dict(nodes=10000, apps=100000, algo=2, mode=0, groups=1,
                         desc="10k nodes x 100k pending apps, minimal-fragmentation (the per-zone packer of "
                              "single-az-minimal-fragmentation), independent; every decision is several full passes over the nodes"),
    "fifo-10k": dict(nodes=10000, apps=10000, algo=0, mode=1, groups=1,
                     desc="10k nodes x 10k pending apps, tightly-pack, FIFO (reference usage accounting), 1 instance group"),
    "fifo-da-50k": dict(nodes=10000, apps=50000, algo=0, mode=1, groups=16, da=True,
                        desc="dynamic-allocation sweep, 10k nodes x 50k apps, FIFO on, 16 instance groups (BASELINE configs[3])"),
    "tightly-50k-1m": dict(nodes=50000, apps=125000, algo=0, mode=0, groups=1,
                           desc="50k nodes x 1M pending apps over 8 GPUs (125k per GPU), tightly-pack (BASELINE configs[4])"),
}
ALGO_NAME = {0: "tightly-pack", 1: "distribute-evenly", 2: "minimal-fragmentation"}
ORC_ALGO = {0: 0, 1: 1, 2: 4}      # gp_algo -> oracle algo id (oracle/gangpack_oracle.h)
MODE_NAME = {0: "independent", 1: "fifo-reference", 2: "fifo-exact"}
APP_KEYS = ("drv_cpu", "drv_mem", "drv_gpu", "exe_cpu", "exe_mem", "exe_gpu", "count", "group", "young")


# ------------------------------------------------------------------------------------------------
def make_workload(w: dict, rank: int):
    from k8s_spark_scheduler_b200 import synth
    nodes = synth.make_nodes(w["nodes"], groups=w["groups"])
    # every rank gets a different slice of the (conceptually N x apps long) queue
    apps = synth.make_apps(w["apps"], seed=synth.APP_SEED + 7919 * rank, groups=w["groups"],
                           da_sweep=bool(w.get("da")))
    eoff, eorder = synth.group_orders(nodes)
    a = {k: apps[k] for k in APP_KEYS}
    a["off"] = synth.exec_offsets(apps["count"])
    return nodes, a, eoff, eorder


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active") and not v.lower().startswith("not"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(w: dict, sample_apps: int, threads: int, repeats: int = 1):
    """Times the literal CPU restatement of the reference path (oracle/, kind 'port') on a bounded
    sample of the same workload, with ComputePackingEfficiencies on (the reference runs it inside
    SparkBinPack on every successful pack, binpack.go:77)."""
    from k8s_spark_scheduler_b200 import synth
    from oracle import oracle as orc
    nodes, a, eoff, eorder = make_workload(w, 0)
    names = synth.node_names(w["nodes"])
    q = min(sample_apps, len(a["count"]))
    drv = orc.res_array(a["drv_cpu"][:q], a["drv_mem"][:q], a["drv_gpu"][:q])
    exe = orc.res_array(a["exe_cpu"][:q], a["exe_mem"][:q], a["exe_gpu"][:q])
    count = a["count"][:q]
    times = []
    if w["groups"] == 1:
        cl = orc.Cluster(names, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"],
                         sched=(nodes["alloc_cpu"], nodes["alloc_mem"], nodes["alloc_gpu"]))
        onames = [names[i] for i in eorder]
        for _ in range(repeats):
            if w["mode"] == 0:
                t0 = time.perf_counter()
                cl.binpack_batch(ORC_ALGO[w["algo"]], drv, exe, count, onames, onames, with_efficiencies=True, n_threads=threads)
                times.append(time.perf_counter() - t0)
            else:
                cl = orc.Cluster(names, nodes["avail_cpu"], nodes["avail_mem"], nodes["avail_gpu"],
                                 sched=(nodes["alloc_cpu"], nodes["alloc_mem"], nodes["alloc_gpu"]))
                t0 = time.perf_counter()
                cl.fifo(ORC_ALGO[w["algo"]], w["mode"], drv, exe, count, a["young"][:q], onames, onames, with_efficiencies=True)
                times.append(time.perf_counter() - t0)
        used_threads = threads if w["mode"] == 0 else 1
    else:
        # FIFO per instance group: groups are independent queues -> one thread per group
        import concurrent.futures as cf
        def run_group(g):
            sel = np.nonzero(a["group"][:q] == g)[0]
            order = eorder[eoff[g]:eoff[g + 1]]
            sub_names = [names[i] for i in order]
            cl = orc.Cluster(sub_names, nodes["avail_cpu"][order], nodes["avail_mem"][order], nodes["avail_gpu"][order],
                             sched=(nodes["alloc_cpu"][order], nodes["alloc_mem"][order], nodes["alloc_gpu"][order]))
            if w["mode"] == 0:
                cl.binpack_batch(ORC_ALGO[w["algo"]], drv[sel], exe[sel], count[sel], sub_names, sub_names, True, 1)
            else:
                cl.fifo(ORC_ALGO[w["algo"]], w["mode"], drv[sel], exe[sel], count[sel], a["young"][:q][sel], sub_names, sub_names, True)
        used_threads = min(threads, w["groups"])
        for _ in range(repeats):
            t0 = time.perf_counter()
            with cf.ThreadPoolExecutor(used_threads) as ex:   # ctypes releases the GIL
                list(ex.map(run_group, range(w["groups"])))
            times.append(time.perf_counter() - t0)
    return q, used_threads, times


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    # warm-up + timed steps on the bounded sample
    q, threads, _ = cpu_reference_run(w, min(sample, 2000), cores, repeats=max(args.warmup, 1) if args.warmup else 0) \
        if args.warmup else (0, cores, [])
    q, threads, times = cpu_reference_run(w, sample, cores, repeats=args.steps)
    t = float(np.mean(times))
    value = q / t
    line = {
        "impl": "reference", "metric": "gang_placements_per_sec", "value": value, "unit": "decisions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": w["desc"], "nodes": w["nodes"], "apps_per_step_sample": q, "algo": ALGO_NAME[w["algo"]],
                   "mode": MODE_NAME[w["mode"]], "instance_groups": w["groups"]},
        "cpu_baseline": {"value": value, "unit": "decisions/s", "cores": threads, "kind": "port",
                         "sample": f"first {q} apps of the workload per step, literal C restatement of the Go path "
                                   f"(string-keyed maps, per-candidate map allocation, ComputePackingEfficiencies on), "
                                   f"{threads} host threads; Go toolchain absent so the reference itself cannot run"},
        "e2e": {"value": value, "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def _watchdog(seconds: float):
    """A hung collective must not hang the caller: hard-exit with an error after `seconds`."""
    def fire():
        print(f"[bench] watchdog: no result after {seconds:.0f} s, aborting", file=sys.stderr, flush=True)
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tightly-100k", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="apps per CPU-baseline step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-seconds", type=float, default=900.0, help="watchdog: abort if the run takes longer")
    args = ap.parse_args()
    _watchdog(args.max_seconds)
    w = WORKLOADS[args.workload]
    if args.cpu_sample == 0:
        # ~10-30 s of CPU work: ~0.2-0.5 ms per decision per thread for the literal port
        args.cpu_sample = min(w["apps"], 40000 if w["mode"] == 0 else 4000)

    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import torch
    import torch.distributed as dist
    import k8s_spark_scheduler_b200 as g
    if not os.path.exists(g.native.LIB_PATH):      # normally prebuilt in-tree by __graft_entry__.build()
        g.native.build()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the library has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    packer = g.GangPacker(device=local_rank)
    stream = torch.cuda.ExternalStream(packer.stream_handle(), device=dev)
    nodes, a, eoff, eorder = make_workload(w, rank)
    q = len(a["count"])
    total_exec = int(a["off"][-1])
    algo, mode = w["algo"], w["mode"]
    n_nodes, n_ord = w["nodes"], len(eorder)

    def dev_t(x, dtype):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dtype).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ================= value: device-resident inputs, no collective in the data path =====================
    # Every rank holds the snapshot and its own block of the queue in HBM; a step = snapshot layout + prep + pack.
    # (The path shards by application with nothing to exchange while packing; distributing the snapshot and
    #  collecting the placements are the multi-GPU analogue of H2D / D2H and are timed in `e2e` below.)
    with torch.cuda.stream(stream):
        # the snapshot lives in ONE flat buffer so that a single NCCL broadcast can move it (e2e):
        # [cpu int64 x N | mem int64 x N | gpu int64 x N | executor/driver order int32 x len(eorder)]
        snapbuf = torch.zeros(3 * n_nodes + (n_ord + 1) // 2, dtype=torch.int64, device=dev)
        tn = {"cpu": snapbuf[0:n_nodes], "mem": snapbuf[n_nodes:2 * n_nodes], "gpu": snapbuf[2 * n_nodes:3 * n_nodes],
              "eorder": snapbuf[3 * n_nodes:].view(torch.int32)[:n_ord], "eoff": dev_t(eoff, torch.int32)}
        tn["cpu"].copy_(dev_t(nodes["avail_cpu"], torch.int64)); tn["mem"].copy_(dev_t(nodes["avail_mem"], torch.int64))
        tn["gpu"].copy_(dev_t(nodes["avail_gpu"], torch.int64)); tn["eorder"].copy_(dev_t(eorder, torch.int32))
        ta = {k: dev_t(a[k], torch.int64 if a[k].dtype == np.int64 else (torch.uint8 if a[k].dtype == np.uint8 else torch.int32))
              for k in APP_KEYS}
        ta["off"] = dev_t(a["off"], torch.int64)
        if w["groups"] == 1:
            ta.pop("group")
        if mode == 0:
            ta.pop("young")
        d_driver = torch.empty(q, dtype=torch.int32, device=dev)
        d_exec = torch.empty(max(total_exec, 1), dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream.synchronize()

    def device_step():
        packer.set_snapshot_device(tn["cpu"], tn["mem"], tn["gpu"], tn["eoff"], tn["eorder"], tn["eoff"], tn["eorder"])
        packer.pack_batch_device(ta, algo, mode, d_driver, d_exec)

    launches_per_step = 3 + 2   # build_groups, build_exec_slots, build_driver_slots, prep_apps, pack
    with torch.cuda.stream(stream):
        # Eager warm-up steps: they also provide the per-kernel times (the library brackets its pack kernel with
        # CUDA events) and the scan statistics that the roofline object needs.
        pack_ns, prep_ns = [], []
        for _ in range(max(args.warmup, 3)):
            flush.fill_(1)
            device_step()
            st = packer.stats()                # synchronises the stream; reads the pack kernel's own event time
            pack_ns.append(st["pack_kernel_ns"]); prep_ns.append(st["prep_kernel_ns"])
        stats = packer.stats()
        # The step is a chain of ~10 small launches; issued from Python its duration depends on how fast the host
        # thread can enqueue them (visibly so with 8 ranks per box).  It is therefore captured ONCE into a CUDA
        # graph and replayed: same kernels, same work, one launch.  BENCH_GRAPH=0 keeps eager launches.
        graph = None
        if os.environ.get("BENCH_GRAPH", "1") != "0":
            try:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    device_step()
                for _ in range(2):
                    graph.replay()
                torch.cuda.synchronize()
            except Exception as e:   # capture not possible: run eagerly
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
                graph = None
                torch.cuda.synchronize()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        wall0 = time.perf_counter()
        for s in range(args.steps):
            flush.fill_(s & 0xff)              # L2 flush, outside the event pair
            evs[s][0].record(stream)
            if graph is not None:
                graph.replay()
            else:
                device_step()
            evs[s][1].record(stream)
        barrier()
        wall1 = time.perf_counter()
        step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    ms = float(np.mean(step_ms))
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    value = q * world / (ms * 1e-3)

    # ================= e2e: host buffers in, host results out ================================================
    # all-zero GPU request columns are passed as NULL (= 0), as the ABI allows; the shim knows while marshalling
    host_keys = [k for k in a if not (k in ("drv_gpu", "exe_gpu") and not a[k].any())]
    i64_cols = [k for k in ("drv_cpu", "drv_mem", "exe_cpu", "exe_mem", "drv_gpu", "exe_gpu") if k in host_keys]
    pin = packer.pinned_columns(q, i64_cols)          # one pinned block, column after column (as the shim allocates it)
    for k in host_keys:
        if k not in pin:
            pin[k] = packer.pinned(len(a[k]), a[k].dtype)
        pin[k][:] = a[k]
    if w["groups"] == 1:
        pin.pop("group", None)
    if mode == 0:
        pin.pop("young", None)
    # the snapshot SoA also lives in pinned host memory (what the shim fills per Predicate)
    pn = {}
    for k, v in (("cpu", nodes["avail_cpu"]), ("mem", nodes["avail_mem"]), ("gpu", nodes["avail_gpu"]),
                 ("eorder", eorder), ("eoff", eoff)):
        pn[k] = packer.pinned(len(v), v.dtype); pn[k][:] = v
    snap_bytes = 3 * 8 * n_nodes + 2 * 4 * n_ord + 2 * 4 * len(eoff)
    in_bytes = sum(v.nbytes for v in pin.values())

    if world == 1:
        out_driver = packer.pinned(q, np.int32)
        out_exec = packer.pinned(max(total_exec, 1), np.int32)
        h2d = in_bytes + snap_bytes
        d2h = out_driver.nbytes + 4 * total_exec

        def e2e_step():
            packer.set_snapshot(pn["cpu"], pn["mem"], pn["gpu"], pn["eorder"], pn["eorder"], pn["eoff"], pn["eoff"])
            packer.pack_batch(pin, algo, mode, out=(out_driver, out_exec))
            return int(out_driver[0])          # the host reads the result

        for _ in range(max(args.warmup, 3)):
            e2e_step()
        torch.cuda.synchronize()
        e2e_t = []
        e2e_launches = 0
        for s in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(s & 0xff)
            stream.synchronize()
            t0 = time.perf_counter()
            e2e_step()
            e2e_t.append(time.perf_counter() - t0)
            e2e_launches += 3 + packer.stats()["kernel_launches"]   # snapshot layout + (prep + pack) per pipelined chunk
        e2e_ms = float(np.mean(e2e_t)) * 1e3
        e2e_value = q / (e2e_ms * 1e-3)
        e2e_step_desc = "gp_set_snapshot + gp_pack_batch from pinned host buffers to host results"
    else:
        # N>1: rank 0 uploads the snapshot (H2D) and ONE NCCL broadcast distributes it; every rank packs
        # its own host-resident block of the queue (inputs read in place from mapped pinned memory) in 4 chunks, and
        # each chunk's placements ([driver | ExecutorNodes], padded to the largest rank so the collective is regular)
        # are all-gathered over NVLink asynchronously while the next chunk is packed; rank 0 finally copies every
        # rank's placements to its host memory.  Timed with wall clock between barriers, max over ranks.
        n_ch = 4 if mode == 0 else 1
        h_snap = torch.empty(snapbuf.numel(), dtype=torch.int64).pin_memory()
        h_snap.copy_(snapbuf.cpu())
        chunks = []
        with torch.cuda.stream(stream):
            for c in range(n_ch):
                lo, hi = (q * c) // n_ch, (q * (c + 1)) // n_ch
                e0, e1 = int(a["off"][lo]), int(a["off"][hi])
                t = torch.tensor([e1 - e0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); pad = max(int(t.item()), 1)
                res = torch.empty((hi - lo) + pad, dtype=torch.int32, device=dev)
                hin = {k: torch.from_numpy(v[lo:hi]) for k, v in pin.items() if k != "off"}      # mapped pinned views
                off_c = packer.pinned(hi - lo + 1, np.int64); off_c[:] = a["off"][lo:hi + 1] - e0
                hin["off"] = torch.from_numpy(off_c)
                gathered = torch.empty(res.numel() * world, dtype=torch.int32, device=dev)
                chunks.append({"apps": hin, "res": res, "driver": res[:hi - lo], "exec": res[hi - lo:], "gathered": gathered,
                               "host": torch.empty(gathered.numel(), dtype=torch.int32).pin_memory() if rank == 0 else None})
        stream.synchronize()
        h2d = in_bytes + (snapbuf.numel() * 8 if rank == 0 else 0)
        d2h = sum(ch["gathered"].numel() * 4 for ch in chunks) if rank == 0 else 0

        def e2e_step_multi():
            with torch.cuda.stream(stream):
                if rank == 0:
                    snapbuf.copy_(h_snap, non_blocking=True)
                dist.broadcast(snapbuf, src=0)
                packer.set_snapshot_device(tn["cpu"], tn["mem"], tn["gpu"], tn["eoff"], tn["eorder"], tn["eoff"], tn["eorder"])
                works = []
                for ch in chunks:
                    packer.pack_batch_device(ch["apps"], algo, mode, ch["driver"], ch["exec"])
                    works.append(dist.all_gather_into_tensor(ch["gathered"], ch["res"], async_op=True))
                for ch, wk in zip(chunks, works):
                    wk.wait()
                    if rank == 0:
                        ch["host"].copy_(ch["gathered"], non_blocking=True)
                stream.synchronize()
            return int(chunks[0]["host"][0]) if rank == 0 else 0

        for _ in range(max(args.warmup, 3)):
            e2e_step_multi()
        barrier()
        e2e_t = []
        for s in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(s & 0xff)
            barrier()
            t0 = time.perf_counter()
            e2e_step_multi()
            barrier()
            e2e_t.append(time.perf_counter() - t0)
        e2e_ms = float(np.mean(e2e_t)) * 1e3
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e_value = q * world / (e2e_ms * 1e-3)
        e2e_launches = (3 + 2 * n_ch) * args.steps
        e2e_step_desc = ("rank 0 H2D snapshot + NCCL broadcast + layout + per-rank pack of host-resident apps (4 chunks) + "
                         "async NCCL all-gather of placements overlapped with the next chunk + rank 0 D2H of all placements")
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (pack) ----------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    R = 2   # cpu + mem; the gpu array is skipped when no request and no negative availability
    k_total = total_exec
    alg_bytes = (stats["nodes_scanned"] * 8 * R + stats["drivers_tried"] * 4 + q * (64 + 8) + 8 * 0 + 4 * k_total)
    nominal_bytes = q * (w["nodes"] * 8 * R + w["nodes"] * 4 + 64 + 8) + 4 * k_total
    pack_s = float(np.mean(pack_ns)) * 1e-9
    kernel_name = f"gp_pack_{'independent' if mode == 0 else 'fifo_cta'}<{ALGO_NAME[algo]}>"
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(kernel_name)
        if t and t.get("workload") == args.workload and world == 1:
            traffic = t["dram_bytes_per_launch"]
    roofline = {
        "bound": "hbm", "kernel": kernel_name,
        "achieved": alg_bytes / pack_s / 1e9, "peak": peak, "unit": "GB/s",
        "frac": alg_bytes / pack_s / 1e9 / peak, "peak_source": peak_src,
        "traffic": traffic,
        "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": pack_s * 1e3,
        "nodes_scanned_per_decision": stats["nodes_scanned"] / q, "drivers_tried_per_decision": stats["drivers_tried"] / q,
        "full_table_equivalent_GBps": nominal_bytes / pack_s / 1e9,
        "note": "algorithmic bytes use the nodes actually visited (early exit is exact); the snapshot is served "
                "from L1/L2, so DRAM traffic is far below this -- see DESIGN.md section 6",
    }

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only) ------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        qs, threads, times = cpu_reference_run(w, args.cpu_sample, cores, repeats=1)
        cpu_baseline = {"value": qs / times[0], "unit": "decisions/s", "cores": threads, "kind": "port",
                        "sample": f"first {qs} apps of the workload, literal C restatement of the Go path (string-keyed maps, "
                                  f"per-candidate map allocation, ComputePackingEfficiencies on), {threads} threads, "
                                  f"{times[0]:.2f} s"}

    if rank == 0:
        line = {
            "metric": "gang_placements_per_sec", "value": value, "unit": "decisions/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": w["desc"], "nodes": w["nodes"], "apps_per_gpu": q, "apps_total": q * world,
                       "algo": ALGO_NAME[algo], "mode": MODE_NAME[mode], "instance_groups": w["groups"],
                       "executors_total_per_gpu": total_exec,
                       "l2": "256 MiB write between steps, outside the per-step CUDA-event pair",
                       "step": "snapshot layout + prep + pack, device-resident inputs, no collective in the data path "
                               "(every rank packs its own block of the queue against its copy of the snapshot); "
                               + ("the launch chain is replayed from one CUDA graph" if graph is not None else "eager launches"),
                       "e2e_step": e2e_step_desc,
                       "fits": None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "decisions/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches_per_step * args.steps + e2e_launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kernel_ms": {"pack": pack_s * 1e3, "prep": float(np.mean(prep_ns)) * 1e-6},
            "wall_s_value_region": wall1 - wall0,
        }
        print(json.dumps(line), flush=True)
    # Leave without tearing anything down: tensors allocated on the library's stream must not outlive that
    # stream (the caching allocator records events on it when they are freed), and destroying a process group
    # with outstanding async work can block.  Every rank has finished its last
    # collective (the all-reduce of the e2e time) at this point.
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()

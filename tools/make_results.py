#!/usr/bin/env python
"""Markdown table of bench.py lines:  python tools/make_results.py profiles/r02/bench_final/*.json"""
import json
import os
import sys


def row(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    name = os.path.basename(path)[6:-5]
    if d.get("impl") == "reference":
        return "| %s (CPU arm) | %.3g (%.1f ms per step) | - | - | - | - | %s | - |" % (name, d["value"], d["ms_per_step"], d["cpu_baseline"]["kind"])
    r = d.get("roofline") or {}
    cpu = d.get("cpu_baseline") or {}
    e = d["e2e"]
    return "| %s | %.3g (%.3f ms) | %.3g (%.3f ms; %.2f + %.2f MB) | %.3f / %.3f | %s | %.0f / %.0f / %.1f | %s | %s |" % (
        name, d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["h2d_bytes_per_step"] / 1e6, e["d2h_bytes_per_step"] / 1e6,
        d["kernel_ms"]["pack"], d["kernel_ms"]["prep"], r.get("scan_path_apps"), r.get("nodes_scanned_per_decision", 0),
        r.get("table_words_per_decision", 0), r.get("drivers_tried_per_decision", 0),
        ("%.3g (%d thr)" % (cpu["value"], cpu["cores"])) if cpu else "-",
        "%s / %s" % (d.get("parity_checked"), d.get("mismatches")))


print("| workload | `value` decisions/s (step) | `e2e` decisions/s (step; H2D + D2H) | pack / prep kernels ms | scan-path apps | nodes / table words / drivers per decision | CPU port decisions/s | parity checked / mismatches |")
print("|---|---|---|---|---|---|---|---|")
for p in sys.argv[1:]:
    try:
        print(row(p))
    except Exception as ex:      # not a bench line
        print("| %s | (unreadable: %s) |" % (os.path.basename(p), ex))

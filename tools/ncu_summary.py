#!/usr/bin/env python
"""Condense ncu output into the small text files committed under profiles/.

  tools/ncu_summary.py launches <launches.csv>            -> per-kernel count / mean us / share of the step
  tools/ncu_summary.py full <report.ncu-rep or raw.csv>   -> per-launch table of the metrics DESIGN.md cites
  tools/ncu_summary.py traffic <metrics.csv> [batch] [steps] [what] -> JSON: dram bytes (read + write) per step and stage,
                                                              the file bench.py reads roofline.traffic from (profiles/r2_traffic.json)
"""
import csv
import subprocess
import sys
from collections import OrderedDict

FULL = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    ki, vi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size")
    d = OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi:
            d.setdefault(r[ki].split("(")[0], []).append((float(r[vi].replace(",", "")), r[gi]))
    tot = sum(t for v in d.values() for t, _ in v)
    print("%-16s %6s %10s %8s  %s" % ("kernel", "count", "mean_us", "share", "grid (first)"))
    for k, v in d.items():
        s = sum(t for t, _ in v)
        print("%-16s %6d %10.1f %8.3f  %s" % (k, len(v), s / len(v) / 1e3, s / tot, v[0][1]))
    print("total_us %.1f over %d launches (cold-cache, serialised: compare shares, not absolutes)" % (tot / 1e3, sum(len(v) for v in d.values())))


def full(path):
    if path.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
    else:
        rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    cols = [(H.index(m), n) for m, n in FULL if m in H]
    print("%-14s " % "kernel" + " ".join("%14s" % n for _, n in cols))
    print("%-14s " % "" + " ".join("%14s" % U[i][:14] for i, _ in cols))
    for r in rows[2:]:
        print("%-14s " % r[ki].split("(")[0][:14] + " ".join("%14s" % r[i][:14] for i, _ in cols))


STAGE_OF = {"k_pyr_level0": "pyramid", "k_pyr_resize": "pyramid", "k_pyr_chain": "pyramid", "k_fast": "fast", "k_octree": "quadtree", "k_blur": "blur",
            "k_describe": "describe", "k_match_frame": "match", "k_unproject_last": "match"}


def _metric_log(path):
    """rows of an `ncu --metrics ... --csv --log-file` capture -> {(launch id, kernel): {metric: value in bytes / us / count}}"""
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    ki, mi, vi, ui, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
    d = OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi:
            v = float(r[vi].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "ms": 1e3}.get(r[ui], 1)
            d.setdefault((r[ii], r[ki].split("(")[0]), {})[r[mi]] = v
    return d


def traffic(path, batch=64, steps=2, what=""):
    """dram__bytes_read.sum + dram__bytes_write.sum per step and stage from a metrics capture that brackets `steps` whole steps
    (bench.py --profile-steps N under ncu --profile-from-start off): the JSON bench.py reads roofline.traffic from."""
    import json
    import os
    d = _metric_log(path)
    steps = int(steps)
    stages = {}
    for (_, name), m in d.items():
        st = STAGE_OF.get(name)
        if st:
            a = stages.setdefault(st, {"dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0, "launches_per_step": 0.0, "time_us": 0.0, "kernels": []})
            a["dram_read"] += m["dram__bytes_read.sum"] / steps
            a["dram_write"] += m["dram__bytes_write.sum"] / steps
            a["dram_bytes"] += (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / steps
            a["launches_per_step"] += 1.0 / steps
            a["time_us"] += m.get("gpu__time_duration.sum", 0.0) / steps
            if name not in a["kernels"]:
                a["kernels"].append(name)
    out = {"source": "profiles/%s: %s" % (os.path.basename(path), what), "batch": int(batch), "captured_steps": steps, "kernels": stages,
           "step_dram_bytes": sum(a["dram_bytes"] for a in stages.values())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

#!/usr/bin/env python
"""Condense ncu output into the small text files committed under profiles/.

  tools/ncu_summary.py launches <launches.csv>            -> per-kernel count / mean us / share of the step
  tools/ncu_summary.py full <report.ncu-rep or raw.csv>   -> per-launch table of the metrics DESIGN.md cites
  tools/ncu_summary.py traffic <report.ncu-rep> [batch] [steps] -> JSON: dram bytes per launch (read + write) of every kernel of one step,
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


def traffic(path, batch=64, steps=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, per kernel (mean over the captured launches of that kernel), and per
    stage of one step (a stage = its kernels x launches per step; the pyramid's launches are distinct levels, so they are summed over
    one step's worth = captured launches / captured steps, taken from the count of k_fast launches)."""
    import json
    import os
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U = rows[0], rows[1]
    ki, ri, wi = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")

    def to_bytes(v, unit):
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    per = OrderedDict()
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        per.setdefault(name, []).append(to_bytes(r[ri], U[ri]) + to_bytes(r[wi], U[wi]))
    steps = int(steps) if steps else max(len(per.get("k_fast", [1])), 1)      # the capture brackets whole steps (bench.py --profile-steps N)
    stages = {}
    for name, v in per.items():
        st = STAGE_OF.get(name)
        if st:
            d = stages.setdefault(st, {"dram_bytes": 0.0, "launches_per_step": 0.0, "kernels": []})
            d["dram_bytes"] += sum(v) / steps
            d["launches_per_step"] += len(v) / steps
            d["kernels"].append(name)
    out = {"source": "profiles/%s (ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per launch at batch %d, "
                     "summed over a stage's launches of one step)" % (os.path.basename(path).replace(".ncu-rep", "_full.txt"), int(batch)),
           "batch": int(batch), "captured_steps": steps, "kernels": stages, "step_dram_bytes": sum(d["dram_bytes"] for d in stages.values())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

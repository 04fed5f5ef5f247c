#!/usr/bin/env python
"""Condense ncu output into the small text files committed under profiles/.

  tools/ncu_summary.py launches <launches.csv>            -> per-kernel count / mean us / share of the step
  tools/ncu_summary.py full <report.ncu-rep or raw.csv>   -> per-launch table of the metrics DESIGN.md cites
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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

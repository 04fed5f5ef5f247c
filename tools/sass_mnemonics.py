#!/usr/bin/env python
"""profiles/r2_sass_mnemonics.txt: per kernel of liborbx.so, how often the SASS mnemonics occur that show how the hot path maps onto
Blackwell (cuobjdump -sass; no GPU needed).  Usage: tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "active-orb-slam2_b200", "liborbx.so")
COLS = ["UTMALDG", "UTMASTG", "UBLKCP", "VABSDIFF4", "VIMNMX3", "VIMNMX", "IDP.4A", "IDP.2A", "PRMT", "POPC", "DMMA", "HMMA", "DFMA", "SHFL", "REDUX",
        "BAR.SYNC", "UCGABAR", "ATOMS", "ATOMG", "SYNCS", "MEMBAR", "CCTL"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern = OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", d).replace("void ", "")
        kern[cur] = {c: 0 for c in COLS}
        kern[cur]["n"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["n"] += 1
        for c in COLS:
            if op.startswith(c):
                if c == "VIMNMX" and op.startswith("VIMNMX3"):
                    continue
                kern[cur][c] += 1
print("# cuobjdump -sass liborbx.so (sm_100a), instruction mnemonics per kernel that show how the hot path maps onto Blackwell (round 2, final binary; tools/sass_mnemonics.py)")
print("# UTMALDG = TMA tile load (cp.async.bulk.tensor), SYNCS = mbarrier, VABSDIFF4 / VIMNMX(3).U16x2 = packed byte / halfword SIMD, IDP.4A / IDP.2A = DP4A / DP2A,")
print("# DMMA = mma.sync f64 (the measured-and-dropped tensor form of the Schur pair accumulation), UCGABAR = cluster barrier, CCTL = L1 invalidation after a cluster- / device-scope fence")
print("%-34s" % "kernel" + "".join("%10s" % c for c in COLS) + "   SASS instructions")
for k, d in kern.items():
    print("%-34s" % k[:34] + "".join("%10d" % d[c] for c in COLS) + "   %d" % d["n"])

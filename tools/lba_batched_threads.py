#!/usr/bin/env python
"""batched LocalBA (orbx_lba_solve_begin / _end, one handle per window): windows/s against the number of host threads that submit and
collect the windows (the C calls release the GIL) and the number of windows in flight"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from orbx import synth
from orbx.optimizer import Optimizer, pack_problem

out = {}
for NW in (16, 32):
    ops = [Optimizer(max_keyframes=32, max_points=4096, max_edges=20000, device=0) for _ in range(NW)]
    probs = [pack_problem(synth.lba_problem(100 + i, n_kf=20, n_pts=3000, stereo=False, n_fixed=1)) for i in range(NW)]
    for nt in (1, 2, 4, 8):
        def work(t):
            mine = range(t, NW, nt)
            n = 0
            for _ in range(4):
                for i in mine:
                    ops[i].begin(probs[i])
                for i in mine:
                    n += ops[i].end()["trials"]
            return n
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(work, range(nt)))          # warm-up
            t0 = time.perf_counter()
            trials = sum(ex.map(work, range(nt)))
            dt = time.perf_counter() - t0
        out["windows_%d_threads_%d" % (NW, nt)] = round(4 * NW / dt, 1)
    for o in ops:
        o.close()
print(json.dumps(out))

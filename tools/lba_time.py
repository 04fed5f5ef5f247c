#!/usr/bin/env python
"""LocalBundleAdjustment timing through orbx_lba_solve_host: the C3 window (20 keyframes, 3000 points, mono) and a window of the
config-4 replay's shape (10 keyframes, 1500 points, stereo), with the cluster kernel's phase timers.  Optionally a pose call between
the windows (argument `mix`), as the replay does."""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from orbx import synth
from orbx.optimizer import Optimizer, PoseOptimizer, pack_problem

mix = "mix" in sys.argv[1:]
po = PoseOptimizer(max_observations=2048, max_frames=1, device=0)
pk = po.pack([synth.pose_problem(77, n=700)])
out = {}
for name, kw in (("c3", dict(n_kf=20, n_pts=3000, stereo=False, n_fixed=1)), ("c4_shape", dict(n_kf=10, n_pts=1500, stereo=True, n_fixed=2))):
    op = Optimizer(max_keyframes=64, max_points=16384, max_edges=131072, device=0)
    prob = synth.lba_problem(5, **kw)
    p = pack_problem(prob)
    for _ in range(3):
        op.begin(p); r = op.end()
    t0 = time.perf_counter()
    for _ in range(20):
        if mix:
            po.call(pk)
        op.begin(p); r = op.end()
    dt = (time.perf_counter() - t0) / 20
    op.LocalBundleAdjustment(prob)
    t0 = time.perf_counter()
    for _ in range(20):
        if mix:
            po.call(pk)
        r2 = op.LocalBundleAdjustment(prob)
    dt2 = (time.perf_counter() - t0) / 20
    out[name + "_solve_host"] = {"ms_per_window_with_packing": 1e3 * dt2, "phase_us": {k: round(float(v), 1) for k, v in op.phase_us().items()}}
    out[name] = {"ms_per_window": 1e3 * dt, "trials": int(r["trials"]), "phase_us": {k: round(float(v), 1) for k, v in op.phase_us().items()}}
    op.close()
print(json.dumps(out))

#!/usr/bin/env python
"""A/B of the pose-optimisation kernels (ORBX_POSE_KERNEL = 1: the one-block form, otherwise the cluster form) on the GPU:
batch of 64 frames and a single frame through orbx_pose_optimize_host, outlier sets checked against the oracle.  Prints one JSON line
per variant.  Usage: python tools/pose_ab.py            (spawns itself once per variant)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import bench
    from orbx import synth
    from orbx.optimizer import PoseOptimizer
    out = {"kernel": os.environ.get("ORBX_POSE_KERNEL", "default")}
    b = bench.bench_pose(0, True)
    out["batch64_ms_with_packing"] = b["ms_per_batch"]
    out["trials_per_frame"] = b["lm_trials_per_frame"]
    probs = [synth.pose_problem(1000 + f, n=500) for f in range(64)]
    po = PoseOptimizer(max_observations=64 * 500, max_frames=64, device=0)
    pk = po.pack(probs)
    po.call(pk)
    t0 = time.perf_counter()
    for _ in range(50):
        po.call(pk)
    out["batch64_ms"] = 1e3 * (time.perf_counter() - t0) / 50
    po.close()
    for n in (400, 1500):
        prob = [synth.pose_problem(77, n=n)]
        po = PoseOptimizer(max_observations=n, max_frames=1, device=0)
        pk = po.pack(prob)
        r = po.run(pk)
        t0 = time.perf_counter()
        for _ in range(200):
            po.call(pk)
        out["single_n%d_ms" % n] = 1e3 * (time.perf_counter() - t0) / 200
        out["single_n%d_trials" % n] = int(r[0]["trials"])
        po.close()
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one()
    else:
        for k in ("1", "0"):
            env = dict(os.environ, ORBX_POSE_KERNEL=k)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, check=False)

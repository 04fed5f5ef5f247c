#!/usr/bin/env python
"""one PoseOptimization call (one frame, 400 observations) through orbx_pose_optimize_host: the workload ncu captures for the pose kernel"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from orbx import synth
from orbx.optimizer import PoseOptimizer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
po = PoseOptimizer(max_observations=n, max_frames=1, device=0)
pk = po.pack([synth.pose_problem(77, n=n)])
for _ in range(3):
    po.call(pk)
print(po.run(pk)[0]["trials"])

"""Handles on two devices of ONE process (the reference runs one process per camera, a serving box may not): per-device kernel
attributes (raised shared-memory limits, cluster sizes) have to be set for every device a handle lives on.  Skipped on a single-GPU box."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_n_devices() < 2")
def test_pose_and_local_ba_on_the_second_device():
    from orbx.optimizer import Optimizer, PoseOptimizer
    for dev in (0, 1):
        po = PoseOptimizer(max_observations=64 * 1500, max_frames=64, device=dev)
        probs = [synth.pose_problem(300 + f, n=1500) for f in range(64)]       # 64 frames x 1500: 90 KB of shared memory per CTA
        rs = po.PoseOptimization(probs)
        for f in (0, 63):
            ref = O.pose_optimize(probs[f])
            assert np.array_equal(rs[f]["outlier"], ref["outlier"]) and rs[f]["n_inliers"] == ref["n_inliers"]
        po.close()
        op = Optimizer(max_keyframes=40, max_points=4000, max_edges=20000, device=dev)
        p = synth.lba_problem(2, n_kf=20, n_pts=3000, stereo=False, n_fixed=1)
        got, ref = op.LocalBundleAdjustment(p), O.lba_solve(p)
        assert got["trials"] == ref["trials"] and np.array_equal(got["erase"], ref["erase"])
        op.begin(p)
        got2 = op.end()
        assert got2["trials"] == ref["trials"] and np.array_equal(got2["erase"], ref["erase"])
        op.close()

"""Config C4 in miniature: the single-stream stereo tracking chain (extract x2 -> ComputeStereoMatches -> SearchByProjection(Cur,
Last) -> PoseOptimization) on a short rendered sequence, CUDA mirrors against the CPU oracle CALL BY CALL on identical inputs, plus
the sanity check that the chain actually tracks (poses near the ground truth)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import replay  # noqa: E402

pytestmark = pytest.mark.gpu


def test_chain_matches_oracle_call_by_call():
    seq = replay.StereoSequence(seed=0)
    gpu, orc = replay.GpuBackend(), replay.OracleBackend()
    rng = np.random.default_rng(42)
    last, rec = None, []
    n_frames = 5
    for t in range(n_frames):
        last = replay.track_frame(gpu, seq, t, last, rng, rec)
    calls = {"extract": 0, "stereo": 0, "match": 0, "pose": 0}
    for name, inp, out in rec:
        calls[name] += 1
        if name == "extract":
            kl, dl = orc.extract(0, inp[0]); kr, dr = orc.extract(1, inp[1])
            assert out[0].tobytes() == kl.tobytes() and np.array_equal(out[1], dl) and out[2].tobytes() == kr.tobytes() and np.array_equal(out[3], dr)
        elif name == "stereo":
            ur, dp = orc.stereo(*inp, seq.K[4], seq.K[5])          # the oracle extractors still hold this frame's pyramids
            assert out[0].tobytes() == ur.tobytes() and out[1].tobytes() == dp.tobytes() and (dp > 0).sum() > 200
        elif name == "match":
            cur, pts, ld, Tcw, fwd, bwd = inp
            n, m = orc.match_last(cur, pts, ld, Tcw[:3, :3], Tcw[:3, 3], fwd, bwd, 7.0)
            assert n == out[0] and np.array_equal(m, out[1]) and n > 150
        else:
            ref = orc.pose(inp)
            assert np.array_equal(ref["outlier"], out["outlier"]) and ref["n_inliers"] == out["n_inliers"]
            upd = np.linalg.norm((out["pose"] - inp["pose"]) - (ref["pose"] - inp["pose"])) / np.linalg.norm(ref["pose"] - inp["pose"])
            assert upd < 1e-4, upd
    assert calls == {"extract": n_frames, "stereo": n_frames, "match": n_frames - 1, "pose": n_frames - 1}
    # the chain tracks: the optimised pose of the last frame is within a centimetre of the truth although every start was ~1 cm off
    err = np.linalg.norm(last["Tcw"][:3, 3] - seq.true_pose(n_frames - 1)[:3, 3])
    assert err < 0.01 and last["n_inliers"] > 100, (err, last["n_inliers"])
    gpu.close()

"""The batched many-sequence mode of the C ABI (orbx_sequences_*, include/orbx.h; SURVEY.md §8e, BASELINE.json configs 2 and 5) against
the CPU oracle, call by call on identical inputs: every step's keypoints / descriptors are the oracle extractor's, mvuRight / mvDepth
are the oracle's Frame::ComputeStereoMatches, and the match arrays are the oracle's SearchByProjection(Cur, Last) on the last frame's
keypoints unprojected as Frame::UnprojectStereo does (restated here in float32 numpy, reference src/Frame.cc:290-296, 695-709)."""
import numpy as np
import pytest

from orbx import synth
from orbx._lib import KP_DTYPE
from orbx.sequences import Sequences, keypoints_of

pytestmark = pytest.mark.gpu
f32 = np.float32


def unproject(kp, depth, Tcw, K):
    """Frame::UnprojectStereo for every keypoint (float32, products summed left to right like cv::Mat)"""
    from oracle import oracle_py as O
    fx, fy, cx, cy = (f32(v) for v in K[:4])
    R, t = Tcw[:3, :3].astype(f32), Tcw[:3, 3].astype(f32)
    Rwc = R.T.copy()
    Ow = np.array([-((R[0, r] * t[0] + R[1, r] * t[1]) + R[2, r] * t[2]) for r in range(3)], f32)
    z = depth.astype(f32)
    x = (kp["x"] - cx) * z * (f32(1.0) / fx)
    y = (kp["y"] - cy) * z * (f32(1.0) / fy)
    pts = np.zeros(len(kp), O.LAST_POINT_DTYPE)
    for a, name in enumerate("xyz"):
        pts[name] = ((Rwc[a, 0] * x + Rwc[a, 1] * y) + Rwc[a, 2] * z) + Ow[a]
    ok = z > 0
    pts["angle"], pts["octave"], pts["valid"], pts["blocks"] = kp["angle"], kp["octave"], ok, ok
    for name in ("x", "y", "z", "angle"):
        pts[name][~ok] = 0
    pts["octave"][~ok] = 0
    return pts


def flags(Tcw, Tlw, b, mono):
    """bForward / bBackward, ORBmatcher.cc:1340-1351"""
    R, t = Tcw[:3, :3].astype(f32), Tcw[:3, 3].astype(f32)
    twc = np.array([-((R[0, r] * t[0] + R[1, r] * t[1]) + R[2, r] * t[2]) for r in range(3)], f32)
    tlc2 = ((Tlw[2, 0] * twc[0] + Tlw[2, 1] * twc[1]) + Tlw[2, 2] * twc[2]) + Tlw[2, 3]
    return bool(tlc2 > f32(b) and not mono), bool(-tlc2 > f32(b) and not mono)


def pose_of(tx, ty, yaw):
    """Tcw of PlaneWorld.render's camera: centre (tx, ty, 0), yaw about y"""
    cs, sn = np.cos(yaw), np.sin(yaw)
    Rwc = np.array([[cs, 0, sn], [0, 1, 0], [-sn, 0, cs]])
    T = np.eye(4)
    T[:3, :3] = Rwc.T
    T[:3, 3] = -Rwc.T @ np.array([tx, ty, 0.0])
    return T.astype(f32)


@pytest.mark.parametrize("stereo,n_sub,pinned", [(False, 1, False), (True, 1, False), (False, 2, True), (True, 3, True)])
def test_steps_match_the_oracle_chain(stereo, n_sub, pinned):
    """pinned: the caller's output arrays are page-locked (results land there by DMA); otherwise ordinary numpy arrays (results go
    through the handle's pinned block and are copied out in _step_end)"""
    from oracle import oracle_py as O
    import torch
    pin = (lambda shape, dtype: torch.zeros(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True).numpy()) if pinned else None
    w, h, ns, steps = 640, 480, 3, 4
    K = synth.TUM1_K
    worlds = [synth.stereo_world(s) for s in range(2)]
    sf = synth.scale_factors(8)
    seq = Sequences(ns, w, h, K, stereo=stereo, th=7.0, mono=False, const_depth=0.0 if stereo else 4.0, n_sub=n_sub, pose=True)
    is2 = (f32(1.0) / (sf * sf)).astype(f32)
    n_pose = 0
    cap = seq.capacity
    oex = [O.Extractor(1000, 1.2, 8, 20, 7) for _ in range(2)]
    sf = synth.scale_factors(8)
    last = [None] * ns
    total = 0
    rng = np.random.default_rng(5)
    for t in range(steps):
        imgs, poses = [], []
        for q in range(ns):
            wd = worlds[q % 2]
            tx, ty, yaw = 0.02 * t + 0.1 * q, 0.01 * q, 0.002 * t
            imgs.append(wd.render(tx, ty, yaw))
            if stereo:
                imgs.append(wd.render(tx, ty, yaw, right=True))
            # the pose the step is given = the motion-model prediction: the true pose off by ~5 mm / 0.1 deg
            poses.append(pose_of(tx + rng.normal(0, 0.005), ty + rng.normal(0, 0.005), yaw + np.deg2rad(rng.normal(0, 0.1))))
        out = seq.step(np.stack(imgs), np.stack([p[:3] for p in poses]), seq.alloc_outputs(pin))
        per = 2 if stereo else 1
        for q in range(ns):
            kl, dl = oex[0](imgs[per * q])
            gk, gd = keypoints_of(out, per * q)
            assert gk.tobytes() == kl.tobytes() and np.array_equal(gd, dl)
            if stereo:
                kr, dr = oex[1](imgs[per * q + 1])
                rk, rd = keypoints_of(out, per * q + 1)
                assert rk.tobytes() == kr.tobytes() and np.array_equal(rd, dr)
                tb = oex[0].tables()
                st = O.stereo_matches(kl, dl, kr, dr, [oex[0].level(l) for l in range(8)], [oex[1].level(l) for l in range(8)], tb["scale"],
                                      tb["inv_scale"], K[4], K[4] / K[0])
                n = len(kl)
                assert out["u_right"][q, :n].tobytes() == st["u_right"].tobytes() and out["depth"][q, :n].tobytes() == st["depth"].tobytes()
                ur, depth = st["u_right"], st["depth"]
            else:
                ur, depth = None, np.full(len(kl), 4.0, f32)
            if last[q] is None:
                assert out["nmatches"][q] == 0 and (out["match"][q] == -1).all()
            else:
                lk, ld, ldepth, lT = last[q]
                pts = unproject(lk, ldepth, lT, K)
                cur = dict(keys_un=kl, desc=dl, u_right=ur, claimed=None, bounds=(0.0, 0.0, float(w), float(h)), K=K, scale_factors=sf)
                fwd, bwd = flags(poses[q], lT, K[4] / K[0], False)
                nref, mref = O.search_by_projection_frame(cur, pts, ld, poses[q][:3, :3], poses[q][:3, 3], fwd, bwd, 7.0, True)
                assert out["nmatches"][q] == nref and np.array_equal(out["match"][q, :len(kl)], mref) and (out["match"][q, len(kl):] == -1).all()
                total += nref
                # Optimizer::PoseOptimization from that match array, starting at the pose the search projected with
                m = np.nonzero(mref >= 0)[0]
                prob = dict(Xw=np.stack([pts["x"], pts["y"], pts["z"]], 1)[mref[m]].astype(np.float64),
                            obs=np.stack([kl["x"][m], kl["y"][m], ur[m] if ur is not None else np.full(len(m), -1, f32)], 1).astype(np.float64),
                            inv_sigma2=is2[kl["octave"][m]], pose=O.to_se3quat(np.vstack([poses[q][:3], [0, 0, 0, 1]])), K=K[:5])
                pref = O.pose_optimize(prob)
                assert out["n_inliers"][q] == pref["n_inliers"] and np.array_equal(out["outlier"][q][m], pref["outlier"])
                assert not out["outlier"][q][:len(mref)][mref < 0].any() and not out["outlier"][q][len(mref):].any()
                upd = np.linalg.norm(pref["pose"] - prob["pose"])
                assert np.linalg.norm(out["pose"][q] - pref["pose"]) <= 1e-4 * upd + 1e-7
                n_pose += 1
            last[q] = (kl.copy(), dl.copy(), depth.copy(), poses[q])
    assert total > 100 * ns * (steps - 1) * (0.3 if stereo else 1.0), total
    assert n_pose == ns * (steps - 1)
    seq.reset()
    out = seq.step(np.stack(imgs), np.stack([p[:3] for p in poses]))
    assert (out["nmatches"] == 0).all()
    seq.close()

"""Replay harness for the single-stream stereo tracking chain: per frame
    ORBextractor x2 -> Frame::ComputeStereoMatches -> ORBmatcher::SearchByProjection(Cur, Last, th, mono=false)
    -> Optimizer::PoseOptimization
with the bookkeeping Tracking does between those calls (unprojection of the last frame's stereo points, Tracking.cc /
Frame::UnprojectStereo Frame.cc:671-685) done here in numpy.  The chain is written once against a small backend interface
and runs on the CUDA mirrors (GpuBackend) or on the CPU oracle (OracleBackend, test / baseline only); tests feed both the same
inputs call by call, bench.py times them.  This is test / measurement scaffolding, not part of the product library."""
import numpy as np

from orbx import synth
from orbx._lib import KP_DTYPE
from orbx.matcher import LAST_POINT_DTYPE, motion_flags

NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH = 1000, 1.2, 8, 20, 7


class GpuBackend:
    def __init__(self, w=640, h=480, device=0):
        from orbx.extractor import ORBextractor
        from orbx.matcher import ORBmatcher
        from orbx.optimizer import PoseOptimizer
        from orbx.stereo import StereoMatcher
        self.ex = [ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=w, max_height=h, max_batch=1, device=device) for _ in range(2)]
        self.mt = ORBmatcher(0.9, True, max_keypoints=4096, max_points=4096, device=device)
        self.st = StereoMatcher(max_keypoints=4096, device=device)
        self.po = PoseOptimizer(max_observations=4096, max_frames=1, device=device)
        self.scale, self.inv_sigma2 = self.ex[0].GetScaleFactors(), self.ex[0].GetInverseScaleSigmaSquares()
        self.seconds = dict(extract=0.0, stereo=0.0, match=0.0, pose=0.0)      # wall time spent inside the backend calls

    def _timed(self, key, fn, *a):
        import time
        t0 = time.perf_counter()
        r = fn(*a)
        self.seconds[key] += time.perf_counter() - t0
        return r

    def extract(self, side, img):
        return self._timed("extract", self.ex[side], img)

    def stereo(self, kl, dl, kr, dr, bf, b):
        ur, dp, _ = self._timed("stereo", self.st.ComputeStereoMatchesFromExtractors, self.ex[0], self.ex[1], len(kl), bf, b)
        return ur, dp

    def match_last(self, cur, pts, desc, R, t, fwd, bwd, th):
        return self._timed("match", self.mt.SearchByProjectionLast, cur, pts, desc, R, t, fwd, bwd, th)

    def pose(self, prob):
        return self._timed("pose", self.po.PoseOptimization, prob)

    def close(self):
        for h in (*self.ex, self.mt, self.st, self.po):
            h.close()


class OracleBackend:
    def __init__(self):
        from oracle import oracle_py as O
        self.O = O
        self.ex = [O.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH) for _ in range(2)]
        t = self.ex[0].tables()
        self.scale, self.inv_scale, self.inv_sigma2 = t["scale"], t["inv_scale"], t["inv_sigma2"]

    def extract(self, side, img):
        return self.ex[side](img)

    def stereo(self, kl, dl, kr, dr, bf, b):
        r = self.O.stereo_matches(kl, dl, kr, dr, [self.ex[0].level(l) for l in range(NLEVELS)], [self.ex[1].level(l) for l in range(NLEVELS)],
                                  self.scale, self.inv_scale, bf, b)
        return r["u_right"], r["depth"]

    def match_last(self, cur, pts, desc, R, t, fwd, bwd, th):
        return self.O.search_by_projection_frame(cur, pts, desc, R, t, fwd, bwd, th, True)

    def pose(self, prob):
        return self.O.pose_optimize(prob)

    def close(self):
        pass


def quat_pose(R, t):
    p = np.zeros(7)
    p[:4] = synth._quat_from_R(np.asarray(R, np.float32).astype(np.float64))     # Converter::toSE3Quat: float matrix -> double quaternion
    p[4:] = np.asarray(t, np.float32)
    return p


def pose_matrix(p):
    from scipy.spatial.transform import Rotation
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = Rotation.from_quat(p[:4]).as_matrix().astype(np.float32)          # Converter::toCvMat(SE3Quat): double -> float
    T[:3, 3] = p[4:].astype(np.float32)
    return T


class StereoSequence:
    """the camera of synth.PlaneWorld translating along x by `step` metres per frame (planes parallel to the image: rigid scene)"""

    def __init__(self, seed=0, w=640, h=480, step=0.02):
        self.world = synth.stereo_world(seed, w, h)
        self.w, self.h, self.step = w, h, step
        wd = self.world
        self.K = (np.float32(wd.fx), np.float32(wd.fy), np.float32(wd.cx), np.float32(wd.cy), np.float32(wd.bf), np.float32(wd.bf / wd.fx))

    def images(self, t):
        return self.world.render(self.step * t, 0.0, 0.0), self.world.render(self.step * t, 0.0, 0.0, right=True)

    def true_pose(self, t):
        T = np.eye(4, dtype=np.float32)
        T[0, 3] = -self.step * t
        return T


def track_frame(be, seq, t, last, rng, record=None):
    """one frame of the chain on backend `be`; `last` = state returned for frame t-1 (None for the first).  `record` collects the
    inputs and outputs of every backend call (for call-by-call comparison with another backend)."""
    fx, fy, cx, cy, bf, b = seq.K
    left, right = seq.images(t)
    kl, dl = be.extract(0, left)
    kr, dr = be.extract(1, right)
    ur, depth = be.stereo(kl, dl, kr, dr, bf, b)
    if record is not None:
        record.append(("extract", (left, right), (kl, dl, kr, dr)))
        record.append(("stereo", (kl, dl, kr, dr), (ur, depth)))
    state = dict(kl=kl, dl=dl, ur=ur, depth=depth, Tcw=seq.true_pose(t), n_match=0, n_inliers=0)
    if last is not None:
        # motion-model initial guess: the true pose off by ~1 cm / 0.2 deg (Tracking::TrackWithMotionModel starts from mVelocity * last pose)
        Tcw = seq.true_pose(t).copy()
        Tcw[:3, 3] += rng.normal(0, 0.01, 3).astype(np.float32)
        Tcw[:3, :3] = (synth._rot(1, np.deg2rad(rng.normal(0, 0.2))) @ Tcw[:3, :3].astype(np.float64)).astype(np.float32)
        # the last frame's stereo points in the world: Frame::UnprojectStereo (Frame.cc:671-685), float
        lk, ld, lz = last["kl"], last["dl"], last["depth"]
        ok = lz > 0
        Tl = last["Tcw"]
        Rwc, Ow = Tl[:3, :3].T, -(Tl[:3, :3].T @ Tl[:3, 3])
        xc = (lk["x"] - cx) * lz / fx
        yc = (lk["y"] - cy) * lz / fy
        Xw = (Rwc @ np.stack([xc, yc, lz]).astype(np.float32)).T + Ow
        pts = np.zeros(len(lk), LAST_POINT_DTYPE)
        pts["x"], pts["y"], pts["z"] = Xw[:, 0], Xw[:, 1], Xw[:, 2]
        pts["angle"], pts["octave"], pts["valid"], pts["blocks"] = lk["angle"], lk["octave"], ok, 1
        cur = dict(keys_un=kl, desc=dl, u_right=ur, claimed=None, bounds=(0.0, 0.0, float(seq.w), float(seq.h)), K=seq.K, scale_factors=be.scale)
        fwd, bwd = motion_flags(Tcw, Tl, b, False)
        n, match = be.match_last(cur, pts, ld, Tcw[:3, :3], Tcw[:3, 3], fwd, bwd, 7.0)
        idx = np.nonzero(match >= 0)[0]
        prob = dict(Xw=Xw[match[idx]].astype(np.float64), obs=np.stack([kl["x"][idx], kl["y"][idx], ur[idx]], 1).astype(np.float64),
                    inv_sigma2=be.inv_sigma2[kl["octave"][idx]], pose=quat_pose(Tcw[:3, :3], Tcw[:3, 3]), K=seq.K[:5])
        r = be.pose(prob)
        if record is not None:
            record.append(("match", (cur, pts, ld, Tcw.copy(), fwd, bwd), (n, match)))
            record.append(("pose", prob, r))
        state.update(Tcw=pose_matrix(r["pose"]), n_match=int(n), n_inliers=int(r["n_inliers"]))
    return state

"""Replay harness for BASELINE.json's config 4: a TUM-RGBD-shaped sequence (640x480, 2000 frames) through the calls Tracking and
LocalMapping make on the hot path, with a live map kept here in numpy:

  every frame      ORBextractor x2, Frame::ComputeStereoMatches                                    (Tracking::GrabImageStereo)
                   ORBmatcher::SearchByProjection(Cur, Last, th, false) + Optimizer::PoseOptimization   (TrackWithMotionModel, Tracking.cc:857-880)
                   Frame::isInFrustum over the local map, ORBmatcher::SearchByProjection(Frame, local map points, th)
                   + Optimizer::PoseOptimization                                                    (TrackLocalMap, Tracking.cc:1011-1060, :1338-1372)
  every kf_every-th frame = keyframe
                   new map points from the stereo depths (Tracking::CreateNewKeyFrame, Tracking.cc:1168-1236), vocabulary transform
                   (KeyFrame::ComputeBoW), ORBmatcher::SearchForTriangulation against up to 10 earlier keyframes
                   (LocalMapping::CreateNewMapPoints, LocalMapping.cc:206-272), Optimizer::LocalBundleAdjustment over the last 10
                   keyframes (+ the fixed observers of their points) (LocalMapping.cc:81)

The ground-truth pose plus noise stands in for the motion model; everything the calls return is used (matches feed the pose
optimisation, its pose and inliers feed the next call, LocalBA moves keyframes and points that the next frames are matched against).
One backend interface, two implementations: the CUDA mirrors (GpuBackend) and the CPU oracle (OracleBackend, test / baseline only);
`record` collects every call's inputs and outputs so that tests replay them on the other backend.  Test / measurement scaffolding,
not part of the product library."""
import numpy as np

import replay
from orbx import synth
from orbx.matcher import LAST_POINT_DTYPE, TRACK_POINT_DTYPE, motion_flags

f32 = np.float32


class GpuBackend(replay.GpuBackend):
    def __init__(self, voc_tree, w=640, h=480, device=0):
        super().__init__(w, h, device)
        from orbx.matcher import ORBmatcher
        from orbx.optimizer import Optimizer
        from orbx.vocabulary import ORBVocabulary
        self.device = device
        self.mt_map = ORBmatcher(0.8, True, max_keypoints=4096, max_points=16384, device=device)      # SearchLocalPoints: ORBmatcher matcher(0.8)
        self.mt_tri = ORBmatcher(0.6, False, max_keypoints=4096, max_points=4096, device=device)      # CreateNewMapPoints: ORBmatcher matcher(0.6, false)
        self.voc = ORBVocabulary(voc_tree, max_features=4096, device=device)
        self.lba_h = Optimizer(max_keyframes=64, max_points=16384, max_edges=131072, device=device)
        self.seconds.update(frustum=0.0, match_map=0.0, bow=0.0, triangulation=0.0, lba=0.0)

    def frustum(self, frame, pts):
        from orbx.frustum import isInFrustum
        return self._timed("frustum", lambda: isInFrustum(frame, pts, device=self.device)[0])

    def match_map(self, F, tp, desc, th):
        return self._timed("match_map", self.mt_map.SearchByProjection, F, tp, desc, th)

    def bow(self, desc):
        return self._timed("bow", self.voc.transform_features, desc, 4)

    def triangulation(self, KF1, KF2, F12, epipole, sigma2, scale):
        return self._timed("triangulation", self.mt_tri.SearchForTriangulation, KF1, KF2, F12, epipole, sigma2, scale, False)

    def lba(self, prob):
        return self._timed("lba", self.lba_h.LocalBundleAdjustment, prob)

    def close(self):
        super().close()
        for h in (self.mt_map, self.mt_tri, self.voc, self.lba_h):
            h.close()


class OracleBackend(replay.OracleBackend):
    def __init__(self, voc_tree):
        super().__init__()
        self.tree = voc_tree

    def frustum(self, frame, pts):
        return self.O.is_in_frustum(frame, pts)

    def match_map(self, F, tp, desc, th):
        return self.O.search_by_projection_points(F, tp, desc, th, 0.8)

    def bow(self, desc):
        return self.O.bow_transform(self.tree, desc, 4)

    def triangulation(self, KF1, KF2, F12, epipole, sigma2, scale):
        A, B = dict(KF1), dict(KF2)
        n, ma = self.O.match_buckets(2, A, B, 0.6, False, False, F12, epipole, sigma2, scale)
        idx1 = np.nonzero(ma >= 0)[0]
        return n, np.stack([idx1, ma[idx1]], 1)

    def lba(self, prob):
        return self.O.lba_solve(prob)


def make_vocabulary(seed=3, k=10, L=4):
    from orbx.vocabulary import tree_from_parents
    return tree_from_parents(*synth.random_vocabulary(seed, k=k, L=L))


class LoopSequence(replay.StereoSequence):
    """the PlaneWorld camera moving there and back along x: n_poses distinct views, frame t shows view tri(t)"""

    def __init__(self, seed=0, n_poses=40, step=0.02):
        super().__init__(seed, 640, 480, step)
        self.n_poses = n_poses
        self._img = {}

    def view(self, t):
        p = t % (2 * self.n_poses - 2)
        return p if p < self.n_poses else 2 * self.n_poses - 2 - p

    def images(self, t):
        v = self.view(t)
        if v not in self._img:
            self._img[v] = super().images(v)
        return self._img[v]

    def true_pose(self, t):
        return super().true_pose(self.view(t))


class Map:
    """map points and keyframes as flat arrays (what Map / MapPoint / KeyFrame hold for the calls above)"""

    def __init__(self, cap=200000):
        self.pos = np.zeros((cap, 3), f32)
        self.desc = np.zeros((cap, 32), np.uint8)
        self.normal = np.zeros((cap, 3), f32)
        self.min_d, self.max_d = np.zeros(cap, f32), np.zeros(cap, f32)
        self.n_obs = np.zeros(cap, np.int32)
        self.first_kf = np.zeros(cap, np.int32)
        self.n = 0
        self.kfs = []                                    # dicts: kl, dl, ur, depth, Tcw, mp (map point of every keypoint or -1), fv


def frustum_frame(Tcw, K, w, h, nlevels=8):
    from orbx.frustum import FRUSTUM_FRAME_DTYPE
    fr = np.zeros((), FRUSTUM_FRAME_DTYPE)
    R, t = Tcw[:3, :3].astype(f32), Tcw[:3, 3].astype(f32)
    fr["Rcw"], fr["tcw"] = R.reshape(9), t
    fr["Ow"] = np.array([-((R[0, r] * t[0] + R[1, r] * t[1]) + R[2, r] * t[2]) for r in range(3)], f32)     # mOw = -mRcw.t()*mtcw
    fr["fx"], fr["fy"], fr["cx"], fr["cy"], fr["bf"] = K[:5]
    fr["min_x"], fr["max_x"], fr["min_y"], fr["max_y"] = 0.0, float(w), 0.0, float(h)
    fr["log_scale_factor"] = f32(np.log(f32(1.2)))
    fr["n_levels"], fr["viewing_cos_limit"] = nlevels, 0.5
    return fr


def unproject(kp, depth, Tcw, K):
    """Frame::UnprojectStereo (Frame.cc:695-709) for every keypoint with depth > 0 -> (world points float32 [n, 3], ok mask)"""
    fx, fy, cx, cy = (f32(v) for v in K[:4])
    Rwc, Ow = Tcw[:3, :3].T.astype(f32), -(Tcw[:3, :3].T.astype(f32) @ Tcw[:3, 3].astype(f32))
    z = depth.astype(f32)
    xc = (kp["x"] - cx) * z / fx
    yc = (kp["y"] - cy) * z / fy
    return (Rwc @ np.stack([xc, yc, z]).astype(f32)).T + Ow, z > 0


def pose_problem(Xw, kl, ur, idx, inv_sigma2, Tcw, K):
    return dict(Xw=Xw.astype(np.float64), obs=np.stack([kl["x"][idx], kl["y"][idx], ur[idx]], 1).astype(np.float64),
                inv_sigma2=inv_sigma2[kl["octave"][idx]], pose=replay.quat_pose(Tcw[:3, :3], Tcw[:3, 3]), K=K[:5])


def fundamental(T1, T2, K):
    """LocalMapping::ComputeF12 (LocalMapping.cc:568-586): F12 = K1^-T [t12]x R12 K2^-1, float"""
    R1, t1, R2, t2 = T1[:3, :3].astype(f32), T1[:3, 3].astype(f32), T2[:3, :3].astype(f32), T2[:3, 3].astype(f32)
    R12 = R1 @ R2.T
    t12 = -R1 @ R2.T @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]], f32)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float64)
    Ki = np.linalg.inv(Km).astype(f32)
    return (Ki.T @ tx @ R12 @ Ki).astype(f32)


def epipole_in_2(T1, T2, K):
    """ORBmatcher.cc:663-670: camera centre of KF1 projected into KF2"""
    C1 = -(T1[:3, :3].T.astype(f32) @ T1[:3, 3].astype(f32))
    C2 = T2[:3, :3].astype(f32) @ C1 + T2[:3, 3].astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        iz = f32(1.0) / C2[2]
        return float(f32(K[0]) * C2[0] * iz + f32(K[2])), float(f32(K[1]) * C2[1] * iz + f32(K[3]))


def local_window(mp, kf_ids, max_fixed=10):
    """Optimizer.cc:456-505 on the flat map: local keyframes, the points they see, the earlier keyframes that also see them"""
    local_pts = sorted(set(int(p) for k in kf_ids for p in mp.kfs[k]["mp"] if p >= 0))
    pset = set(local_pts)
    fixed = []
    for k in range(len(mp.kfs) - 1, -1, -1):
        if k in kf_ids or len(fixed) >= max_fixed:
            continue
        if any(int(p) in pset for p in mp.kfs[k]["mp"] if p >= 0):
            fixed.append(k)
    return local_pts, fixed


def lba_problem(mp, kf_ids, K, inv_sigma2):
    local_pts, fixed = local_window(mp, kf_ids)
    kfs = list(kf_ids) + fixed
    pidx = {p: i for i, p in enumerate(local_pts)}
    e_kf, e_pt, e_obs, e_is2, e_st, e_ref = [], [], [], [], [], []
    for ki, k in enumerate(kfs):
        kf = mp.kfs[k]
        sel = np.nonzero(kf["mp"] >= 0)[0]
        for i in sel:
            p = int(kf["mp"][i])
            if p in pidx:
                e_kf.append(ki); e_pt.append(pidx[p])
                e_obs.append((kf["kl"]["x"][i], kf["kl"]["y"][i], kf["ur"][i]))
                e_is2.append(inv_sigma2[kf["kl"]["octave"][i]]); e_st.append(0 if kf["ur"][i] < 0 else 1); e_ref.append((k, int(i)))
    kf_pose = np.stack([replay.quat_pose(mp.kfs[k]["Tcw"][:3, :3], mp.kfs[k]["Tcw"][:3, 3]) for k in kfs])
    fixed_mask = np.array([0] * len(kf_ids) + [1] * len(fixed), np.uint8)
    if kfs[0] == 0 or 0 in kf_ids:
        fixed_mask[kfs.index(0)] = 1                       # vSE3->setFixed(pKFi->mnId == 0)
    prob = dict(kf_pose=kf_pose, kf_fixed=fixed_mask, pts=mp.pos[local_pts].astype(np.float64), e_kf=np.array(e_kf, np.int32),
                e_pt=np.array(e_pt, np.int32), e_obs=np.array(e_obs, np.float64).reshape(-1, 3), e_inv_sigma2=np.array(e_is2, f32),
                e_stereo=np.array(e_st, np.uint8), K=tuple(float(v) for v in K[:5]))
    return prob, kfs, local_pts, e_ref


class Replay:
    def __init__(self, be, seq, kf_every=10, seed=7, record=None):
        self.be, self.seq, self.kf_every, self.record = be, seq, kf_every, record
        self.rng = np.random.default_rng(seed)
        self.map = Map()
        self.last = None
        self.local_pts = None
        self.sigma2 = (be.scale * be.scale).astype(f32)
        self.stats = dict(frames=0, keyframes=0, a12=0, a11=0, tri_pairs=0, lba_windows=0, lba_trials=0, inliers=[], err=[])

    def _rec(self, name, inp, out):
        if self.record is not None:
            self.record.append((name, inp, out))

    def step(self, t):
        be, seq, mp = self.be, self.seq, self.map
        K = seq.K
        fx, fy, cx, cy, bf, b = K
        left, right = seq.images(t)
        kl, dl = be.extract(0, left)
        kr, dr = be.extract(1, right)
        ur, depth = be.stereo(kl, dl, kr, dr, bf, b)
        self._rec("extract", (left, right), (kl, dl, kr, dr))
        self._rec("stereo", (kl, dl, kr, dr), (ur, depth))
        Tcw = seq.true_pose(t).copy()
        kp_mp = np.full(len(kl), -1, np.int64)               # map point of every keypoint after TrackLocalMap
        n_inl = 0
        if self.last is not None:
            # motion-model guess: the true pose off by ~1 cm / 0.2 deg
            Tcw[:3, 3] += self.rng.normal(0, 0.01, 3).astype(f32)
            Tcw[:3, :3] = (synth._rot(1, np.deg2rad(self.rng.normal(0, 0.2))) @ Tcw[:3, :3].astype(np.float64)).astype(f32)
            lk, ld, lz, Tl = self.last["kl"], self.last["dl"], self.last["depth"], self.last["Tcw"]
            Xl, ok = unproject(lk, lz, Tl, K)
            pts = np.zeros(len(lk), LAST_POINT_DTYPE)
            pts["x"], pts["y"], pts["z"] = Xl[:, 0], Xl[:, 1], Xl[:, 2]
            pts["angle"], pts["octave"], pts["valid"], pts["blocks"] = lk["angle"], lk["octave"], ok, 1
            cur = dict(keys_un=kl, desc=dl, u_right=ur, claimed=None, bounds=(0.0, 0.0, float(seq.w), float(seq.h)), K=K, scale_factors=be.scale)
            fwd, bwd = motion_flags(Tcw, Tl, b, False)
            n12, m12 = be.match_last(cur, pts, ld, Tcw[:3, :3], Tcw[:3, 3], fwd, bwd, 7.0)
            self._rec("match", (cur, pts, ld, Tcw.copy(), fwd, bwd), (n12, m12))
            idx = np.nonzero(m12 >= 0)[0]
            prob = pose_problem(Xl[m12[idx]], kl, ur, idx, be.inv_sigma2, Tcw, K)
            r = be.pose(prob)
            self._rec("pose", prob, r)
            Tcw = replay.pose_matrix(r["pose"])
            self.stats["a12"] += int(n12)
            n_inl = int(r["n_inliers"])
            # ---- TrackLocalMap: the points of the last 10 keyframes ----
            if mp.kfs:
                from orbx.frustum import FRUSTUM_POINT_DTYPE
                if self.local_pts is None:               # the local map changes only when a keyframe is inserted / LocalBA erases observations
                    allp = np.concatenate([kf["mp"] for kf in mp.kfs[-10:]])
                    self.local_pts = np.unique(allp[allp >= 0]).astype(np.int64)
                lp = self.local_pts
                fp = np.zeros(len(lp), FRUSTUM_POINT_DTYPE)
                fp["x"], fp["y"], fp["z"] = mp.pos[lp, 0], mp.pos[lp, 1], mp.pos[lp, 2]
                fp["nx"], fp["ny"], fp["nz"] = mp.normal[lp, 0], mp.normal[lp, 1], mp.normal[lp, 2]
                fp["min_distance"], fp["max_distance"], fp["blocks"] = mp.min_d[lp], mp.max_d[lp], 1
                fr = frustum_frame(Tcw, K, seq.w, seq.h)
                tp = be.frustum(fr, fp)
                self._rec("frustum", (fr, fp), tp)
                n11, m11 = be.match_map(cur, tp, mp.desc[lp], 3.0)          # th = 3 for stereo / RGB-D (Tracking.cc:1364-1369)
                self._rec("match_map", (cur, tp, mp.desc[lp].copy()), (n11, m11))
                self.stats["a11"] += int(n11)
                idx = np.nonzero(m11 >= 0)[0]
                if len(idx) >= 10:
                    prob = pose_problem(mp.pos[lp[m11[idx]]], kl, ur, idx, be.inv_sigma2, Tcw, K)
                    r = be.pose(prob)
                    self._rec("pose", prob, r)
                    Tcw = replay.pose_matrix(r["pose"])
                    inl = idx[r["outlier"] == 0]
                    kp_mp[inl] = lp[m11[inl]]
                    n_inl = int(r["n_inliers"])
        state = dict(kl=kl, dl=dl, ur=ur, depth=depth, Tcw=Tcw)
        self.stats["frames"] += 1
        self.stats["inliers"].append(n_inl)
        self.stats["err"].append(float(np.linalg.norm(Tcw[:3, 3] - seq.true_pose(t)[:3, 3])))
        if t % self.kf_every == 0:
            self.keyframe(state, kp_mp)
        self.last = state
        return state

    def keyframe(self, st, kp_mp):
        be, mp, K = self.be, self.map, self.seq.K
        kl, dl, ur, depth, Tcw = st["kl"], st["dl"], st["ur"], st["depth"], st["Tcw"]
        kid = len(mp.kfs)
        # new map points from the stereo depths of keypoints without one: the closest first, all the close ones and at least 100
        # (Tracking::CreateNewKeyFrame, Tracking.cc:1186-1232; mThDepth = 35 baselines)
        Xw, ok = unproject(kl, depth, Tcw, K)
        Ow = -(Tcw[:3, :3].T.astype(f32) @ Tcw[:3, 3].astype(f32))
        cand = np.nonzero(ok & (kp_mp < 0))[0]
        cand = cand[np.argsort(depth[cand], kind="stable")]
        th_depth = f32(35.0) * f32(K[5])
        take = [i for n_, i in enumerate(cand) if n_ < 100 or depth[i] <= th_depth]
        for i in take:
            p = mp.n
            mp.pos[p], mp.desc[p] = Xw[i], dl[i]
            d = Xw[i] - Ow
            dist = f32(np.sqrt(np.sum(d.astype(np.float64) ** 2)))
            mp.normal[p] = d / dist                                         # MapPoint::UpdateNormalAndDepth (MapPoint.cc:342-378)
            mp.max_d[p] = dist * be.scale[kl["octave"][i]]
            mp.min_d[p] = mp.max_d[p] / be.scale[-1]
            mp.first_kf[p] = kid
            kp_mp[i] = p
            mp.n += 1
        np.add.at(mp.n_obs, kp_mp[kp_mp >= 0], 1)
        word, node, wt = be.bow(dl)
        self._rec("bow", dl.copy(), (word, node, wt))
        from orbx.vocabulary import bow_maps, feature_vector_csr
        _, fv = bow_maps(word, node, wt)
        ids, start, feat = feature_vector_csr(fv)
        kf = dict(kl=kl, dl=dl, ur=ur, depth=depth, Tcw=Tcw.copy(), mp=kp_mp.copy(), node_id=ids, node_start=start, node_feat=feat)
        mp.kfs.append(kf)
        self.local_pts = None
        self.stats["keyframes"] += 1
        # ---- LocalMapping::CreateNewMapPoints: SearchForTriangulation against up to 10 earlier keyframes ----
        A = dict(keys_un=kl, desc=dl, u_right=ur, has_mp=(kp_mp >= 0).astype(np.uint8), node_id=ids, node_start=start, node_feat=feat)
        for k2 in range(max(0, kid - 10), kid):
            o = mp.kfs[k2]
            base = np.linalg.norm((-(o["Tcw"][:3, :3].T @ o["Tcw"][:3, 3])) - Ow)
            if base < K[5]:                                                 # baseline shorter than the stereo baseline (LocalMapping.cc:240-244)
                continue
            B = dict(keys_un=o["kl"], desc=o["dl"], u_right=o["ur"], has_mp=(o["mp"] >= 0).astype(np.uint8), node_id=o["node_id"],
                     node_start=o["node_start"], node_feat=o["node_feat"])
            F12 = fundamental(Tcw, o["Tcw"], K)
            epi = epipole_in_2(Tcw, o["Tcw"], K)
            n, pairs = be.triangulation(A, B, F12, epi, self.sigma2, be.scale)
            self._rec("triangulation", (A, B, F12, epi), (n, pairs))
            self.stats["tri_pairs"] += int(n)
        # ---- Optimizer::LocalBundleAdjustment over the last 10 keyframes ----
        if len(mp.kfs) >= 3:
            ids_local = list(range(max(0, len(mp.kfs) - 10), len(mp.kfs)))[::-1]     # the new keyframe first, like lLocalKeyFrames
            prob, kfs, pts, e_ref = lba_problem(mp, ids_local, K, be.inv_sigma2)
            if len(prob["e_kf"]) > 50 and (prob["kf_fixed"] == 0).any():
                r = be.lba(prob)
                self._rec("lba", prob, r)
                self.stats["lba_windows"] += 1
                self.stats["lba_trials"] += int(r["trials"])
                for i, k in enumerate(kfs):
                    if not prob["kf_fixed"][i]:
                        mp.kfs[k]["Tcw"] = replay.pose_matrix(r["kf"][i])
                mp.pos[pts] = r["pts"].astype(f32)
                for e in np.nonzero(r["erase"])[0]:                          # Optimizer.cc:745-756
                    k, i = e_ref[e]
                    p = mp.kfs[k]["mp"][i]
                    if p >= 0:
                        mp.kfs[k]["mp"][i] = -1
                        mp.n_obs[p] -= 1
                self.local_pts = None

    def summary(self):
        s = self.stats
        return dict(frames=s["frames"], keyframes=s["keyframes"], map_points=int(self.map.n), a12_matches_per_frame=s["a12"] / max(s["frames"] - 1, 1),
                    a11_matches_per_frame=s["a11"] / max(s["frames"] - 1, 1), triangulation_pairs_per_keyframe=s["tri_pairs"] / max(s["keyframes"], 1),
                    lba_windows=s["lba_windows"], lba_trials_per_window=s["lba_trials"] / max(s["lba_windows"], 1),
                    inliers_per_frame=float(np.mean(s["inliers"][1:])) if len(s["inliers"]) > 1 else 0.0,
                    mean_position_error_m=float(np.mean(s["err"])), max_position_error_m=float(np.max(s["err"])))

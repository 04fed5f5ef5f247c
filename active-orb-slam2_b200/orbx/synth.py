"""Seeded synthetic inputs for the parity tests and the bench.

Pure numpy (no cv2), so the same frames are produced in this container and on the GPU box.
"""
import numpy as np


def _smooth3(img):
    """cheap separable [1 2 1]/4 smoothing in integer arithmetic (keeps generators cv2-free)"""
    a = img.astype(np.int32)
    p = np.pad(a, 1, mode="reflect")
    h = p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]
    v = h[:-2] + 2 * h[1:-1] + h[2:]
    return ((v + 8) >> 4).astype(np.uint8)


def g_rect(seed, w=640, h=480, nrect=900, noise=3.0):
    """G-rect: background 110 + filled axis-aligned rectangles of random grey + N(0, noise^2)."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 110, np.float32)
    cx = rng.uniform(0, w, nrect)
    cy = rng.uniform(0, h, nrect)
    side = rng.uniform(4, 40, nrect)
    asp = rng.uniform(0.4, 1.6, nrect)
    grey = rng.uniform(0, 255, nrect)
    for i in range(nrect):
        hw, hh = 0.5 * side[i] * asp[i], 0.5 * side[i]
        x0, x1 = int(max(cx[i] - hw, 0)), int(min(cx[i] + hw, w))
        y0, y1 = int(max(cy[i] - hh, 0)), int(min(cy[i] + hh, h))
        img[y0:y1, x0:x1] = grey[i]
    img += rng.normal(0, noise, (h, w)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def g_noise(seed, w=640, h=480):
    """G-noise: uniform noise, lightly smoothed (dense corners everywhere)."""
    rng = np.random.default_rng(seed)
    return _smooth3(rng.integers(0, 256, (h, w), dtype=np.uint8))


def g_sparse(seed, w=640, h=480):
    """G-sparse: few rectangles, low noise -> empty cells and the minThFAST retry path."""
    return g_rect(seed, w, h, nrect=40, noise=1.0)


def g_flat(w=640, h=480, value=128):
    """G-flat: constant image -> zero keypoints (reference ORBextractor.cc:1064-1065)."""
    return np.full((h, w), value, np.uint8)


GENERATORS = {"rect": g_rect, "noise": g_noise, "sparse": g_sparse}


def frame(kind, seed, w=640, h=480):
    if kind == "flat":
        return g_flat(w, h)
    return GENERATORS[kind](seed, w, h)


# ---- stereo / sequence world: textured fronto-parallel planes seen by a moving rig ----
class PlaneWorld:
    """Three textured fronto-parallel planes; renders rectified left/right views for a camera pose.

    The renderer is a nearest-texel lookup with per-pixel plane selection by image row band, which keeps
    it exact integer work (deterministic across machines) while giving real disparity structure.
    """

    def __init__(self, seed, w=640, h=480, fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, bf=40.0):
        self.w, self.h, self.fx, self.fy, self.cx, self.cy, self.bf = w, h, fx, fy, cx, cy, bf
        self.depths = (2.0, 4.0, 8.0)
        self.tex = [g_rect(seed * 16 + i, 2048, 2048, nrect=9000) for i in range(3)]
        self.texscale = 220.0  # texels per metre

    def plane_of_row(self, v):
        b = (v * 3) // self.h
        return np.clip(b, 0, 2)

    def render(self, tx, ty, yaw, right=False):
        """view from camera at (tx, ty, 0) with yaw (rad) about y; right camera is offset by baseline."""
        h, w = self.h, self.w
        vs, us = np.mgrid[0:h, 0:w]
        pl = self.plane_of_row(vs)
        z = np.choose(pl, self.depths)
        base = self.bf / self.fx
        xc = (us - self.cx) / self.fx * z + (base if right else 0.0)
        yc = (vs - self.cy) / self.fy * z
        cs, sn = np.cos(yaw), np.sin(yaw)
        xw = cs * xc + sn * z + tx
        yw = yc + ty
        out = np.zeros((h, w), np.uint8)
        for i in range(3):
            m = pl == i
            tu = np.mod(np.rint(xw[m] * self.texscale + 1024).astype(np.int64), 2048)
            tv = np.mod(np.rint(yw[m] * self.texscale + 1024).astype(np.int64), 2048)
            out[m] = self.tex[i][tv, tu]
        return out


_WORLDS = {}


def stereo_world(seed, w=640, h=480, **kw):
    """cached PlaneWorld (building the three textures takes half a second)"""
    key = (seed, w, h, tuple(sorted(kw.items())))
    if key not in _WORLDS:
        _WORLDS[key] = PlaneWorld(seed, w, h, **kw)
    return _WORLDS[key]


# ---- matcher scenes: a current frame, and points that project near its keypoints ----------------
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
TUM1_K = (517.306408, 516.469215, 318.643040, 255.313989, 40.0, 40.0 / 517.306408)   # fx fy cx cy bf b


def scale_factors(nlevels=8, sf=1.2):
    s = np.ones(nlevels, np.float32)
    for i in range(1, nlevels):
        s[i] = np.float32(np.float64(s[i - 1]) * np.float64(np.float32(sf)))
    return s


def random_frame(rng, n, w=640, h=480, nlevels=8, stereo_frac=0.5, claimed_frac=0.05):
    """a Frame view with n keypoints laid out like extractor output (integer level coordinates times the level scale)"""
    sf = scale_factors(nlevels)
    oct_ = rng.choice(nlevels, n, p=np.array([1.2 ** -l for l in range(nlevels)]) / sum(1.2 ** -l for l in range(nlevels))).astype(np.int32)
    k = np.zeros(n, KP_DTYPE)
    lx = rng.integers(19, (w / sf[oct_]).astype(np.int64) - 19)
    ly = rng.integers(19, (h / sf[oct_]).astype(np.int64) - 19)
    k["x"] = lx.astype(np.float32) * sf[oct_]
    k["y"] = ly.astype(np.float32) * sf[oct_]
    k["octave"] = oct_
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["size"] = (31 * sf[oct_]).astype(np.int32)
    k["response"] = rng.integers(7, 200, n)
    k["class_id"] = -1
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ur = np.where(rng.random(n) < stereo_frac, k["x"] - rng.uniform(2, 40, n).astype(np.float32), np.float32(-1)).astype(np.float32)
    claimed = (rng.random(n) < claimed_frac).astype(np.uint8)
    return dict(keys_un=k, desc=desc, u_right=ur, claimed=claimed, bounds=(0.0, 0.0, float(w), float(h)), K=TUM1_K, scale_factors=sf)


def flip_bits(rng, desc, nbits):
    out = desc.copy()
    for r in range(len(out)):
        for b in rng.choice(256, int(nbits[r]), replace=False):
            out[r, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def last_frame_points(rng, cur, n_pts, dup_frac=0.15, angle_jitter=8.0):
    """points of a 'last frame' that project near keypoints of `cur` under pose (Rcw, tcw); some share a target"""
    LP = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("valid", "u1"),
                   ("blocks", "u1"), ("pad", "u1", (2,))])
    fx, fy, cx, cy, bf, b = cur["K"]
    n = len(cur["keys_un"])
    yaw = rng.uniform(-0.05, 0.05)
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = rng.uniform(-0.1, 0.1, 3)
    tgt = rng.integers(0, n, n_pts)
    ndup = int(dup_frac * n_pts)
    if ndup and n_pts > ndup:
        tgt[-ndup:] = tgt[:ndup]                       # several points aim at the same keypoint -> claim conflicts
    k = cur["keys_un"][tgt]
    u = k["x"].astype(np.float64) + rng.normal(0, 2.0, n_pts)
    v = k["y"].astype(np.float64) + rng.normal(0, 2.0, n_pts)
    z = rng.uniform(1.0, 8.0, n_pts)
    behind = rng.random(n_pts) < 0.03
    z[behind] *= -1
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    pw = (pc - t) @ R                                   # R^T (pc - t)
    pts = np.zeros(n_pts, LP)
    pts["x"], pts["y"], pts["z"] = pw[:, 0], pw[:, 1], pw[:, 2]
    pts["octave"] = np.clip(k["octave"] + rng.integers(-1, 2, n_pts), 0, len(cur["scale_factors"]) - 1)
    pts["angle"] = np.mod(k["angle"] + rng.normal(0, angle_jitter, n_pts) + np.where(rng.random(n_pts) < 0.1, 90, 0), 360).astype(np.float32)
    pts["valid"] = rng.random(n_pts) > 0.05
    pts["blocks"] = rng.random(n_pts) > 0.2
    desc = flip_bits(rng, cur["desc"][tgt], rng.integers(0, 90, n_pts))
    return pts, desc, R.astype(np.float32), t.astype(np.float32)


def track_points(rng, cur, n_pts, dup_frac=0.15):
    """map points as Frame::isInFrustum leaves them (mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos)"""
    TP = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                   ("in_view", "u1"), ("blocks", "u1"), ("pad", "u1", (2,))])
    n = len(cur["keys_un"])
    tgt = rng.integers(0, n, n_pts)
    ndup = int(dup_frac * n_pts)
    if ndup and n_pts > ndup:
        tgt[-ndup:] = tgt[:ndup]
    k = cur["keys_un"][tgt]
    pts = np.zeros(n_pts, TP)
    pts["proj_x"] = k["x"] + rng.normal(0, 1.5, n_pts)
    pts["proj_y"] = k["y"] + rng.normal(0, 1.5, n_pts)
    ur = cur["u_right"][tgt]
    pts["proj_xr"] = np.where(ur > 0, ur + rng.normal(0, 2.0, n_pts), pts["proj_x"] - 10)
    pts["view_cos"] = rng.uniform(0.99, 1.0, n_pts)
    pts["level"] = np.clip(k["octave"] + rng.integers(0, 2, n_pts), 0, len(cur["scale_factors"]) - 1)
    pts["in_view"] = rng.random(n_pts) > 0.1
    pts["blocks"] = rng.random(n_pts) > 0.1
    desc = flip_bits(rng, cur["desc"][tgt], rng.integers(0, 110, n_pts))
    return pts, desc


# ---- local bundle adjustment problems -----------------------------------------------------------
def _quat_from_R(R):
    """(x, y, z, w), w >= 0"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R))); j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s; q[3] = (R[k, j] - R[j, k]) / s; q[j] = (R[j, i] + R[i, j]) / s; q[k] = (R[k, i] + R[i, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    return {0: np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), 1: np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
            2: np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]


def lba_problem(seed, n_kf=20, n_pts=3000, obs_per_pt=4, stereo=False, n_fixed=0, outlier_frac=0.05, w=640, h=480):
    """P keyframes looking down +z at L points in a 6 x 4 x 6 m box 2-8 m ahead; every point is observed by
    `obs_per_pt` keyframes; pixel noise N(0, sigma_octave^2), 5 % outliers (+20 px); poses perturbed by 1 cm / 0.5 deg,
    points by 2 cm.  The first n_fixed keyframes are fixed.  Returns a dict of numpy arrays (orbx_lba_problem fields)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf, _ = TUM1_K
    Rs, ts = [], []
    for k in range(n_kf):
        R = _rot(1, rng.uniform(-0.15, 0.15)) @ _rot(0, rng.uniform(-0.05, 0.05))
        c = np.array([rng.uniform(-1.0, 1.0), rng.uniform(-0.3, 0.3), rng.uniform(-0.5, 0.5)])   # camera centre
        Rs.append(R.T); ts.append(-R.T @ c)                                                     # Tcw
    X = np.stack([rng.uniform(-3, 3, n_pts), rng.uniform(-2, 2, n_pts), rng.uniform(2, 8, n_pts)], 1)
    e_kf, e_pt, e_obs, e_is2, e_st = [], [], [], [], []
    sf = scale_factors(8)
    for l in range(n_pts):
        seen = 0
        for k in rng.permutation(n_kf):
            if seen == obs_per_pt:
                break
            Xc = Rs[k] @ X[l] + ts[k]
            if Xc[2] < 0.5:
                continue
            u, v = fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy
            if not (0 <= u < w and 0 <= v < h):
                continue
            octv = int(rng.integers(0, 8))
            sig = float(sf[octv])
            du, dv = rng.normal(0, sig, 2)
            if rng.random() < outlier_frac:
                du += 20.0
            ur = u + du - bf / Xc[2] + rng.normal(0, sig)
            e_kf.append(k); e_pt.append(l); e_obs.append((np.float32(u + du), np.float32(v + dv), np.float32(ur)))
            e_is2.append(np.float32(1.0) / (sf[octv] * sf[octv])); e_st.append(1 if stereo else 0)
            seen += 1
    kf_pose = np.zeros((n_kf, 7))
    for k in range(n_kf):
        Rn = _rot(int(rng.integers(0, 3)), np.deg2rad(rng.normal(0, 0.5))) @ Rs[k] if k >= n_fixed else Rs[k]
        tn = ts[k] + (rng.normal(0, 0.01, 3) if k >= n_fixed else 0)
        Rf = Rn.astype(np.float32).astype(np.float64)          # poses enter as float cv::Mat (Converter::toSE3Quat)
        kf_pose[k, :4] = _quat_from_R(Rf)
        kf_pose[k, 4:] = tn.astype(np.float32)
    pts = (X + rng.normal(0, 0.02, X.shape)).astype(np.float32).astype(np.float64)
    fixed = np.zeros(n_kf, np.uint8); fixed[:n_fixed] = 1
    return dict(kf_pose=kf_pose, kf_fixed=fixed, pts=pts, e_kf=np.array(e_kf, np.int32), e_pt=np.array(e_pt, np.int32),
                e_obs=np.array(e_obs, np.float64).reshape(-1, 3), e_inv_sigma2=np.array(e_is2, np.float32),
                e_stereo=np.array(e_st, np.uint8), K=(float(np.float32(fx)), float(np.float32(fy)), float(np.float32(cx)),
                                                      float(np.float32(cy)), float(np.float32(bf))))


# ---- vocabulary-node matcher scenes: two keyframes seeing the same 3-D points ---------------------------------------
def bow_pair(seed, n_a=1000, n_b=1000, n_common=600, n_nodes=90, nlevels=8):
    """Two feature sets with `n_common` true correspondences (similar descriptors, same vocabulary node, consistent
    epipolar geometry), the rest unrelated.  Returns (A, B, F12, (ex, ey), sigma2, scale) with A/B dicts holding
    keys_un, desc, u_right, has_mp (the feature already has a map point), node_id, node_start, node_feat."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf, _ = TUM1_K
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    sf = scale_factors(nlevels)
    # camera 1 at the origin, camera 2 displaced and slightly rotated
    R21 = _rot(1, rng.uniform(-0.1, 0.1)) @ _rot(0, rng.uniform(-0.03, 0.03))
    t21 = np.array([rng.uniform(0.2, 0.5), rng.uniform(-0.05, 0.05), rng.uniform(-0.1, 0.1)])
    X = np.stack([rng.uniform(-2, 2, n_common), rng.uniform(-1.5, 1.5, n_common), rng.uniform(2, 8, n_common)], 1)
    x1 = (K @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
    X2 = (R21 @ X.T).T + t21
    x2 = (K @ X2.T).T; x2 = x2[:, :2] / x2[:, 2:]

    def make_set(n, xy_common, desc_common, node_common, oct_common):
        k = np.zeros(n, KP_DTYPE)
        k["x"] = rng.uniform(20, 620, n); k["y"] = rng.uniform(20, 460, n)
        k["octave"] = rng.integers(0, nlevels, n)
        k["angle"] = rng.uniform(0, 360, n)
        desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        node = rng.integers(0, n_nodes, n)
        slots = rng.permutation(n)[:len(xy_common)]
        k["x"][slots] = xy_common[:, 0] + rng.normal(0, 0.4, len(slots))
        k["y"][slots] = xy_common[:, 1] + rng.normal(0, 0.4, len(slots))
        k["octave"][slots] = oct_common
        desc[slots] = desc_common
        node[slots] = node_common
        ur = np.where(rng.random(n) < 0.4, k["x"] - rng.uniform(2, 30, n), -1).astype(np.float32)
        has_mp = (rng.random(n) < 0.5).astype(np.uint8)
        # FeatureVector: ascending node ids (sparse ids like DBoW2 node numbers), lists in ascending feature order
        ids = np.unique(node)
        start = np.zeros(len(ids) + 1, np.int32)
        feat = []
        for j, nid in enumerate(ids):
            f = np.nonzero(node == nid)[0]
            feat.append(f); start[j + 1] = start[j] + len(f)
        return dict(keys_un=k, desc=desc, u_right=ur, has_mp=has_mp, node_id=(ids * 7 + 3).astype(np.uint32), node_start=start,
                    node_feat=np.concatenate(feat).astype(np.int32) if feat else np.zeros(0, np.int32)), slots

    dcom = rng.integers(0, 256, (n_common, 32), dtype=np.uint8)
    ncom = rng.integers(0, n_nodes, n_common)
    ocom = rng.integers(0, nlevels, n_common)
    A, sa = make_set(n_a, x1, dcom, ncom, ocom)
    B, sb = make_set(n_b, x2, flip_bits(rng, dcom, rng.integers(0, 70, n_common)), ncom, np.clip(ocom + rng.integers(-1, 2, n_common), 0, nlevels - 1))
    A["angle_ref"] = None
    B["keys_un"]["angle"][sb] = np.mod(A["keys_un"]["angle"][sa] + rng.normal(0, 6, n_common) + np.where(rng.random(n_common) < 0.08, 120, 0), 360)
    tx = np.array([[0, -t21[2], t21[1]], [t21[2], 0, -t21[0]], [-t21[1], t21[0], 0]])
    F12 = (np.linalg.inv(K).T @ (tx @ R21) @ np.linalg.inv(K)).T          # x1^T F12 x2 = 0  (ORB-SLAM2's convention)
    F12 = (F12 / np.abs(F12).max()).astype(np.float32)
    C2 = t21                                                              # camera-1 centre in camera-2 coordinates
    ex, ey = np.float32(fx * C2[0] / C2[2] + cx), np.float32(fy * C2[1] / C2[2] + cy)
    return A, B, F12, (ex, ey), (sf * sf).astype(np.float32), sf


def window_points(rng, cur, n_pts, th=3.0, dup_frac=0.2):
    """points already projected into a keyframe (u, v, ur, radius, level window) near its keypoints, plus descriptors"""
    WP = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"),
                   ("valid", "u1"), ("pad", "u1", (3,))])
    n = len(cur["keys_un"])
    tgt = rng.integers(0, n, n_pts)
    ndup = int(dup_frac * n_pts)
    if ndup and n_pts > ndup:
        tgt[-ndup:] = tgt[:ndup]
    k = cur["keys_un"][tgt]
    sf = cur["scale_factors"]
    lvl = np.clip(k["octave"] + rng.integers(0, 2, n_pts), 0, len(sf) - 1)
    pts = np.zeros(n_pts, WP)
    pts["u"] = k["x"] + rng.normal(0, 1.2, n_pts)
    pts["v"] = k["y"] + rng.normal(0, 1.2, n_pts)
    ur = cur["u_right"][tgt]
    pts["ur"] = np.where(ur >= 0, ur + rng.normal(0, 1.5, n_pts), pts["u"] - 8)
    pts["radius"] = np.float32(th) * sf[lvl]
    pts["min_level"], pts["max_level"] = lvl - 1, lvl
    pts["valid"] = rng.random(n_pts) > 0.1
    desc = flip_bits(rng, cur["desc"][tgt], rng.integers(0, 80, n_pts))
    return pts, desc


def pose_problem(seed, n=400, stereo_frac=0.6, outlier_frac=0.15, w=640, h=480, rot_deg=1.0, trans=0.03):
    """One frame's PoseOptimization input: n map points 1-9 m ahead seen by a camera whose initial pose is
    off by ~rot_deg / trans metres; pixel noise N(0, sigma_octave^2); outlier_frac of the observations are off by 15-40 px.
    Map points and the pose enter as float (cv::Mat CV_32F), observations as float keypoints, like the reference."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf, _ = TUM1_K
    R = _rot(1, rng.uniform(-0.3, 0.3)) @ _rot(0, rng.uniform(-0.1, 0.1))
    t = rng.uniform(-0.5, 0.5, 3)
    sf = scale_factors(8)
    Xw, obs, is2 = [], [], []
    while len(Xw) < n:
        Xc = np.array([rng.uniform(-4, 4), rng.uniform(-3, 3), rng.uniform(1, 9)])
        u, v = fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy
        if not (0 <= u < w and 0 <= v < h):
            continue
        octv = int(rng.integers(0, 8))
        sig = float(sf[octv])
        du, dv, dr = rng.normal(0, sig, 3)
        if rng.random() < outlier_frac:
            a = rng.uniform(0, 2 * np.pi)
            m = rng.uniform(15, 40)
            du += m * np.cos(a); dv += m * np.sin(a)
        ur = u + dr - bf / Xc[2] if rng.random() < stereo_frac else -1.0
        Xw.append((R.T @ (Xc - t)).astype(np.float32))
        obs.append((np.float32(u + du), np.float32(v + dv), np.float32(ur)))
        is2.append(np.float32(1.0) / (sf[octv] * sf[octv]))
    Rn = _rot(int(rng.integers(0, 3)), np.deg2rad(rng.normal(0, rot_deg))) @ R
    tn = t + rng.normal(0, trans, 3)
    pose = np.zeros(7)
    pose[:4] = _quat_from_R(Rn.astype(np.float32).astype(np.float64))
    pose[4:] = tn.astype(np.float32)
    return dict(Xw=np.asarray(Xw, np.float64), obs=np.asarray(obs, np.float64), inv_sigma2=np.asarray(is2, np.float32), pose=pose,
                K=(np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy), np.float32(bf)))


def random_vocabulary(seed, k=10, L=4, prune=0.1, dup=0.1, stop=0.02, shuffle=True):
    """A vocabulary tree as a DBoW2 file would describe it (parent, descriptor, weight, is_leaf per node 1..n in file order):
    up to k children per node, some branches end early (prune), some siblings share a descriptor (dup: exercises `d < best_d`
    keeping the first child), some words are stopped (weight 0).  shuffle permutes the node ids so that siblings are not
    consecutive.  Returns the arguments of orbx.vocabulary.tree_from_parents."""
    rng = np.random.default_rng(seed)
    parent, depth = [], []
    frontier = [0]
    for lvl in range(1, L + 1):
        nxt = []
        for p in frontier:
            if p != 0 and rng.random() < prune:
                continue                                           # p stays a leaf above the last level
            for _ in range(int(rng.integers(2, k + 1))):
                parent.append(p); depth.append(lvl)
                nxt.append(len(parent))
        frontier = nxt
    n = len(parent)
    parent = np.array(parent, np.int64)
    desc = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    for i in range(1, n):
        if parent[i] == parent[i - 1] and rng.random() < dup:
            desc[i] = desc[i - 1]
    has_child = np.zeros(n + 1, bool)
    has_child[parent] = True
    is_leaf = ~has_child[1:]
    weight = np.where(is_leaf, rng.uniform(0.1, 9.0, n), 0.0).astype(np.float32)
    weight[is_leaf & (rng.random(n) < stop)] = 0.0
    if shuffle:
        perm = rng.permutation(n)                                  # new id of old node i+1 is perm[i]+1
        new_parent = np.zeros(n, np.int64)
        new_parent[perm] = np.where(parent == 0, 0, perm[np.maximum(parent - 1, 0)] + 1)
        nd, nw, nl = np.zeros_like(desc), np.zeros_like(weight), np.zeros_like(is_leaf)
        nd[perm], nw[perm], nl[perm] = desc, weight, is_leaf
        parent, desc, weight, is_leaf = new_parent, nd, nw, nl
    return parent.astype(np.int32), desc, weight, is_leaf.astype(np.uint8), k, L


def descriptors_near_words(rng, tree, n, flip=20):
    """n descriptors: leaf descriptors of the tree with `flip` random bits flipped (so descents are not uniform noise)"""
    leaves = np.nonzero(tree["word_id"] >= 0)[0]
    pick = rng.choice(leaves, n)
    return flip_bits(rng, tree["desc"][pick].copy(), np.full(n, flip))


def frustum_scene(seed, n=4000, w=640, h=480, nlevels=8, sf=1.2):
    """A frame pose and n map points for Frame::isInFrustum: points scattered around the camera (in front, behind, outside the
    image, too near / too far for their scale-invariance range, seen from behind), as float records.  Returns
    (frame dict of orbx_frustum_frame fields, structured point array fields as a dict of arrays)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, bf, _ = TUM1_K
    R = (_rot(1, rng.uniform(-0.5, 0.5)) @ _rot(0, rng.uniform(-0.2, 0.2))).astype(np.float32)
    t = rng.uniform(-1, 1, 3).astype(np.float32)
    Ow = (-R.T.astype(np.float64) @ t.astype(np.float64)).astype(np.float32)
    frame = dict(Rcw=R.reshape(9), tcw=t, Ow=Ow, fx=fx, fy=fy, cx=cx, cy=cy, bf=bf, min_x=0.0, max_x=float(w), min_y=0.0, max_y=float(h),
                 log_scale_factor=np.float32(np.log(np.float32(sf))), n_levels=nlevels, viewing_cos_limit=0.5)
    Pc = np.stack([rng.uniform(-6, 6, n), rng.uniform(-4, 4, n), rng.uniform(-2, 12, n)], 1)
    P = ((Pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)        # R^T (Pc - t)
    d = np.linalg.norm(P.astype(np.float64) - Ow, axis=1)
    ref_d = d * rng.uniform(0.4, 2.5, n)                                             # distance at which the point was created
    lvl = rng.integers(0, nlevels, n)
    max_d = (ref_d * np.float32(sf) ** lvl).astype(np.float32)                      # MapPoint::UpdateNormalAndDepth, MapPoint.cc:413-415
    min_d = (max_d / np.float32(sf) ** (nlevels - 1)).astype(np.float32)
    view = (Ow - P.astype(np.float64)); view /= np.maximum(np.linalg.norm(view, axis=1, keepdims=True), 1e-9)
    nrm = -view + rng.normal(0, 0.6, (n, 3))                                         # normals roughly towards ... away from the camera
    flip = rng.random(n) < 0.15
    nrm[flip] = -nrm[flip]
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-9)
    pts = dict(x=P[:, 0], y=P[:, 1], z=P[:, 2], nx=nrm[:, 0].astype(np.float32), ny=nrm[:, 1].astype(np.float32), nz=nrm[:, 2].astype(np.float32),
               min_distance=min_d, max_distance=max_d, skip=(rng.random(n) < 0.05).astype(np.uint8), blocks=(rng.random(n) < 0.8).astype(np.uint8))
    return frame, pts

"""Host-side mirror of ORB_SLAM2::ORBmatcher (reference include/ORBmatcher.h:37-102) over the orbx C ABI.

A `frame` is a dict with the Frame members the matchers read: keys_un (cv::KeyPoint records), desc (N x 32),
u_right or None, claimed or None (mvpMapPoints[i] && Observations()>0), bounds (mnMinX, mnMinY, mnMaxX, mnMaxY),
K (fx, fy, cx, cy, mbf, mb), scale_factors (mvScaleFactors).  `match` plays the role of Frame::mvpMapPoints as
indices into the point array (-1 = NULL)."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib

TRACK_POINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                              ("in_view", "u1"), ("blocks", "u1"), ("pad", "u1", (2,))])
LAST_POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("valid", "u1"),
                             ("blocks", "u1"), ("pad", "u1", (2,))])


class FrameView(C.Structure):
    """orbx_frame_view (include/orbx.h)"""
    _fields_ = [("n", C.c_int32), ("n_dev", C.c_void_p), ("keys_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p),
                ("claimed", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("b", C.c_float),
                ("scale_factors", C.c_void_p), ("nlevels", C.c_int32)]


class FrameMatchJob(C.Structure):
    """orbx_frame_match_job (include/orbx.h)"""
    _fields_ = [("cur", FrameView), ("n_last", C.c_int32), ("pts", C.c_void_p), ("last_desc", C.c_void_p),
                ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("forward", C.c_int32), ("backward", C.c_int32),
                ("th", C.c_float), ("check_ori", C.c_int32), ("match", C.c_void_p), ("nmatches", C.c_void_p),
                ("max_dist", C.c_int32), ("variant", C.c_int32)]


def fill_view(f, bounds, K, nlevels):
    mnx, mny, mxx, mxy = (np.float32(v) for v in bounds)
    f.min_x, f.min_y, f.max_x, f.max_y = mnx, mny, mxx, mxy
    f.grid_w_inv = np.float32(64) / (mxx - mnx)      # Frame.cc:127-128
    f.grid_h_inv = np.float32(48) / (mxy - mny)
    f.fx, f.fy, f.cx, f.cy, f.bf, f.b = (np.float32(v) for v in K)
    f.nlevels = nlevels


def _view(fr):
    keys = np.ascontiguousarray(fr["keys_un"], KP_DTYPE)
    desc = np.ascontiguousarray(fr["desc"], np.uint8)
    sf = np.ascontiguousarray(fr["scale_factors"], np.float32)
    keep = [keys, desc, sf]
    f = FrameView()
    f.n = len(keys)
    f.keys_un, f.desc, f.scale_factors = keys.ctypes.data, desc.ctypes.data, sf.ctypes.data
    if fr.get("u_right") is not None:
        ur = np.ascontiguousarray(fr["u_right"], np.float32); keep.append(ur); f.u_right = ur.ctypes.data
    if fr.get("claimed") is not None:
        cl = np.ascontiguousarray(fr["claimed"], np.uint8); keep.append(cl); f.claimed = cl.ctypes.data
    fill_view(f, fr["bounds"], fr["K"], len(sf))
    return f, keep


def motion_flags(Tcw_cur, Tcw_last, mb, mono):
    """bForward / bBackward of ORBmatcher.cc:1338-1351 in float32"""
    Tc, Tl = np.asarray(Tcw_cur, np.float32), np.asarray(Tcw_last, np.float32)
    twc = -(Tc[:3, :3].T @ Tc[:3, 3])
    tlc = Tl[:3, :3] @ twc + Tl[:3, 3]
    return bool(tlc[2] > np.float32(mb) and not mono), bool(-tlc[2] > np.float32(mb) and not mono)


class ORBmatcher:
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30      # ORBmatcher.cc:35-37

    def __init__(self, nnratio=0.6, checkOri=True, max_keypoints=4096, max_points=4096, max_jobs=1, device=0):
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.orbx_matcher_create(C.byref(self._h), max_keypoints, max_points, max_jobs, device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_matcher_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @staticmethod
    def DescriptorDistance(a, b):
        a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
        return lib().orbx_hamming256(a.ctypes.data, b.ctypes.data)

    def SearchByProjection(self, F, pts, pt_desc, th, match=None):
        """SearchByProjection(Frame &F, const vector<MapPoint*>&, th), ORBmatcher.cc:45 -> (nmatches, match)"""
        f, keep = _view(F)
        pts = np.ascontiguousarray(pts, TRACK_POINT_DTYPE)
        pd = np.ascontiguousarray(pt_desc, np.uint8)
        m = np.full(f.n, -1, np.int32) if match is None else np.ascontiguousarray(match, np.int32).copy()
        n = C.c_int32()
        check(self._L.orbx_match_projection_points_host(self._h, C.byref(f), len(pts), pts.ctypes.data, pd.ctypes.data, th,
                                                        self.mfNNratio, m.ctypes.data, C.byref(n)))
        return n.value, m

    def SearchByProjectionLast(self, Cur, last_pts, last_desc, Rcw, tcw, forward, backward, th, match=None):
        """SearchByProjection(Frame &Cur, const Frame &Last, th, bMono), ORBmatcher.cc:1328 -> (nmatches, match)"""
        f, keep = _view(Cur)
        pts = np.ascontiguousarray(last_pts, LAST_POINT_DTYPE)
        pd = np.ascontiguousarray(last_desc, np.uint8)
        R = (C.c_float * 9)(*np.asarray(Rcw, np.float32).reshape(9).tolist())
        t = (C.c_float * 3)(*np.asarray(tcw, np.float32).reshape(3).tolist())
        m = np.full(f.n, -1, np.int32) if match is None else np.ascontiguousarray(match, np.int32).copy()
        n = C.c_int32()
        check(self._L.orbx_match_projection_frame_host(self._h, C.byref(f), len(pts), pts.ctypes.data, pd.ctypes.data, R, t,
                                                       int(forward), int(backward), th, int(self.mbCheckOrientation),
                                                       m.ctypes.data, C.byref(n)))
        return n.value, m

    def SearchByProjectionKF(self, Cur, pts, pt_desc, Rcw, tcw, th, ORBdist, match=None):
        """SearchByProjection(Frame &Cur, KeyFrame*, sAlreadyFound, th, ORBdist), ORBmatcher.cc:1472 -> (nmatches, match);
        pts[i].valid / .octave carry the adapter's host-side gates and PredictScale (see include/orbx.h)"""
        f, keep = _view(Cur)
        pts = np.ascontiguousarray(pts, LAST_POINT_DTYPE)
        pd = np.ascontiguousarray(pt_desc, np.uint8)
        R = (C.c_float * 9)(*np.asarray(Rcw, np.float32).reshape(9).tolist())
        t = (C.c_float * 3)(*np.asarray(tcw, np.float32).reshape(3).tolist())
        m = np.full(f.n, -1, np.int32) if match is None else np.ascontiguousarray(match, np.int32).copy()
        n = C.c_int32()
        check(self._L.orbx_match_projection_keyframe_host(self._h, C.byref(f), len(pts), pts.ctypes.data, pd.ctypes.data, R, t, th,
                                                          int(ORBdist), int(self.mbCheckOrientation), m.ctypes.data, C.byref(n)))
        return n.value, m

    def search_frames_device(self, d_jobs, n_jobs, stream=0):
        """batched device-resident SearchByProjection(Cur, Last): d_jobs = device pointer to orbx_frame_match_job[n_jobs]"""
        check(self._L.orbx_match_projection_frame_device(self._h, d_jobs, n_jobs, stream))

    def last_sweeps(self, n_jobs=1):
        out = np.zeros(n_jobs, np.int32)
        check(self._L.orbx_matcher_last_sweeps(self._h, out.ctypes.data, n_jobs))
        return out


class BowSet(C.Structure):
    """orbx_bow_set (include/orbx.h)"""
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("valid", C.c_void_p),
                ("n_nodes", C.c_int32), ("node_id", C.c_void_p), ("node_start", C.c_void_p), ("node_feat", C.c_void_p)]


class BucketJob(C.Structure):
    """orbx_bucket_job (include/orbx.h)"""
    _fields_ = [("a", BowSet), ("b", BowSet), ("mode", C.c_int32), ("nnratio", C.c_float), ("check_ori", C.c_int32),
                ("only_stereo", C.c_int32), ("F12", C.c_float * 9), ("ex", C.c_float), ("ey", C.c_float),
                ("sigma2_b", C.c_void_p), ("scale_b", C.c_void_p), ("nlevels", C.c_int32)]


def _bow_set(S, d, valid, keep):
    arrs = dict(keys_un=np.ascontiguousarray(d["keys_un"], KP_DTYPE), desc=np.ascontiguousarray(d["desc"], np.uint8),
                u_right=np.ascontiguousarray(d["u_right"], np.float32), valid=np.ascontiguousarray(valid, np.uint8),
                node_id=np.ascontiguousarray(d["node_id"], np.uint32), node_start=np.ascontiguousarray(d["node_start"], np.int32),
                node_feat=np.ascontiguousarray(d["node_feat"], np.int32))
    keep.append(arrs)
    S.n, S.n_nodes = len(arrs["keys_un"]), len(arrs["node_id"])
    for k, v in arrs.items():
        setattr(S, k, v.ctypes.data)


def _buckets(self, mode, A, B, va, vb, only_stereo=False, F12=None, epipole=(0, 0), sigma2=None, scale=None):
    J, keep = BucketJob(), []
    _bow_set(J.a, A, va, keep); _bow_set(J.b, B, vb, keep)
    J.mode, J.nnratio, J.check_ori, J.only_stereo = mode, self.mfNNratio, int(self.mbCheckOrientation), int(only_stereo)
    if F12 is not None:
        J.F12[:] = np.asarray(F12, np.float32).reshape(9).tolist()
    J.ex, J.ey = float(epipole[0]), float(epipole[1])
    s2 = np.ascontiguousarray(sigma2 if sigma2 is not None else np.ones(8), np.float32)
    sc = np.ascontiguousarray(scale if scale is not None else np.ones(8), np.float32)
    J.sigma2_b, J.scale_b, J.nlevels = s2.ctypes.data, sc.ctypes.data, len(s2)
    m = np.zeros(max(J.a.n, 1), np.int32)
    n = C.c_int32()
    check(self._L.orbx_match_buckets_host(self._h, C.byref(J), m.ctypes.data, C.byref(n)))
    return n.value, m[:J.a.n]


def _search_by_bow_frame(self, KF, F):
    """SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), ORBmatcher.cc:159 -> (nmatches, matches indexed by the frame's
    features = keyframe feature index or -1)"""
    n, ma = _buckets(self, 0, KF, F, KF["has_mp"], F["has_mp"])
    out = np.full(len(F["keys_un"]), -1, np.int32)
    sel = ma >= 0
    out[ma[sel]] = np.nonzero(sel)[0]
    return n, out


def _search_by_bow_kf(self, KF1, KF2):
    """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), ORBmatcher.cc:522 -> (nmatches, match12)"""
    return _buckets(self, 1, KF1, KF2, KF1["has_mp"], KF2["has_mp"])


def _search_for_triangulation(self, KF1, KF2, F12, epipole, sigma2_2, scale_2, bOnlyStereo=False):
    """SearchForTriangulation, ORBmatcher.cc:657 -> (nmatches, vMatchedPairs as an (n, 2) array sorted by idx1)"""
    n, ma = _buckets(self, 2, KF1, KF2, 1 - KF1["has_mp"], 1 - KF2["has_mp"], bOnlyStereo, F12, epipole, sigma2_2, scale_2)
    idx1 = np.nonzero(ma >= 0)[0]
    return n, np.stack([idx1, ma[idx1]], 1)


ORBmatcher.SearchByBoW = _search_by_bow_frame
ORBmatcher.SearchByBoWKF = _search_by_bow_kf
ORBmatcher.SearchForTriangulation = _search_for_triangulation


WINDOW_POINT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"),
                               ("valid", "u1"), ("pad", "u1", (3,))])


def _match_window(self, KF, pts, pt_desc, flags, inv_sigma2, max_dist):
    """window + Hamming core of SearchByProjection(KF, Scw, ...) [flags=2], Fuse(KF, pts, th) [1], Fuse(KF, Scw, ...) [0] and
    SearchBySim3 [0]: -> (accepted, best_idx[n], best_dist[n]); see include/orbx.h for what the adapter does on the host"""
    f, keep = _view(KF)
    pts = np.ascontiguousarray(pts, WINDOW_POINT_DTYPE)
    pd = np.ascontiguousarray(pt_desc, np.uint8)
    s2 = np.ascontiguousarray(inv_sigma2, np.float32)
    bi, bd = np.full(max(len(pts), 1), -1, np.int32), np.full(max(len(pts), 1), 256, np.int32)
    n = C.c_int32()
    check(self._L.orbx_match_window_host(self._h, C.byref(f), len(pts), pts.ctypes.data, pd.ctypes.data, flags, s2.ctypes.data, max_dist,
                                         bi.ctypes.data, bd.ctypes.data, C.byref(n)))
    return n.value, bi[:len(pts)], bd[:len(pts)]


ORBmatcher.MatchWindow = _match_window


def _search_for_initialization(self, F1, F2, vbPrevMatched, windowSize=100):
    """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:405 -> (nmatches, vnMatches12,
    updated vbPrevMatched)"""
    f1, k1 = _view(F1)
    f2, k2 = _view(F2)
    prev = np.ascontiguousarray(vbPrevMatched, np.float32).reshape(-1, 2).copy()
    m = np.full(max(f1.n, 1), -1, np.int32)
    n = C.c_int32()
    check(self._L.orbx_match_initialization_host(self._h, C.byref(f1), C.byref(f2), prev.ctypes.data, int(windowSize), self.mfNNratio,
                                                 int(self.mbCheckOrientation), m.ctypes.data, C.byref(n)))
    m = m[:f1.n]
    sel = m >= 0                                   # ORBmatcher.cc:513-516
    prev[sel, 0] = F2["keys_un"]["x"][m[sel]]
    prev[sel, 1] = F2["keys_un"]["y"][m[sel]]
    return n.value, m, prev


ORBmatcher.SearchForInitialization = _search_for_initialization

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
                ("th", C.c_float), ("check_ori", C.c_int32), ("match", C.c_void_p), ("nmatches", C.c_void_p)]


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

    def search_frames_device(self, d_jobs, n_jobs, stream=0):
        """batched device-resident SearchByProjection(Cur, Last): d_jobs = device pointer to orbx_frame_match_job[n_jobs]"""
        check(self._L.orbx_match_projection_frame_device(self._h, d_jobs, n_jobs, stream))

    def last_sweeps(self, n_jobs=1):
        out = np.zeros(n_jobs, np.int32)
        check(self._L.orbx_matcher_last_sweeps(self._h, out.ctypes.data, n_jobs))
        return out

"""Host-side mirror of Frame::isInFrustum (reference src/Frame.cc:298-354) over the orbx C ABI, for all points of the local
map at once.  The device evaluates every point; the few whose predicted level sits within a last bit of logf of a level
boundary come back flagged and get MapPoint::PredictScale (MapPoint.cc:444-459) here on the host with libm's logf, exactly the
function the reference calls."""
import ctypes as C
import ctypes.util

import numpy as np

from ._lib import check, lib
from .matcher import TRACK_POINT_DTYPE

FRUSTUM_POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                                ("min_distance", "<f4"), ("max_distance", "<f4"), ("skip", "u1"), ("blocks", "u1"), ("pad", "u1", (2,))])
FRUSTUM_FRAME_DTYPE = np.dtype([("Rcw", "<f4", (9,)), ("tcw", "<f4", (3,)), ("Ow", "<f4", (3,)), ("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"),
                                ("cy", "<f4"), ("bf", "<f4"), ("min_x", "<f4"), ("max_x", "<f4"), ("min_y", "<f4"), ("max_y", "<f4"),
                                ("log_scale_factor", "<f4"), ("n_levels", "<i4"), ("viewing_cos_limit", "<f4")])
assert FRUSTUM_POINT_DTYPE.itemsize == 36 and FRUSTUM_FRAME_DTYPE.itemsize == 108

_libm = None


def _logf(x):
    global _libm
    if _libm is None:
        _libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _libm.logf.restype = C.c_float
        _libm.logf.argtypes = [C.c_float]
    return np.float32(_libm.logf(C.c_float(x)))


def predict_scale(max_distance, dist, log_scale_factor, n_levels):
    """MapPoint::PredictScale(currentDist, Frame*) in float, MapPoint.cc:444-459"""
    ratio = np.float32(max_distance) / np.float32(dist)
    n = int(np.ceil(_logf(ratio) / np.float32(log_scale_factor)))
    return 0 if n < 0 else min(n, n_levels - 1)


def isInFrustum(frame, points, device=0, resolve=True):
    """frame: FRUSTUM_FRAME_DTYPE record; points: FRUSTUM_POINT_DTYPE array -> (TRACK_POINT_DTYPE array, number flagged by the device)"""
    fr = np.ascontiguousarray(frame, FRUSTUM_FRAME_DTYPE).reshape(1)
    pts = np.ascontiguousarray(points, FRUSTUM_POINT_DTYPE)
    out = np.zeros(max(len(pts), 1), TRACK_POINT_DTYPE)
    amb = C.c_int32()
    check(lib().orbx_frustum_host(fr.ctypes.data, len(pts), pts.ctypes.data, out.ctypes.data, C.byref(amb), device))
    out = out[:len(pts)]
    if resolve and amb.value:
        f = fr[0]
        for i in np.nonzero(out["pad"][:, 0])[0]:
            p = pts[i]
            o = np.array([p["x"], p["y"], p["z"]], np.float32) - f["Ow"]
            dist = np.float32(np.sqrt(np.sum(o.astype(np.float64) ** 2)))        # cv::norm(P - mOw)
            out["level"][i] = predict_scale(p["max_distance"], dist, f["log_scale_factor"], int(f["n_levels"]))
            out["pad"][i, 0] = 0
    return out, amb.value

"""Oracle of Frame::isInFrustum (oracle/frustum_oracle.c, reference src/Frame.cc:298-354) against an independent numpy float32
statement (cv2.gemm / cv2.norm for the two OpenCV primitives that are reachable from Python).  CPU only."""
import ctypes as C
import ctypes.util

import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth

cv2 = pytest.importorskip("cv2")
F = np.float32
libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
libm.logf.restype = C.c_float
libm.logf.argtypes = [C.c_float]


def pack(frame, pts):
    fr = np.zeros(1, O.FRUSTUM_FRAME_DTYPE)
    for k, v in frame.items():
        fr[k] = v
    p = np.zeros(len(pts["x"]), O.FRUSTUM_POINT_DTYPE)
    for k, v in pts.items():
        p[k] = v
    return fr[0], p


def numpy_frustum(fr, p):
    out = np.zeros(len(p), O.TRACK_POINT_DTYPE)
    out["blocks"] = p["blocks"]
    R = fr["Rcw"].reshape(3, 3)
    for i in range(len(p)):
        if p["skip"][i]:
            continue
        P = np.array([[p["x"][i]], [p["y"][i]], [p["z"][i]]], F)
        Pc = cv2.gemm(R, P, 1.0, fr["tcw"].reshape(3, 1), 1.0).ravel()             # mRcw*P+mtcw
        if Pc[2] < 0:
            continue
        invz = F(1.0) / Pc[2]
        u, v = fr["fx"] * Pc[0] * invz + fr["cx"], fr["fy"] * Pc[1] * invz + fr["cy"]
        if u < fr["min_x"] or u > fr["max_x"] or v < fr["min_y"] or v > fr["max_y"]:
            continue
        PO = P.ravel() - fr["Ow"]
        dist = F(cv2.norm(PO))
        if dist < F(0.8) * p["min_distance"][i] or dist > F(1.2) * p["max_distance"][i]:
            continue
        Pn = np.array([p["nx"][i], p["ny"][i], p["nz"][i]], F)
        view_cos = F(float(np.dot(PO.astype(np.float64), Pn.astype(np.float64))) / float(dist))
        if view_cos < fr["viewing_cos_limit"]:
            continue
        ratio = p["max_distance"][i] / dist
        n = int(np.ceil(F(libm.logf(C.c_float(ratio))) / fr["log_scale_factor"]))
        n = 0 if n < 0 else min(n, int(fr["n_levels"]) - 1)
        out[i] = (u, v, u - fr["bf"] * invz, view_cos, n, 1, p["blocks"][i], (0, 0))
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_equals_numpy(seed):
    fr, p = pack(*synth.frustum_scene(seed, n=1500))
    got, ref = O.is_in_frustum(fr, p), numpy_frustum(fr, p)
    assert got.tobytes() == ref.tobytes()
    assert 100 < got["in_view"].sum() < 1400 and len(set(got["level"][got["in_view"] == 1].tolist())) >= 6


def test_predict_scale_helper():
    lsf = F(np.log(F(1.2)))
    assert O.predict_scale(10.0, 10.0, lsf, 8) == 0
    assert O.predict_scale(10.0, 1.0, lsf, 8) == 7            # clamped to nlevels - 1
    assert O.predict_scale(1.0, 10.0, lsf, 8) == 0            # negative -> 0
    assert O.predict_scale(12.0, 10.0, lsf, 8) in (1, 2)       # exactly on a boundary: whatever libm's logf says

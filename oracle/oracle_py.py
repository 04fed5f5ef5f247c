"""ctypes binding of the CPU oracle (oracle/liborbx_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (active-orb-slam2_b200/orbx) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    """Compile oracle/*.c -> liborbx_oracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "liborbx_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".inc", "Makefile"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        try:
            so = build()
        except Exception:
            so = os.path.join(_HERE, "liborbx_oracle.so")  # GPU box without a changed tree: use the prebuilt file
        _LIB = C.CDLL(so)
        _declare(_LIB)
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _declare(L):
    u8p, i32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.orbo_extractor_create.restype = C.c_void_p
    L.orbo_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.orbo_extractor_destroy.argtypes = [C.c_void_p]
    L.orbo_extractor_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.orbo_extractor_capacity.argtypes = [C.c_void_p]
    L.orbo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.orbo_level_info.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p]
    L.orbo_level_ptr.restype = C.c_void_p
    L.orbo_level_ptr.argtypes = [C.c_void_p, C.c_int]
    L.orbo_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orbo_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orbo_stage_seconds.argtypes = [C.c_void_p, C.c_void_p]
    L.orbo_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orbo_gaussian7_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.orbo_border_reflect101.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orbo_fast9.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orbo_fast_atan2.restype = C.c_float
    L.orbo_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.orbo_sincos_f.argtypes = [C.c_float, f32p, f32p]
    L.orbo_cv_round_f.argtypes = [C.c_float]
    L.orbo_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    L.orbo_ic_angle.restype = C.c_float
    L.orbo_ic_angle.argtypes = [C.c_void_p, C.c_int]


class Extractor:
    """Mirror of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111) over the C oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.h = self.L.orbo_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("bad extractor parameters")
        self.nlevels = nlevels
        self.cap = self.L.orbo_extractor_capacity(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbo_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        um = np.zeros(16, np.int32)
        self.L.orbo_extractor_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(q), _p(um))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, quota=q, umax=um)

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape if img.ndim == 2 else (0, 0)
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = self.L.orbo_extract(self.h, _p(img), w, h, img.strides[0] if img.ndim == 2 and h else 0, _p(kps), _p(desc), self.cap)
        if n < 0:
            raise RuntimeError("oracle extract failed: %d" % n)
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        w, h, st = C.c_int(), C.c_int(), C.c_int()
        if self.L.orbo_level_info(self.h, l, C.byref(w), C.byref(h), C.byref(st)):
            raise IndexError(l)
        ptr = self.L.orbo_level_ptr(self.h, l)
        buf = (C.c_uint8 * (st.value * h.value)).from_address(ptr)
        a = np.frombuffer(buf, np.uint8).reshape(h.value, st.value)[:, : w.value]
        return a.copy()

    def level_padded(self, l, pad=19):
        w, h, st = C.c_int(), C.c_int(), C.c_int()
        self.L.orbo_level_info(self.h, l, C.byref(w), C.byref(h), C.byref(st))
        ptr = self.L.orbo_level_ptr(self.h, l) - pad * st.value - pad
        buf = (C.c_uint8 * (st.value * (h.value + 2 * pad))).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(h.value + 2 * pad, st.value)[:, : w.value + 2 * pad].copy()

    def candidates(self, l):
        n = self.L.orbo_level_candidates(self.h, l, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.L.orbo_level_candidates(self.h, l, _p(out), n)
        return out[:n]

    def stage_seconds(self):
        t = np.zeros(6, np.float64)
        self.L.orbo_stage_seconds(self.h, _p(t))
        return dict(zip(["pyramid", "fast", "octree", "orient", "blur", "desc"], t.tolist()))


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orbo_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def gaussian7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orbo_gaussian7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def border101(src, pad=19):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    buf = np.zeros((h + 2 * pad, w + 2 * pad), np.uint8)
    buf[pad:pad + h, pad:pad + w] = src
    lib().orbo_border_reflect101(buf.ctypes.data + pad * buf.strides[0] + pad, w, h, buf.strides[0], pad)
    return buf


def fast9(img, threshold, cap=1 << 20):
    img = np.ascontiguousarray(img, np.uint8)
    xs, ys, sc = (np.zeros(cap, np.int32) for _ in range(3))
    n = lib().orbo_fast9(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, _p(xs), _p(ys), _p(sc), cap)
    assert n <= cap
    return xs[:n], ys[:n], sc[:n]


def fast_atan2(y, x):
    return lib().orbo_fast_atan2(float(y), float(x))


def sincos(x):
    s, c = C.c_float(), C.c_float()
    lib().orbo_sincos_f(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def descriptor(blurred, x, y, angle_deg):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orbo_descriptor(blurred.ctypes.data + y * blurred.strides[0] + x, blurred.strides[0], float(angle_deg), _p(d))
    return d


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return lib().orbo_ic_angle(img.ctypes.data + y * img.strides[0] + x, img.strides[0])


def distribute(cands, min_x, max_x, min_y, max_y, n):
    """DistributeOctTree on a KP_DTYPE candidate array (x, y relative to min_x/min_y)."""
    cands = np.ascontiguousarray(cands, KP_DTYPE)
    w, h = max_x - min_x, max_y - min_y
    n_ini = max(int(round(w / max(h, 1))), 1)
    out = np.zeros(max(n + 3, 4 * n_ini) + 8, KP_DTYPE)
    r = lib().orbo_distribute(_p(cands), len(cands), min_x, max_x, min_y, max_y, n, _p(out))
    if r < 0:
        raise RuntimeError("distribute failed: %d" % r)
    return out[:r].copy()


# ---- matchers (oracle/match_oracle.c) --------------------------------------------------------------------------
TRACK_POINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                              ("in_view", "u1"), ("blocks", "u1"), ("pad", "u1", (2,))])
LAST_POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("valid", "u1"),
                             ("blocks", "u1"), ("pad", "u1", (2,))])
assert TRACK_POINT_DTYPE.itemsize == 24 and LAST_POINT_DTYPE.itemsize == 24


class OFrame(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("claimed", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("b", C.c_float),
                ("scale_factors", C.c_void_p), ("nlevels", C.c_int32)]


def _oframe(fr):
    """fr: dict with keys_un, desc, u_right (or None), claimed (or None), bounds=(minx,miny,maxx,maxy), K=(fx,fy,cx,cy,bf,b),
    scale_factors.  Returns (OFrame, keepalive list)."""
    keys = np.ascontiguousarray(fr["keys_un"], KP_DTYPE)
    desc = np.ascontiguousarray(fr["desc"], np.uint8)
    sf = np.ascontiguousarray(fr["scale_factors"], np.float32)
    keep = [keys, desc, sf]
    f = OFrame()
    f.n = len(keys)
    f.keys_un, f.desc, f.scale_factors, f.nlevels = keys.ctypes.data, desc.ctypes.data, sf.ctypes.data, len(sf)
    if fr.get("u_right") is not None:
        ur = np.ascontiguousarray(fr["u_right"], np.float32); keep.append(ur); f.u_right = ur.ctypes.data
    if fr.get("claimed") is not None:
        cl = np.ascontiguousarray(fr["claimed"], np.uint8); keep.append(cl); f.claimed = cl.ctypes.data
    mnx, mny, mxx, mxy = (np.float32(v) for v in fr["bounds"])
    f.min_x, f.min_y, f.max_x, f.max_y = mnx, mny, mxx, mxy
    f.grid_w_inv = np.float32(64) / (mxx - mnx)      # Frame.cc:127-128
    f.grid_h_inv = np.float32(48) / (mxy - mny)
    f.fx, f.fy, f.cx, f.cy, f.bf, f.b = (np.float32(v) for v in fr["K"])
    return f, keep


def _declare_match(L):
    if getattr(L, "_match_declared", False):
        return
    L.orbo_hamming256.argtypes = [C.c_void_p, C.c_void_p]
    L.orbo_features_in_area.argtypes = [C.POINTER(OFrame), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    L.orbo_three_maxima.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
    L.orbo_search_by_projection_points.argtypes = [C.POINTER(OFrame), C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    L.orbo_search_by_projection_frame.argtypes = [C.POINTER(OFrame), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
    L._match_declared = True


def hamming256(a, b):
    L = lib(); _declare_match(L)
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return L.orbo_hamming256(_p(a), _p(b))


def features_in_area(fr, x, y, r, min_level=-1, max_level=-1):
    L = lib(); _declare_match(L)
    f, keep = _oframe(fr)
    out = np.zeros(max(f.n, 1), np.int32)
    n = L.orbo_features_in_area(C.byref(f), x, y, r, min_level, max_level, _p(out))
    return out[:n].copy()


def three_maxima(hist):
    L = lib(); _declare_match(L)
    h = np.ascontiguousarray(hist, np.int32)
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    L.orbo_three_maxima(_p(h), len(h), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def search_by_projection_points(fr, pts, pt_desc, th, nnratio=0.8, match=None):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) -> (nmatches, match[F.n])"""
    L = lib(); _declare_match(L)
    f, keep = _oframe(fr)
    pts = np.ascontiguousarray(pts, TRACK_POINT_DTYPE)
    pd = np.ascontiguousarray(pt_desc, np.uint8)
    m = np.full(f.n, -1, np.int32) if match is None else np.ascontiguousarray(match, np.int32).copy()
    n = L.orbo_search_by_projection_points(C.byref(f), len(pts), _p(pts), _p(pd), th, nnratio, _p(m))
    return n, m


def search_by_projection_frame(cur, last_pts, last_desc, Rcw, tcw, forward, backward, th, check_ori=True, match=None):
    """ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, mono) -> (nmatches, match[Cur.n])"""
    L = lib(); _declare_match(L)
    f, keep = _oframe(cur)
    pts = np.ascontiguousarray(last_pts, LAST_POINT_DTYPE)
    pd = np.ascontiguousarray(last_desc, np.uint8)
    R = np.ascontiguousarray(Rcw, np.float32).reshape(9)
    t = np.ascontiguousarray(tcw, np.float32).reshape(3)
    m = np.full(f.n, -1, np.int32) if match is None else np.ascontiguousarray(match, np.int32).copy()
    n = L.orbo_search_by_projection_frame(C.byref(f), len(pts), _p(pts), _p(pd), _p(R), _p(t), int(forward), int(backward), th,
                                          int(check_ori), _p(m))
    return n, m


# ---- local bundle adjustment (oracle/lba_oracle.c) ------------------------------------------------------------
class OLbaProblem(C.Structure):
    _fields_ = [("n_kf", C.c_int32), ("kf_pose", C.c_void_p), ("kf_fixed", C.c_void_p), ("n_pts", C.c_int32), ("pts", C.c_void_p),
                ("n_edges", C.c_int32), ("e_kf", C.c_void_p), ("e_pt", C.c_void_p), ("e_obs", C.c_void_p),
                ("e_inv_sigma2", C.c_void_p), ("e_stereo", C.c_void_p),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double),
                ("stop_flag", C.c_void_p)]


class OLbaTrace(C.Structure):
    _fields_ = [("n_trials", C.c_int32), ("chi2", C.c_double * 256), ("lambda_", C.c_double * 256), ("dim", C.c_int32),
                ("lambda0", C.c_double), ("Hschur", C.c_void_p), ("bschur", C.c_void_p), ("xp", C.c_void_p),
                ("t_build", C.c_double), ("t_schur", C.c_double), ("t_solve", C.c_double), ("n_builds", C.c_int32)]


def lba_pack(prob, cls=OLbaProblem):
    """prob: dict from synth.lba_problem -> (ctypes struct, keepalive)"""
    keep = dict(kf_pose=np.ascontiguousarray(prob["kf_pose"], np.float64), kf_fixed=np.ascontiguousarray(prob["kf_fixed"], np.uint8),
                pts=np.ascontiguousarray(prob["pts"], np.float64), e_kf=np.ascontiguousarray(prob["e_kf"], np.int32),
                e_pt=np.ascontiguousarray(prob["e_pt"], np.int32), e_obs=np.ascontiguousarray(prob["e_obs"], np.float64),
                e_inv_sigma2=np.ascontiguousarray(prob["e_inv_sigma2"], np.float32), e_stereo=np.ascontiguousarray(prob["e_stereo"], np.uint8))
    P = cls()
    P.n_kf, P.n_pts, P.n_edges = len(keep["kf_pose"]), len(keep["pts"]), len(keep["e_kf"])
    for k, v in keep.items():
        setattr(P, k, v.ctypes.data)
    P.fx, P.fy, P.cx, P.cy, P.bf = prob["K"]
    if prob.get("stop_flag") is not None:
        keep["stop_flag"] = prob["stop_flag"]
        P.stop_flag = prob["stop_flag"].ctypes.data
    return P, keep


def lba_solve(prob, its1=5, its2=10, want_system=False):
    """Optimizer::LocalBundleAdjustment on a synthetic problem -> dict(kf, pts, chi2, erase, trials, chi2_trace, lambda_trace[, Hschur, bschur, xp])"""
    L = lib()
    L.orbo_lba_solve.argtypes = [C.POINTER(OLbaProblem), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OLbaTrace)]
    P, keep = lba_pack(prob)
    kf, pt = np.zeros((P.n_kf, 7)), np.zeros((P.n_pts, 3))
    chi2, erase = np.zeros(P.n_edges), np.zeros(P.n_edges, np.uint8)
    tr = OLbaTrace()
    dim = 6 * int((np.asarray(prob["kf_fixed"]) == 0).sum())
    if want_system:
        Hs, bs, xp = np.zeros((dim, dim)), np.zeros(dim), np.zeros(dim)
        tr.Hschur, tr.bschur, tr.xp = Hs.ctypes.data, bs.ctypes.data, xp.ctypes.data
    rc = L.orbo_lba_solve(C.byref(P), its1, its2, _p(kf), _p(pt), _p(chi2), _p(erase), C.byref(tr))
    n = min(tr.n_trials, 256)
    out = dict(kf=kf, pts=pt, chi2=chi2, erase=erase, trials=tr.n_trials, chi2_trace=np.array(tr.chi2[:n]), lambda_trace=np.array(tr.lambda_[:n]),
               stopped=rc, t_build=tr.t_build, t_schur=tr.t_schur, t_solve=tr.t_solve, n_builds=tr.n_builds)
    if want_system:
        d = tr.dim
        out.update(Hschur=Hs.reshape(-1)[:d * d].reshape(d, d).copy(), bschur=bs[:d].copy(), xp=xp[:d].copy(), lambda0=tr.lambda0)
    return out


# ---- vocabulary-node matchers --------------------------------------------------------------------------------
class OBowSet(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("valid", C.c_void_p),
                ("n_nodes", C.c_int32), ("node_id", C.c_void_p), ("node_start", C.c_void_p), ("node_feat", C.c_void_p)]


class OBucketJob(C.Structure):
    _fields_ = [("a", OBowSet), ("b", OBowSet), ("mode", C.c_int32), ("nnratio", C.c_float), ("check_ori", C.c_int32),
                ("only_stereo", C.c_int32), ("F12", C.c_float * 9), ("ex", C.c_float), ("ey", C.c_float),
                ("sigma2_b", C.c_void_p), ("scale_b", C.c_void_p)]


def fill_bow_set(S, d, valid, keep):
    arrs = dict(keys_un=np.ascontiguousarray(d["keys_un"], KP_DTYPE), desc=np.ascontiguousarray(d["desc"], np.uint8),
                u_right=np.ascontiguousarray(d["u_right"], np.float32), valid=np.ascontiguousarray(valid, np.uint8),
                node_id=np.ascontiguousarray(d["node_id"], np.uint32), node_start=np.ascontiguousarray(d["node_start"], np.int32),
                node_feat=np.ascontiguousarray(d["node_feat"], np.int32))
    keep.append(arrs)
    S.n, S.n_nodes = len(arrs["keys_un"]), len(arrs["node_id"])
    for k, v in arrs.items():
        setattr(S, k, v.ctypes.data)


def bucket_valid(mode, A, B):
    """per-feature map-point tests of the three reference functions"""
    if mode == 2:
        return 1 - A["has_mp"], 1 - B["has_mp"]
    return A["has_mp"], B["has_mp"]


def match_buckets(mode, A, B, nnratio=0.75, check_ori=True, only_stereo=False, F12=None, epipole=(0, 0), sigma2=None, scale=None,
                  use_reference=False):
    """mode 0 SearchByBoW(KF,F), 1 SearchByBoW(KF,KF), 2 SearchForTriangulation -> (nmatches, match_a).
    use_reference: run the reference's own functions on KeyFrames built from the same records (oracle/_ref/liborbmatcher_ref.so)
    -> (nmatches, match_a, the epipole the reference derived from its two poses, to be given to the oracle for mode 2)"""
    L = lib()
    L.orbo_match_buckets.argtypes = [C.POINTER(OBucketJob), C.c_void_p]
    J, keep = OBucketJob(), []
    va, vb = bucket_valid(mode, A, B)
    fill_bow_set(J.a, A, va, keep); fill_bow_set(J.b, B, vb, keep)
    J.mode, J.nnratio, J.check_ori, J.only_stereo = mode, nnratio, int(check_ori), int(only_stereo)
    if F12 is not None:
        J.F12[:] = np.asarray(F12, np.float32).reshape(9).tolist()
    J.ex, J.ey = float(epipole[0]), float(epipole[1])
    s2 = np.ascontiguousarray(sigma2 if sigma2 is not None else np.ones(8), np.float32)
    sc = np.ascontiguousarray(scale if scale is not None else np.ones(8), np.float32)
    J.sigma2_b, J.scale_b = s2.ctypes.data, sc.ctypes.data
    m = np.zeros(max(J.a.n, 1), np.int32)
    if use_reference:
        epi = np.zeros(2, np.float32)
        n = ref_matcher_lib().orbmref_match_buckets(C.byref(J), len(sc), _p(m), _p(epi))
        return n, m[:J.a.n], (float(epi[0]), float(epi[1]))
    n = L.orbo_match_buckets(C.byref(J), _p(m))
    return n, m[:J.a.n]


def search_by_projection_kf(cur, pts, pt_desc, Rcw, tcw, th, orb_dist, check_ori=True, match=None):
    """ORBmatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) -> (nmatches, match[Cur.n])"""
    L = lib(); _declare_match(L)
    L.orbo_search_by_projection_kf.argtypes = [C.POINTER(OFrame), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                               C.c_int, C.c_int, C.c_void_p]
    f, keep = _oframe(cur)
    pts = np.ascontiguousarray(pts, LAST_POINT_DTYPE)
    pd = np.ascontiguousarray(pt_desc, np.uint8)
    R = np.ascontiguousarray(Rcw, np.float32).reshape(9)
    t = np.ascontiguousarray(tcw, np.float32).reshape(3)
    m = np.full(f.n, -1, np.int32) if match is None else np.ascontiguousarray(match, np.int32).copy()
    n = L.orbo_search_by_projection_kf(C.byref(f), len(pts), _p(pts), _p(pd), _p(R), _p(t), th, orb_dist, int(check_ori), _p(m))
    return n, m


WINDOW_POINT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"),
                               ("valid", "u1"), ("pad", "u1", (3,))])


def match_window(F, pts, pt_desc, flags, inv_sigma2, max_dist):
    """window + Hamming core of SearchByProjection(KF, Scw, ...), Fuse x2, SearchBySim3 -> (accepted, best_idx, best_dist)"""
    L = lib(); _declare_match(L)
    L.orbo_match_window.argtypes = [C.POINTER(OFrame), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f, keep = _oframe(F)
    pts = np.ascontiguousarray(pts, WINDOW_POINT_DTYPE)
    pd = np.ascontiguousarray(pt_desc, np.uint8)
    s2 = np.ascontiguousarray(inv_sigma2, np.float32)
    bi, bd = np.zeros(max(len(pts), 1), np.int32), np.zeros(max(len(pts), 1), np.int32)
    n = L.orbo_match_window(C.byref(f), len(pts), _p(pts), _p(pd), flags, _p(s2), max_dist, _p(bi), _p(bd))
    return n, bi[:len(pts)], bd[:len(pts)]


def search_for_initialization(F1, F2, prev_xy, window_size=100, nnratio=0.9, check_ori=True):
    """ORBmatcher::SearchForInitialization -> (nmatches, match12)"""
    L = lib(); _declare_match(L)
    L.orbo_search_for_initialization.argtypes = [C.POINTER(OFrame), C.POINTER(OFrame), C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    f1, k1 = _oframe(F1)
    f2, k2 = _oframe(F2)
    pv = np.ascontiguousarray(prev_xy, np.float32)
    m = np.zeros(max(f1.n, 1), np.int32)
    n = L.orbo_search_for_initialization(C.byref(f1), C.byref(f2), _p(pv), window_size, nnratio, int(check_ori), _p(m))
    return n, m[:f1.n]


# ---- stereo association (oracle/stereo_oracle.c) ---------------------------------------------------------------
class OStereoJob(C.Structure):
    _fields_ = [("n_left", C.c_int32), ("n_right", C.c_int32), ("keys_l", C.c_void_p), ("keys_r", C.c_void_p),
                ("desc_l", C.c_void_p), ("desc_r", C.c_void_p), ("nlevels", C.c_int32), ("scale", C.c_void_p), ("inv_scale", C.c_void_p),
                ("lvl_l", C.c_void_p), ("lvl_r", C.c_void_p), ("lvl_w", C.c_void_p), ("lvl_h", C.c_void_p),
                ("pitch_l", C.c_void_p), ("pitch_r", C.c_void_p), ("bf", C.c_float), ("b", C.c_float)]


def stereo_matches(keys_l, desc_l, keys_r, desc_r, pyr_l, pyr_r, scale, inv_scale, bf, b):
    """Frame::ComputeStereoMatches.  pyr_l / pyr_r: lists of 2-D uint8 arrays (interior of every pyramid level).
    Returns dict(u_right, depth, best_right, sad, kept)."""
    L = lib()
    L.orbo_stereo_matches.argtypes = [C.POINTER(OStereoJob)] + [C.c_void_p] * 4
    kl, kr = np.ascontiguousarray(keys_l, KP_DTYPE), np.ascontiguousarray(keys_r, KP_DTYPE)
    dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
    pl = [np.ascontiguousarray(a, np.uint8) for a in pyr_l]
    pr = [np.ascontiguousarray(a, np.uint8) for a in pyr_r]
    n = len(pl)
    sc, isc = np.ascontiguousarray(scale, np.float32), np.ascontiguousarray(inv_scale, np.float32)
    ptr_l = np.array([a.ctypes.data for a in pl], np.uint64)
    ptr_r = np.array([a.ctypes.data for a in pr], np.uint64)
    lw = np.array([a.shape[1] for a in pl], np.int32)
    lh = np.array([a.shape[0] for a in pl], np.int32)
    sl = np.array([a.strides[0] for a in pl], np.int32)
    sr = np.array([a.strides[0] for a in pr], np.int32)
    J = OStereoJob(len(kl), len(kr), kl.ctypes.data, kr.ctypes.data, dl.ctypes.data, dr.ctypes.data, n, sc.ctypes.data,
                   isc.ctypes.data, ptr_l.ctypes.data, ptr_r.ctypes.data, lw.ctypes.data, lh.ctypes.data, sl.ctypes.data,
                   sr.ctypes.data, np.float32(bf), np.float32(b))
    m = max(len(kl), 1)
    ur, dp = np.zeros(m, np.float32), np.zeros(m, np.float32)
    br, sad = np.zeros(m, np.int32), np.zeros(m, np.int32)
    kept = L.orbo_stereo_matches(C.byref(J), _p(ur), _p(dp), _p(br), _p(sad))
    k = len(kl)
    return dict(u_right=ur[:k], depth=dp[:k], best_right=br[:k], sad=sad[:k], kept=kept)


# ---- pose-only optimisation (oracle/pose_oracle.c) -------------------------------------------------------------
class OPoseProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("Xw", C.c_void_p), ("obs", C.c_void_p), ("inv_sigma2", C.c_void_p), ("pose", C.c_double * 7),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double)]


def pose_pack(prob, cls=OPoseProblem):
    """prob: dict(Xw[n,3], obs[n,3], inv_sigma2[n], pose[7], K=(fx,fy,cx,cy,bf)) -> (struct, keepalive)"""
    Xw = np.ascontiguousarray(prob["Xw"], np.float64)
    obs = np.ascontiguousarray(prob["obs"], np.float64)
    s2 = np.ascontiguousarray(prob["inv_sigma2"], np.float32)
    P = cls()
    P.n, P.Xw, P.obs, P.inv_sigma2 = len(Xw), Xw.ctypes.data, obs.ctypes.data, s2.ctypes.data
    for i, v in enumerate(np.asarray(prob["pose"], np.float64)):
        P.pose[i] = v
    P.fx, P.fy, P.cx, P.cy, P.bf = (float(v) for v in prob["K"][:5])
    return P, [Xw, obs, s2]


def pose_optimize(prob):
    """Optimizer::PoseOptimization -> dict(pose[7], outlier[n], n_inliers, n_bad, trials)"""
    L = lib()
    L.orbo_pose_optimize.argtypes = [C.POINTER(OPoseProblem), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    P, keep = pose_pack(prob)
    pose = np.zeros(7, np.float64)
    out = np.zeros(max(P.n, 1), np.uint8)
    nb, tr = C.c_int32(), C.c_int32()
    r = L.orbo_pose_optimize(C.byref(P), _p(pose), _p(out), C.byref(nb), C.byref(tr))
    return dict(pose=pose, outlier=out[:P.n].copy(), n_inliers=r, n_bad=nb.value, trials=tr.value)


# ---- bag-of-words transform (oracle/bow_oracle.c) ---------------------------------------------------------------
class OVocabulary(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("child_start", C.c_void_p), ("children", C.c_void_p), ("desc", C.c_void_p),
                ("weight", C.c_void_p), ("word_id", C.c_void_p), ("L", C.c_int32)]


def bow_transform(voc, desc, levelsup=4):
    """voc: dict(child_start, children, desc, weight, word_id, L) -> (word[n], node[n], weight[n]) per feature"""
    L = lib()
    L.orbo_bow_transform.restype = None
    L.orbo_bow_transform.argtypes = [C.POINTER(OVocabulary), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    cs, ch = np.ascontiguousarray(voc["child_start"], np.int32), np.ascontiguousarray(voc["children"], np.int32)
    nd, w = np.ascontiguousarray(voc["desc"], np.uint8), np.ascontiguousarray(voc["weight"], np.float64)
    wid = np.ascontiguousarray(voc["word_id"], np.int32)
    V = OVocabulary(len(cs) - 1, cs.ctypes.data, ch.ctypes.data, nd.ctypes.data, w.ctypes.data, wid.ctypes.data, int(voc["L"]))
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(d)
    word, node, wt = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
    L.orbo_bow_transform(C.byref(V), _p(d), n, levelsup, _p(word), _p(node), _p(wt))
    return word[:n], node[:n], wt[:n]


# ---- MapPoint::ComputeDistinctiveDescriptors (oracle/mappoint_oracle.c) -----------------------------------------
def distinctive_descriptors(start, desc):
    """start[n+1] CSR over desc[total,32] -> (best_idx[n], best_median[n])"""
    L = lib()
    L.orbo_distinctive_descriptors.restype = None
    L.orbo_distinctive_descriptors.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    st = np.ascontiguousarray(start, np.int32)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(st) - 1
    bi, bm = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    L.orbo_distinctive_descriptors(n, _p(st), _p(d), _p(bi), _p(bm))
    return bi[:n], bm[:n]


# ---- Frame::isInFrustum (oracle/frustum_oracle.c) ---------------------------------------------------------------
FRUSTUM_POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                                ("min_distance", "<f4"), ("max_distance", "<f4"), ("skip", "u1"), ("blocks", "u1"), ("pad", "u1", (2,))])
FRUSTUM_FRAME_DTYPE = np.dtype([("Rcw", "<f4", (9,)), ("tcw", "<f4", (3,)), ("Ow", "<f4", (3,)), ("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"),
                                ("cy", "<f4"), ("bf", "<f4"), ("min_x", "<f4"), ("max_x", "<f4"), ("min_y", "<f4"), ("max_y", "<f4"),
                                ("log_scale_factor", "<f4"), ("n_levels", "<i4"), ("viewing_cos_limit", "<f4")])
assert FRUSTUM_POINT_DTYPE.itemsize == 36 and FRUSTUM_FRAME_DTYPE.itemsize == 108


def is_in_frustum(frame, pts):
    """frame: FRUSTUM_FRAME_DTYPE record, pts: FRUSTUM_POINT_DTYPE array -> TRACK_POINT_DTYPE array"""
    L = lib()
    L.orbo_is_in_frustum.restype = None
    L.orbo_is_in_frustum.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    fr = np.ascontiguousarray(frame, FRUSTUM_FRAME_DTYPE).reshape(1)
    p = np.ascontiguousarray(pts, FRUSTUM_POINT_DTYPE)
    out = np.zeros(max(len(p), 1), TRACK_POINT_DTYPE)
    L.orbo_is_in_frustum(_p(fr), len(p), _p(p), _p(out))
    return out[:len(p)]


def predict_scale(max_distance, dist, log_scale_factor, n_levels):
    L = lib()
    L.orbo_predict_scale.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    return L.orbo_predict_scale(max_distance, dist, log_scale_factor, n_levels)


# ---- the reference's own ORBextractor (oracle/_ref/liborbextractor_ref.so, `make -C oracle ref`) ----
REF_EXTRACTOR_SO = os.path.join(_HERE, "_ref", "liborbextractor_ref.so")
_REF = None


def ref_extractor_lib():
    """The reference's src/ORBextractor.cc compiled against oracle/cvmini (None when it has not been built: the
    reference tree exists only in the build container; the built file travels with the snapshot)."""
    global _REF
    if _REF is None and os.path.exists(REF_EXTRACTOR_SO):
        lib()                                               # liborbx_oracle.so first: the stand-in's primitives live there
        _REF = C.CDLL(REF_EXTRACTOR_SO)
        _REF.orbref_extractor_create.restype = C.c_void_p
        _REF.orbref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _REF.orbref_extractor_destroy.argtypes = [C.c_void_p]
        _REF.orbref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _REF.orbref_levels.argtypes = [C.c_void_p]
        _REF.orbref_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _REF.orbref_level.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_void_p)]
    return _REF


class RefExtractor:
    """ORB_SLAM2::ORBextractor itself (reference src/ORBextractor.cc:410-1132), for pinning `Extractor` above."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.R = ref_extractor_lib()
        if self.R is None:
            raise RuntimeError("oracle/_ref/liborbextractor_ref.so is missing (make -C oracle ref, needs /root/reference)")
        self.h = self.R.orbref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels, self.cap = nlevels, 4 * nfeatures + 1024

    def close(self):
        if getattr(self, "h", None):
            self.R.orbref_extractor_destroy(self.h)
            self.h = None

    __del__ = close

    def tables(self):
        a = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        self.R.orbref_tables(self.h, *[_p(x) for x in a])
        return dict(zip(["scale", "inv_scale", "sigma2", "inv_sigma2"], a))

    def __call__(self, img):
        img = np.asarray(img, np.uint8)
        h, w = img.shape if img.ndim == 2 and img.size else (0, 0)
        assert h == 0 or img.strides[1] == 1
        kps, desc = np.zeros(self.cap, KP_DTYPE), np.zeros((self.cap, 32), np.uint8)
        n = self.R.orbref_extract(self.h, img.ctypes.data if h else None, w, h, img.strides[0] if h else 0, _p(kps), _p(desc), self.cap)
        assert n <= self.cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        """mvImagePyramid[l] with its 19-pixel pad (the view's parent buffer)"""
        w, h, st, ptr = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
        if self.R.orbref_level(self.h, l, C.byref(w), C.byref(h), C.byref(st), C.byref(ptr)):
            raise IndexError(l)
        base = ptr.value - 19 * st.value - 19
        buf = (C.c_uint8 * (st.value * (h.value + 38))).from_address(base)
        return np.frombuffer(buf, np.uint8).reshape(h.value + 38, st.value)[:, : w.value + 38].copy()


# ---- the reference's own ORBmatcher / Frame / MapPoint (oracle/_ref/liborbmatcher_ref.so, `make -C oracle ref`) ----
REF_MATCHER_SO = os.path.join(_HERE, "_ref", "liborbmatcher_ref.so")
ADAPTER_SO = os.path.join(_HERE, "_ref", "liborbmatcher_adapter.so")
_REFM = None
_ADPM = None
USE_ADAPTER = False          # tests flip this: every ref_* function below then runs the DROP-IN build (adapters + liborbx.so, needs a GPU)


def ref_matcher_lib():
    """src/ORBmatcher.cc + Frame.cc + MapPoint.cc + KeyFrame.cc of the reference compiled against oracle/cvmini (None if not built).
    With USE_ADAPTER set: the same shim and reference objects, but ORBextractor / ORBmatcher / Frame::ComputeStereoMatches /
    Frame::ComputeBoW are the product's adapters on top of the CUDA library (oracle/_ref/liborbmatcher_adapter.so)."""
    global _REFM, _ADPM
    if USE_ADAPTER:
        if _ADPM is None and os.path.exists(ADAPTER_SO):
            lib()
            _ADPM = C.CDLL(ADAPTER_SO)
            _declare_refm(_ADPM)
        return _ADPM
    if _REFM is None and os.path.exists(REF_MATCHER_SO):
        lib()
        _REFM = C.CDLL(REF_MATCHER_SO)
        _declare_refm(_REFM)
    return _REFM


def _declare_refm(R):
    """argument / result types of the C entry points of oracle/orbmatcher_ref_shim.cpp (same for both builds)"""
    R.orbmref_hamming256.argtypes = [C.c_void_p, C.c_void_p]
    R.orbmref_features_in_area.argtypes = [C.POINTER(OFrame), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    R.orbmref_search_by_projection_frame.argtypes = [C.POINTER(OFrame), C.c_int] + [C.c_void_p] * 6 + [
        C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]
    R.orbmref_search_by_projection_points.argtypes = [C.POINTER(OFrame), C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                          C.c_void_p]
    R.orbmref_is_in_frustum.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    R.orbmref_match_buckets.argtypes = [C.POINTER(OBucketJob), C.c_int, C.c_void_p, C.c_void_p]
    R.orbmref_search_for_initialization.argtypes = [C.POINTER(OFrame), C.POINTER(OFrame), C.c_void_p, C.c_int, C.c_float, C.c_int,
                                                        C.c_void_p]
    R.orbmref_distinctive_descriptors.restype = None
    R.orbmref_distinctive_descriptors.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    R.orbvref_load_text.restype = C.c_void_p
    R.orbvref_load_text.argtypes = [C.c_char_p]
    R.orbvref_destroy.argtypes = [C.c_void_p]
    R.orbvref_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8
    R.orbmref_search_by_projection_kf.argtypes = [C.POINTER(OFrame), C.c_int] + [C.c_void_p] * 5 + [
        C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    R.orbmref_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    R.orbmref_window.argtypes = [C.c_int, C.POINTER(OFrame), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 8
    R.orbmref_search_by_sim3.argtypes = [C.POINTER(OFrame), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OFrame), C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    R.orbmref_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_int]
    if hasattr(R, "orbmref_three_maxima"):          # reference build only
        R.orbmref_three_maxima.restype = None
        R.orbmref_three_maxima.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int32)] * 3
    R.orbmref_keyframe_features_in_area.argtypes = [C.POINTER(OFrame), C.c_float, C.c_float, C.c_float, C.c_void_p]
    R.orbmref_extractor_quota_umax.restype = None
    R.orbmref_extractor_quota_umax.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    R.orbvref_compute_bow.restype = None
    R.orbvref_compute_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7


def ref_hamming256(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return ref_matcher_lib().orbmref_hamming256(_p(a), _p(b))


def ref_features_in_area(fr, x, y, r, min_level=-1, max_level=-1):
    """Frame::GetFeaturesInArea of the reference on a grid built by its AssignFeaturesToGrid"""
    f, keep = _oframe(fr)
    out = np.zeros(max(f.n, 1), np.int32)
    n = ref_matcher_lib().orbmref_features_in_area(C.byref(f), x, y, r, min_level, max_level, _p(out))
    return out[:n].copy()


def ref_search_by_projection_frame(cur, last_pts, last_desc, Rcw, tcw, Rlw, tlw, mono, th, nnratio=0.9, check_ori=True):
    """the reference's ORBmatcher::SearchByProjection(Cur, Last, th, mono) -> (nmatches, match[Cur.n]); Tlw decides forward / backward"""
    f, keep = _oframe(cur)
    pts = np.ascontiguousarray(last_pts, LAST_POINT_DTYPE)
    pd = np.ascontiguousarray(last_desc, np.uint8)
    a = [np.ascontiguousarray(v, np.float32).reshape(-1) for v in (Rcw, tcw, Rlw, tlw)]
    m = np.full(f.n, -1, np.int32)
    n = ref_matcher_lib().orbmref_search_by_projection_frame(C.byref(f), len(pts), _p(pts), _p(pd), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]),
                                                            int(mono), th, nnratio, int(check_ori), _p(m))
    return n, m


def ref_search_by_projection_points(fr, pts, pt_desc, th, nnratio=0.8):
    """the reference's ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) -> (nmatches, match[F.n])"""
    f, keep = _oframe(fr)
    pts = np.ascontiguousarray(pts, TRACK_POINT_DTYPE)
    pd = np.ascontiguousarray(pt_desc, np.uint8)
    m = np.full(f.n, -1, np.int32)
    n = ref_matcher_lib().orbmref_search_by_projection_points(C.byref(f), len(pts), _p(pts), _p(pd), th, nnratio, _p(m))
    return n, m


def ref_stereo(left, right, bf, b, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """the reference's two ORBextractors + Frame::ComputeStereoMatches on a rectified pair -> dict(keys, desc, u_right, depth)"""
    left, right = np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(right, np.uint8)
    h, w = left.shape
    cap = 4 * nfeatures + 1024
    k, d = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
    ur, dp = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    n = ref_matcher_lib().orbmref_stereo(_p(left), _p(right), w, h, w, nfeatures, scale_factor, nlevels, ini_th, min_th, bf, b,
                                         _p(k), _p(d), cap, _p(ur), _p(dp))
    assert n <= cap
    return dict(keys=k[:n].copy(), desc=d[:n].copy(), u_right=ur[:n].copy(), depth=dp[:n].copy())


def ref_is_in_frustum(frame, pts):
    """the reference's Frame::isInFrustum (+ MapPoint::PredictScale) on every record -> (TRACK_POINT_DTYPE array, the camera
    centre mOw that the reference's Frame::UpdatePoseMatrices derived from the pose)"""
    fr = np.ascontiguousarray(frame, FRUSTUM_FRAME_DTYPE).reshape(1)
    p = np.ascontiguousarray(pts, FRUSTUM_POINT_DTYPE)
    out, ow = np.zeros(max(len(p), 1), TRACK_POINT_DTYPE), np.zeros(3, np.float32)
    ref_matcher_lib().orbmref_is_in_frustum(_p(fr), len(p), _p(p), _p(out), _p(ow))
    return out[:len(p)], ow


def ref_search_for_initialization(F1, F2, prev_xy, window_size=100, nnratio=0.9, check_ori=True):
    """the reference's ORBmatcher::SearchForInitialization -> (nmatches, match12)"""
    f1, k1 = _oframe(F1)
    f2, k2 = _oframe(F2)
    pv = np.ascontiguousarray(prev_xy, np.float32)
    m = np.full(max(f1.n, 1), -1, np.int32)
    n = ref_matcher_lib().orbmref_search_for_initialization(C.byref(f1), C.byref(f2), _p(pv), window_size, nnratio, int(check_ori), _p(m))
    return n, m[:f1.n]


def ref_distinctive_descriptors(start, desc):
    """the reference's MapPoint::ComputeDistinctiveDescriptors per CSR group -> chosen descriptors [n, 32]"""
    st = np.ascontiguousarray(start, np.int32)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    out = np.zeros((max(len(st) - 1, 1), 32), np.uint8)
    ref_matcher_lib().orbmref_distinctive_descriptors(len(st) - 1, _p(st), _p(d), _p(out))
    return out[:len(st) - 1]


def write_vocabulary_text(path, parent, desc, weight, is_leaf, k, L, scoring=0, weighting=0):
    """a vocabulary in the text format of TemplatedVocabulary::loadFromTextFile (TemplatedVocabulary.h:1349-1440): header `k L scoring
    weighting`, then one line per node `parent is_leaf d0 .. d31 weight`.  No trailing newline (the loader's eof() loop would
    otherwise append an empty node); weights are printed with 17 digits so that the parsed double is the float32 value."""
    lines = ["%d %d %d %d" % (k, L, scoring, weighting)]
    for p, d, w, lf in zip(parent.tolist(), np.asarray(desc, np.uint8).tolist(), np.asarray(weight, np.float32).tolist(), is_leaf.tolist()):
        lines.append("%d %d %s %.17g" % (p, int(lf), " ".join(str(x) for x in d), w))
    with open(path, "w") as f:
        f.write("\n".join(lines))


class RefVocabulary:
    """ORB_SLAM2::ORBVocabulary (DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>) loaded by the reference's own loader"""

    def __init__(self, path):
        self.R = ref_matcher_lib()
        self.h = self.R.orbvref_load_text(path.encode())
        if not self.h:
            raise RuntimeError("the reference could not load %s" % path)

    def close(self):
        if getattr(self, "h", None):
            self.R.orbvref_destroy(self.h)
            self.h = None

    __del__ = close

    def compute_bow(self, desc):
        """Frame::ComputeBoW of a frame with these descriptors -> (BowVector dict, FeatureVector dict)"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        bid, bval, nb = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64), C.c_int32()
        fid, fst, ff, nf = np.zeros(max(n, 1), np.uint32), np.zeros(n + 2, np.int32), np.zeros(max(n, 1), np.uint32), C.c_int32()
        self.R.orbvref_compute_bow(self.h, _p(d), n, _p(bid), _p(bval), C.byref(nb), _p(fid), _p(fst), _p(ff), C.byref(nf))
        return ({int(bid[i]): float(bval[i]) for i in range(nb.value)},
                {int(fid[i]): ff[fst[i]:fst[i + 1]].astype(int).tolist() for i in range(nf.value)})

    def transform(self, desc, levelsup=4):
        """-> (word id per feature, BowVector dict, FeatureVector dict) as Frame::ComputeBoW obtains them"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        word = np.zeros(max(n, 1), np.int32)
        bid, bval, nb = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64), C.c_int32()
        fid, fst, ff, nf = np.zeros(max(n, 1), np.uint32), np.zeros(n + 2, np.int32), np.zeros(max(n, 1), np.uint32), C.c_int32()
        self.R.orbvref_transform(self.h, _p(d), n, levelsup, _p(word), _p(bid), _p(bval), C.byref(nb), _p(fid), _p(fst), _p(ff), C.byref(nf))
        bow = {int(bid[i]): float(bval[i]) for i in range(nb.value)}
        fv = {int(fid[i]): ff[fst[i]:fst[i + 1]].astype(int).tolist() for i in range(nf.value)}
        return word[:n].copy(), bow, fv


def ref_search_by_projection_kf(cur, pts, pt_desc, dist_range, Rcw, tcw, th, orb_dist, nnratio=0.9, check_ori=True):
    """the reference's relocalisation SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist) -> (nmatches, match[Cur.n], gate[n_pts],
    level[n_pts]); gate / level are the distance-invariance gate and MapPoint::PredictScale the adapter evaluates for the oracle"""
    f, keep = _oframe(cur)
    pts = np.ascontiguousarray(pts, LAST_POINT_DTYPE)
    pd = np.ascontiguousarray(pt_desc, np.uint8)
    dr = np.ascontiguousarray(dist_range, np.float32).reshape(-1)
    R, t = np.ascontiguousarray(Rcw, np.float32).reshape(9), np.ascontiguousarray(tcw, np.float32).reshape(3)
    m = np.full(f.n, -1, np.int32)
    gate, level = np.zeros(max(len(pts), 1), np.uint8), np.zeros(max(len(pts), 1), np.int32)
    n = ref_matcher_lib().orbmref_search_by_projection_kf(C.byref(f), len(pts), _p(pts), _p(pd), _p(dr), _p(R), _p(t), th, orb_dist, nnratio,
                                                         int(check_ori), _p(m), _p(gate), _p(level))
    return n, m, gate[:len(pts)], level[:len(pts)]


def ref_extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """ORBextractor::operator() through the matcher-side library (the reference's class, or with USE_ADAPTER the adapter's)"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = 4 * nfeatures + 1024
    k, d = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
    n = ref_matcher_lib().orbmref_extract(_p(img), w, h, w, nfeatures, scale_factor, nlevels, ini_th, min_th, _p(k), _p(d), cap)
    assert n <= cap
    return k[:n].copy(), d[:n].copy()


def ref_window(which, kf, R, t, scale, kf_has, kf_extra, pts, pt_desc, th):
    """the map-mutating window searches on a KeyFrame built by the reference's constructor: which = 0 Fuse(pKF, vpMapPoints, th),
    1 Fuse(pKF, Scw, vpPoints, th, vpReplacePoint), 2 SearchByProjection(pKF, Scw, vpPoints, vpMatched, th); see orbmref_window in
    oracle/orbmatcher_ref_shim.cpp for the scene and the MapPoint* codes.  Returns a dict of the map state afterwards."""
    f, keep = _oframe(kf)
    n = f.n
    p = np.ascontiguousarray(pts, FRUSTUM_POINT_DTYPE)
    pd = np.ascontiguousarray(pt_desc, np.uint8)
    Rm, tv = np.ascontiguousarray(R, np.float32).reshape(9), np.ascontiguousarray(t, np.float32).reshape(3)
    has, extra = np.ascontiguousarray(kf_has, np.uint8), np.ascontiguousarray(kf_extra, np.int32)
    o = dict(kf_slot=np.zeros(n, np.int32), pt_bad=np.zeros(len(p), np.uint8), pt_obs=np.zeros(len(p), np.int32),
             pt_replaced=np.zeros(len(p), np.int32), kfmp_bad=np.zeros(n, np.uint8), kfmp_obs=np.zeros(n, np.int32),
             kfmp_replaced=np.zeros(n, np.int32), aux=np.full(max(n, len(p)), -1, np.int32))
    o["ret"] = ref_matcher_lib().orbmref_window(which, C.byref(f), _p(Rm), _p(tv), scale, _p(has), _p(extra), len(p), _p(p), _p(pd), th,
                                                *[_p(o[k]) for k in ("kf_slot", "pt_bad", "pt_obs", "pt_replaced", "kfmp_bad", "kfmp_obs",
                                                                     "kfmp_replaced", "aux")])
    return o


def ref_search_by_sim3(kf1, R1, t1, p1, kf2, R2, t2, p2, preset12, s12, R12, t12, th):
    """the reference's ORBmatcher::SearchBySim3 on two KeyFrames built by its constructor -> (nFound, match12[kf1.n] = keypoint of kf2)"""
    f1, keep1 = _oframe(kf1)
    f2, keep2 = _oframe(kf2)
    a = [np.ascontiguousarray(v, np.float32).reshape(-1) for v in (R1, t1, R2, t2, R12, t12)]
    q1, q2 = np.ascontiguousarray(p1, FRUSTUM_POINT_DTYPE), np.ascontiguousarray(p2, FRUSTUM_POINT_DTYPE)
    pre = np.ascontiguousarray(preset12, np.int32)
    m = np.full(f1.n, -1, np.int32)
    n = ref_matcher_lib().orbmref_search_by_sim3(C.byref(f1), _p(a[0]), _p(a[1]), _p(q1), C.byref(f2), _p(a[2]), _p(a[3]), _p(q2), _p(pre),
                                                s12, _p(a[4]), _p(a[5]), th, _p(m))
    return n, m


def ref_three_maxima(hist):
    """the reference's ORBmatcher::ComputeThreeMaxima on a histogram of bin sizes"""
    h = np.ascontiguousarray(hist, np.int32)
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    ref_matcher_lib().orbmref_three_maxima(_p(h), len(h), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def ref_keyframe_features_in_area(fr, x, y, r):
    """the reference's KeyFrame::GetFeaturesInArea on a KeyFrame built from the record by its constructor"""
    f, keep = _oframe(fr)
    out = np.zeros(max(f.n, 1), np.int32)
    n = ref_matcher_lib().orbmref_keyframe_features_in_area(C.byref(f), x, y, r, _p(out))
    return out[:n].copy()


def ref_extractor_quota_umax(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """mnFeaturesPerLevel and umax as the reference's ORBextractor constructor computes them"""
    q, u = np.zeros(nlevels, np.int32), np.zeros(16, np.int32)
    ref_matcher_lib().orbmref_extractor_quota_umax(nfeatures, scale_factor, nlevels, ini_th, min_th, _p(q), _p(u))
    return q, u


# ---- the reference's own Optimizer + g2o (oracle/_ref/liboptimizer_ref.so, `make -C oracle ref_opt`) -----------------------
REF_OPTIMIZER_SO = os.path.join(_HERE, "_ref", "liboptimizer_ref.so")
OPT_ADAPTER_SO = os.path.join(_HERE, "_ref", "liboptimizer_adapter.so")
_REFO = None
_ADPO = None


class OptrefLbaGraph(C.Structure):
    """optref_lba_graph (oracle/optimizer_ref_shim.cpp)"""
    _fields_ = [("n_kf", C.c_int32), ("kf_Tcw", C.c_void_p), ("kf_start", C.c_void_p), ("kp_xy_ur", C.c_void_p), ("kp_octave", C.c_void_p),
                ("kp_point", C.c_void_p), ("n_pts", C.c_int32), ("pts", C.c_void_p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("nlevels", C.c_int32), ("scale_factor", C.c_float), ("center_kf", C.c_int32), ("first_kf_id", C.c_int32),
                ("stop_before", C.c_int32)]


class OptrefPoseFrame(C.Structure):
    """optref_pose_frame (oracle/optimizer_ref_shim.cpp)"""
    _fields_ = [("n", C.c_int32), ("kp_xy_ur", C.c_void_p), ("kp_octave", C.c_void_p), ("Xw", C.c_void_p), ("has_point", C.c_void_p),
                ("Tcw", C.c_float * 16), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("nlevels", C.c_int32), ("scale_factor", C.c_float)]


def _declare_refo(R):
    dp, ip = C.c_void_p, C.POINTER(C.c_int)
    R.optref_edge_binary.argtypes = [C.c_int, dp, dp, dp, dp, C.c_float, dp, dp, ip, dp, dp]
    R.optref_edge_pose_only.argtypes = [C.c_int, dp, dp, dp, dp, C.c_float, dp, dp, ip, dp]
    R.optref_se3_oplus.argtypes = [dp, dp, dp]
    R.optref_se3_map.argtypes = [dp, dp, dp]
    R.optref_huber.argtypes = [C.c_double, C.c_double, dp]
    R.optref_to_se3quat.argtypes = [dp, dp]
    R.optref_to_cvmat.argtypes = [dp, dp]
    R.optref_local_ba.argtypes = [C.POINTER(OptrefLbaGraph), dp, dp, dp, dp, dp]
    R.optref_pose_optimization.argtypes = [C.POINTER(OptrefPoseFrame), dp, dp, C.POINTER(C.c_int32)]


def ref_optimizer_lib(adapter=False):
    """The reference's src/Optimizer.cc + src/Converter.cc + Thirdparty/g2o compiled against oracle/eigenmini + oracle/cvmini (None
    if not built).  adapter=True: the drop-in build, where Optimizer::LocalBundleAdjustment / PoseOptimization are the product's
    adapter/Optimizer_orbx.cc on the CUDA library (needs a GPU); the leaf entry points are the reference's in both."""
    global _REFO, _ADPO
    if adapter:
        if _ADPO is None and os.path.exists(OPT_ADAPTER_SO):
            lib()
            _ADPO = C.CDLL(OPT_ADAPTER_SO)
            _declare_refo(_ADPO)
        return _ADPO
    if _REFO is None and os.path.exists(REF_OPTIMIZER_SO):
        lib()
        _REFO = C.CDLL(REF_OPTIMIZER_SO)
        _declare_refo(_REFO)
    return _REFO


def _leaf_decl():
    L = lib()
    dp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.orbo_lba_edge_eval.argtypes = [C.c_int, dp, dp, dp, dp, C.c_float, dp, dp, ip, dp, dp]
    L.orbo_pose_edge_eval.argtypes = [dp, dp, dp, dp, C.c_float, dp, dp, ip, dp]
    L.orbo_se3_oplus.argtypes = [dp, dp, dp]
    L.orbo_se3_map.argtypes = [dp, dp, dp]
    L.orbo_huber.argtypes = [C.c_double, C.c_double, dp]
    L.orbo_to_se3quat.argtypes = [dp, dp]
    L.orbo_to_cvmat.argtypes = [dp, dp]
    return L


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def edge_binary(stereo, pose, X, obs, K, inv_sigma2, ref=False):
    """EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ on one (pose, point, observation): dict(err, chi2, depth_positive, JX, Jxi);
    ref=True evaluates the reference's g2o classes, otherwise the oracle's restatement"""
    D = 3 if stereo else 2
    pose, X, obs, K = _f64(pose), _f64(X), _f64(obs), _f64(K)
    err, chi2, JX, Jxi, dpos = np.zeros(D), np.zeros(1), np.zeros((D, 3)), np.zeros((D, 6)), C.c_int()
    f = ref_optimizer_lib().optref_edge_binary if ref else _leaf_decl().orbo_lba_edge_eval
    f(int(stereo), _p(pose), _p(X), _p(obs), _p(K), float(inv_sigma2), _p(err), _p(chi2), C.byref(dpos), _p(JX), _p(Jxi))
    return dict(err=err, chi2=chi2[0], depth_positive=dpos.value, JX=JX, Jxi=Jxi)


def edge_pose_only(pose, X, obs, K, inv_sigma2, ref=False):
    """EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose (obs[2] < 0: monocular): dict(err, chi2, depth_positive, Jxi)"""
    stereo = not (obs[2] < 0)
    D = 3 if stereo else 2
    pose, X, obs, K = _f64(pose), _f64(X), _f64(obs), _f64(K)
    err, chi2, Jxi, dpos = np.zeros(D), np.zeros(1), np.zeros((D, 6)), C.c_int()
    if ref:
        ref_optimizer_lib().optref_edge_pose_only(int(stereo), _p(pose), _p(X), _p(obs), _p(K), float(inv_sigma2), _p(err), _p(chi2), C.byref(dpos), _p(Jxi))
    else:
        _leaf_decl().orbo_pose_edge_eval(_p(pose), _p(X), _p(obs), _p(K), float(inv_sigma2), _p(err), _p(chi2), C.byref(dpos), _p(Jxi))
    return dict(err=err, chi2=chi2[0], depth_positive=dpos.value, Jxi=Jxi)


def se3_oplus(pose, update, ref=False):
    pose, update, out = _f64(pose), _f64(update), np.zeros(7)
    (ref_optimizer_lib().optref_se3_oplus if ref else _leaf_decl().orbo_se3_oplus)(_p(pose), _p(update), _p(out))
    return out


def se3_map(pose, X, ref=False):
    pose, X, out = _f64(pose), _f64(X), np.zeros(3)
    (ref_optimizer_lib().optref_se3_map if ref else _leaf_decl().orbo_se3_map)(_p(pose), _p(X), _p(out))
    return out


def huber(e2, delta, ref=False):
    rho = np.zeros(3)
    (ref_optimizer_lib().optref_huber if ref else _leaf_decl().orbo_huber)(float(e2), float(delta), _p(rho))
    return rho


def to_se3quat(Tcw, ref=False):
    T, out = np.ascontiguousarray(Tcw, np.float32).reshape(16), np.zeros(7)
    (ref_optimizer_lib().optref_to_se3quat if ref else _leaf_decl().orbo_to_se3quat)(_p(T), _p(out))
    return out


def to_cvmat(pose, ref=False):
    pose, out = _f64(pose), np.zeros(16, np.float32)
    (ref_optimizer_lib().optref_to_cvmat if ref else _leaf_decl().orbo_to_cvmat)(_p(pose), _p(out))
    return out.reshape(4, 4)


def ref_local_ba(g, adapter=False):
    """Optimizer::LocalBundleAdjustment of the reference (or, adapter=True, of the drop-in build) on a KeyFrame / MapPoint graph
    built by the reference's constructors from g = dict(kf_Tcw[n,4,4] f32, kf_start[n+1], kp_xy_ur[m,3] f32, kp_octave[m], kp_point[m],
    pts[L,3] f32, K=(fx,fy,cx,cy,bf), center_kf, first_kf_id[, nlevels, scale_factor, stop_before])
    -> dict(Tcw, pts, kp_kept, pt_bad, kf_role)"""
    R = ref_optimizer_lib(adapter)
    G = OptrefLbaGraph()
    keep = dict(kf_Tcw=np.ascontiguousarray(g["kf_Tcw"], np.float32), kf_start=np.ascontiguousarray(g["kf_start"], np.int32),
                kp_xy_ur=np.ascontiguousarray(g["kp_xy_ur"], np.float32), kp_octave=np.ascontiguousarray(g["kp_octave"], np.int32),
                kp_point=np.ascontiguousarray(g["kp_point"], np.int32), pts=np.ascontiguousarray(g["pts"], np.float32))
    for k, v in keep.items():
        setattr(G, k, v.ctypes.data)
    G.n_kf, G.n_pts = len(keep["kf_Tcw"]), len(keep["pts"])
    G.fx, G.fy, G.cx, G.cy, G.bf = (float(v) for v in g["K"][:5])
    G.nlevels, G.scale_factor = int(g.get("nlevels", 8)), float(g.get("scale_factor", 1.2))
    G.center_kf, G.first_kf_id, G.stop_before = int(g["center_kf"]), int(g.get("first_kf_id", 1)), int(g.get("stop_before", 0))
    m = len(keep["kp_point"])
    Tcw, pts = np.zeros((G.n_kf, 4, 4), np.float32), np.zeros((G.n_pts, 3), np.float32)
    kept, bad, role = np.zeros(max(m, 1), np.uint8), np.zeros(max(G.n_pts, 1), np.uint8), np.zeros(max(G.n_kf, 1), np.uint8)
    R.optref_local_ba(C.byref(G), _p(Tcw), _p(pts), _p(kept), _p(bad), _p(role))
    return dict(Tcw=Tcw, pts=pts, kp_kept=kept[:m], pt_bad=bad[:G.n_pts], kf_role=role[:G.n_kf])


def ref_pose_optimization(f, adapter=False):
    """Optimizer::PoseOptimization of the reference (or of the drop-in build) on a Frame built from f = dict(kp_xy_ur[n,3] f32,
    kp_octave[n], Xw[n,3] f32, has_point[n], Tcw[4,4] f32, K) -> dict(ret, Tcw, outlier, n_bad)"""
    R = ref_optimizer_lib(adapter)
    F = OptrefPoseFrame()
    keep = dict(kp_xy_ur=np.ascontiguousarray(f["kp_xy_ur"], np.float32), kp_octave=np.ascontiguousarray(f["kp_octave"], np.int32),
                Xw=np.ascontiguousarray(f["Xw"], np.float32), has_point=np.ascontiguousarray(f["has_point"], np.uint8))
    for k, v in keep.items():
        setattr(F, k, v.ctypes.data)
    F.n = len(keep["kp_octave"])
    for i, v in enumerate(np.asarray(f["Tcw"], np.float32).reshape(16)):
        F.Tcw[i] = v
    F.fx, F.fy, F.cx, F.cy, F.bf = (float(v) for v in f["K"][:5])
    F.nlevels, F.scale_factor = int(f.get("nlevels", 8)), float(f.get("scale_factor", 1.2))
    Tcw, out, nb = np.zeros((4, 4), np.float32), np.zeros(max(F.n, 1), np.uint8), C.c_int32()
    r = R.optref_pose_optimization(C.byref(F), _p(Tcw), _p(out), C.byref(nb))
    return dict(ret=r, Tcw=Tcw, outlier=out[:F.n].copy(), n_bad=nb.value)

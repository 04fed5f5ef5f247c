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

"""Pins the oracle's OpenCV primitives bit-for-bit against cv2 (the reference delegates to OpenCV at
ORBextractor.cc:81,103,115,119-120,809,814,1086,1120,1122,1127).  CPU only."""
import ctypes
import math

import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth

cv2 = pytest.importorskip("cv2")


def level_sizes(w, h, n=8, sf=1.2):
    t = O.Extractor(1000, sf, n, 20, 7).tables()
    return [(O.lib().orbo_cv_round_f(float(np.float32(w) * t["inv_scale"][l])),
             O.lib().orbo_cv_round_f(float(np.float32(h) * t["inv_scale"][l]))) for l in range(n)]


@pytest.mark.parametrize("wh", [(640, 480), (1241, 376), (752, 480)])
def test_resize_chain_matches_cv2(wh):
    w, h = wh
    img = synth.g_rect(1, w, h)
    sizes = level_sizes(w, h)
    cur_cv, cur_or = img, img
    for (lw, lh) in sizes[1:]:
        nxt_cv = cv2.resize(cur_cv, (lw, lh), interpolation=cv2.INTER_LINEAR)
        nxt_or = O.resize_linear(cur_or, lw, lh)
        assert np.array_equal(nxt_cv, nxt_or), (lw, lh)
        cur_cv, cur_or = nxt_cv, nxt_or


@pytest.mark.parametrize("dst", [(533, 400), (641, 479), (320, 241), (700, 500), (100, 77)])
def test_resize_odd_ratios_match_cv2(dst):
    img = synth.g_noise(3, 640, 480)
    assert np.array_equal(cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR), O.resize_linear(img, *dst))


def test_border_matches_cv2():
    img = synth.g_noise(5, 179, 134)
    ref = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, O.border101(img, 19))


@pytest.mark.parametrize("wh", [(640, 480), (1241, 376), (179, 134), (37, 41)])
def test_gaussian_matches_cv2(wh):
    for kind in ("rect", "noise"):
        img = synth.frame(kind, 7, *wh)
        ref = cv2.GaussianBlur(img, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, O.gaussian7(img))


@pytest.mark.parametrize("th", [20, 7])
@pytest.mark.parametrize("kind", ["rect", "noise", "sparse"])
def test_fast_matches_cv2(th, kind):
    img = synth.frame(kind, 11, 640, 480)
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(img, None)
    xs, ys, sc = O.fast9(img, th)
    assert len(kps) == len(xs) and len(xs) > 0
    ref = np.array([(k.pt[0], k.pt[1], k.response) for k in kps], np.float32)
    got = np.stack([xs, ys, sc], 1).astype(np.float32)
    assert np.array_equal(ref, got)  # same set, same raster order, same scores


def test_fast_small_cells_match_cv2():
    """the reference calls cv::FAST on ~37x37 sub-images (ORBextractor.cc:809): NMS must see zeros outside"""
    img = synth.g_noise(2, 640, 480)
    det = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
    rng = np.random.default_rng(0)
    for _ in range(60):
        x0, y0 = int(rng.integers(0, 600)), int(rng.integers(0, 440))
        cw, ch = int(rng.integers(7, 40)), int(rng.integers(7, 40))
        cell = np.ascontiguousarray(img[y0:y0 + ch, x0:x0 + cw])
        kps = det.detect(cell, None)
        xs, ys, sc = O.fast9(cell, 20)
        ref = np.array([(k.pt[0], k.pt[1], k.response) for k in kps], np.float32).reshape(-1, 3)
        assert np.array_equal(ref, np.stack([xs, ys, sc], 1).astype(np.float32).reshape(-1, 3))


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    ys = rng.integers(-60000, 60000, 20000).astype(np.float32)
    xs = rng.integers(-60000, 60000, 20000).astype(np.float32)
    ys[:50] = 0
    xs[25:75] = 0
    for y, x in zip(ys.tolist(), xs.tolist()):
        assert np.float32(cv2.fastAtan2(y, x)) == np.float32(O.fast_atan2(y, x)), (y, x)


def test_sincos_is_correctly_rounded_and_within_1ulp_of_libm():
    """orbo_sincos_f stands in for std::cos(float)/std::sin(float) (ORBextractor.cc:113).

    glibc's cosf/sinf are not correctly rounded (<=0.56 ulp) and glibc picks FMA / non-FMA variants per
    CPU (ifunc), so the reference's own value is machine dependent in the last ulp.  The oracle (and the
    CUDA path, same operation sequence) returns the correctly rounded value; it must stay within 1 ulp
    of this machine's libm and agree with it on the vast majority of angles."""
    libm = ctypes.CDLL("libm.so.6")
    for f in (libm.cosf, libm.sinf):
        f.restype = ctypes.c_float
        f.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(1)
    factor = np.float32(math.pi / 180.0)
    deg = np.concatenate([rng.uniform(0, 360, 100000).astype(np.float32), np.arange(0, 360.5, 0.5, dtype=np.float32)])
    xs = (deg * factor).astype(np.float32)
    got = np.array([O.sincos(float(x)) for x in xs], np.float32)
    cr = np.stack([np.sin(xs.astype(np.float64)), np.cos(xs.astype(np.float64))], 1).astype(np.float32)
    assert np.array_equal(got, cr)
    lm = np.array([(libm.sinf(float(x)), libm.cosf(float(x))) for x in xs], np.float32)
    ulp = np.abs(got.view(np.int32).astype(np.int64) - lm.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
    assert (ulp > 0).mean() < 0.05


def test_descriptor_steering_matches_cv2_orb_level0():
    """Pins the learned pattern + steering + cvRound + sin/cos of computeOrbDescriptor (ORBextractor.cc:108-147)
    against cv2.ORB.compute at octave 0 (same formula in OpenCV's orb.cpp).

    cv2.ORB blurs a SUB-matrix of its pyramid atlas, which makes OpenCV 4.x take the generic (float-kernel)
    separable filter instead of the 8U fixed-point one; ORB-SLAM2 blurs a clone() (ORBextractor.cc:1085-1086),
    i.e. the fixed-point path pinned by test_gaussian_matches_cv2.  So here the blurred image handed to the
    oracle's descriptor is the float-kernel blur, and only the sampling logic is under test."""
    img = synth.g_rect(4, 640, 480)
    e = O.Extractor()
    kps, desc = e(img)
    lvl0 = kps[kps["octave"] == 0]
    assert len(lvl0) > 100
    orb = cv2.ORB_create(nfeatures=5000, scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2,
                         scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=20)
    cvk = [cv2.KeyPoint(float(k["x"]), float(k["y"]), 31.0, float(k["angle"]), float(k["response"]), 0, -1) for k in lvl0]
    cvk2, cvd = orb.compute(img, cvk)
    assert [a.pt for a in cvk] == [b.pt for b in cvk2]
    fblur = cv2.GaussianBlur(img.astype(np.float32), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    fblur = np.clip(np.rint(fblur), 0, 255).astype(np.uint8)
    got = np.stack([O.descriptor(fblur, int(k["x"]), int(k["y"]), k["angle"]) for k in lvl0])
    assert np.array_equal(got, cvd)
    # and the oracle's own pipeline = fixed-point blur + the same sampling
    blurred = cv2.GaussianBlur(img, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    d0 = desc[kps["octave"] == 0]
    for k, d in zip(lvl0, d0):
        assert np.array_equal(O.descriptor(blurred, int(k["x"]), int(k["y"]), k["angle"]), d)
        assert np.float32(O.ic_angle(img, int(k["x"]), int(k["y"]))) == k["angle"]


def test_ic_angle_matches_cv2_orb_detect():
    """IC_Angle (ORBextractor.cc:77-104) is OpenCV's ICAngles: feed cv2.ORB-detected level-0 keypoints back."""
    img = synth.g_rect(9, 640, 480)
    orb = cv2.ORB_create(nfeatures=300, scaleFactor=1.2, nlevels=1, edgeThreshold=19, scoreType=cv2.ORB_FAST_SCORE)
    kps = orb.detect(img, None)
    assert len(kps) > 50
    for k in kps:
        x, y = int(round(k.pt[0])), int(round(k.pt[1]))
        assert np.float32(O.ic_angle(img, x, y)) == np.float32(k.angle)


def test_tables_match_survey():
    t = O.Extractor(1000, 1.2, 8, 20, 7).tables()
    assert t["quota"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert level_sizes(640, 480) == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    t2 = O.Extractor(2000, 1.2, 8, 20, 7).tables()
    assert t2["quota"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]

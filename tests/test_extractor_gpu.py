"""GPU parity of the extractor (liborbx.so through the C ABI) against the CPU oracle, stage by stage and end to
end.  Bar: bit-exact keypoints (every cv::KeyPoint field), descriptors, counts and order."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth
from orbx.extractor import ORBextractor

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vga():
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=8)
    yield ex
    ex.close()


@pytest.fixture(scope="module")
def oracle_vga():
    return O.Extractor(1000, 1.2, 8, 20, 7)


def assert_same_output(kp, de, rkp, rde, tag=""):
    assert len(kp) == len(rkp), "%s count %d vs oracle %d" % (tag, len(kp), len(rkp))
    for f in O.KP_DTYPE.names:
        a, b = kp[f], rkp[f]
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b), \
            "%s field %s differs at %s" % (tag, f, np.nonzero(a != b)[0][:5])
    assert np.array_equal(de, rde), "%s descriptors differ in rows %s" % (tag, np.nonzero((de != rde).any(1))[0][:5])


def test_tables_match_oracle(vga, oracle_vga):
    t = oracle_vga.tables()
    assert np.array_equal(vga.GetScaleFactors(), t["scale"])
    assert np.array_equal(vga.GetInverseScaleFactors(), t["inv_scale"])
    assert np.array_equal(vga.GetScaleSigmaSquares(), t["sigma2"])
    assert np.array_equal(vga.GetInverseScaleSigmaSquares(), t["inv_sigma2"])
    assert np.array_equal(vga.features_per_level(), t["quota"])
    assert vga.capacity == oracle_vga.cap


@pytest.mark.parametrize("kind,seed", [("rect", 0), ("noise", 1), ("sparse", 2)])
def test_stages_vga(vga, oracle_vga, kind, seed):
    img = synth.frame(kind, seed)
    kp, de = vga(img)
    rkp, rde = oracle_vga(img)
    for l in range(8):
        # pyramid incl. the 19-pixel REFLECT_101 border (mvImagePyramid)
        assert np.array_equal(vga.pyramid_level(l, with_border=True), oracle_vga.level_padded(l)), "pyramid level %d" % l
        # FAST candidates as a set
        x, y, s = vga.candidates(l)
        rc = oracle_vga.candidates(l)
        got = sorted(zip(x.tolist(), y.tolist(), s.tolist()))
        ref = sorted(zip(rc["x"].astype(int).tolist(), rc["y"].astype(int).tolist(), rc["response"].astype(int).tolist()))
        assert got == ref, "FAST candidates level %d: %d vs %d" % (l, len(got), len(ref))
        # quadtree selection, ordered
        x, y, s = vga.level_keypoints(l)
        m = rkp["octave"] == l
        sc = oracle_vga.tables()["scale"][l]
        assert len(x) == int(m.sum()), "octree level %d count %d vs %d" % (l, len(x), int(m.sum()))
        rx = rkp["x"][m] if l == 0 else None
        if l == 0:
            assert np.array_equal(x + 16, rx.astype(np.int32)) and np.array_equal(y + 16, rkp["y"][m].astype(np.int32))
        else:
            assert np.array_equal(((x + 16).astype(np.float32) * sc), rkp["x"][m])
            assert np.array_equal(((y + 16).astype(np.float32) * sc), rkp["y"][m])
        assert np.array_equal(s.astype(np.float32), rkp["response"][m])
        # blur
        assert np.array_equal(vga.blurred_level(l), O.gaussian7(oracle_vga.level(l))), "blur level %d" % l
    assert_same_output(kp, de, rkp, rde, kind)


@pytest.mark.parametrize("kind", ["rect", "noise", "sparse"])
@pytest.mark.parametrize("seed", [3, 4, 5, 6])
def test_end_to_end_vga(vga, oracle_vga, kind, seed):
    img = synth.frame(kind, seed)
    kp, de = vga(img)
    rkp, rde = oracle_vga(img)
    assert_same_output(kp, de, rkp, rde, "%s/%d" % (kind, seed))


def test_flat_and_empty(vga):
    kp, de = vga(synth.g_flat())
    assert len(kp) == 0 and de.shape == (0, 32)      # ORBextractor.cc:1064-1065
    kp, de = vga(np.zeros((0, 0), np.uint8))
    assert len(kp) == 0 and de.shape == (0, 32)      # ORBextractor.cc:1046


def test_batch_matches_single(vga, oracle_vga):
    imgs = [synth.frame(k, s) for k, s in [("rect", 10), ("noise", 11), ("sparse", 12), ("flat", 0), ("rect", 13), ("rect", 10)]]
    kps, des = vga.extract_batch(imgs)
    for i, img in enumerate(imgs):
        rkp, rde = oracle_vga(img)
        assert_same_output(kps[i], des[i], rkp, rde, "batch %d" % i)
    assert vga.last_launches() >= 12


def test_strided_input(vga, oracle_vga):
    big = synth.g_rect(21, 800, 600)
    roi = big[60:540, 80:720]
    assert roi.strides[0] == 800
    kp, de = vga(roi)
    rkp, rde = oracle_vga(np.ascontiguousarray(roi))
    assert_same_output(kp, de, rkp, rde, "roi")


@pytest.mark.parametrize("w,h,nf", [(1241, 376, 2000), (752, 480, 1200), (321, 243, 500), (640, 480, 5000), (640, 480, 30)])
def test_other_geometries(w, h, nf):
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
    orc = O.Extractor(nf, 1.2, 8, 20, 7)
    try:
        for kind, seed in [("rect", 1), ("noise", 2)]:
            img = synth.frame(kind, seed, w, h)
            kp, de = ex(img)
            rkp, rde = orc(img)
            assert_same_output(kp, de, rkp, rde, "%dx%d/%s" % (w, h, kind))
    finally:
        ex.close()


def test_other_parameters():
    for nf, sf, nl, a, b in [(1000, 1.5, 4, 30, 10), (800, 1.1, 10, 12, 5), (2000, 2.0, 3, 20, 7)]:
        ex = ORBextractor(nf, sf, nl, a, b, max_width=640, max_height=480)
        orc = O.Extractor(nf, sf, nl, a, b)
        try:
            img = synth.g_rect(nf)
            kp, de = ex(img)
            rkp, rde = orc(img)
            assert_same_output(kp, de, rkp, rde, str((nf, sf, nl, a, b)))
        finally:
            ex.close()


def test_size_change_reconfigures(vga, oracle_vga):
    small = synth.g_rect(31, 512, 384)
    kp, de = vga(small)
    rkp, rde = oracle_vga(small)
    assert_same_output(kp, de, rkp, rde, "512x384")
    img = synth.g_rect(0)
    kp, de = vga(img)
    rkp, rde = oracle_vga(img)
    assert_same_output(kp, de, rkp, rde, "back to vga")


def test_unsupported_shapes_fail_loudly():
    from orbx._lib import OrbxError
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    try:
        with pytest.raises(OrbxError):
            ex(np.zeros((100, 100), np.uint8))       # level 7 would be 28x28 < border
        with pytest.raises(OrbxError):
            ex(np.zeros((500, 700), np.uint8))       # larger than the handle
    finally:
        ex.close()

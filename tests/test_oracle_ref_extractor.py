"""Pins the extractor oracle (and the CUDA path) against the REFERENCE'S OWN ORBextractor.

oracle/_ref/liborbextractor_ref.so is the reference's src/ORBextractor.cc compiled unmodified (make -C oracle ref) against
the functional OpenCV stand-in oracle/cvmini: everything ORB-SLAM2 wrote itself (tables, level sizes, per-cell FAST loop and
minThFAST retry, DivideNode / DistributeOctTree, IC_Angle, computeOrbDescriptor, concatenation and rescale) runs from the
reference's source; the cv:: primitives are the oracle's restatements, which tests/test_oracle_cv2.py pins against cv2.
Heap order (ORBextractor.cc:684 sorts node pointers) is fixed to creation order by the shim's bump allocator.

 - live comparison (needs the .so, i.e. a snapshot taken from the build container): oracle == reference, every field;
 - tests/golden/ref_extract_digests.txt (tools/make_ref_golden.py, written from the reference build): oracle on CPU and the
   CUDA extractor on the GPU reproduce the reference's results without the reference being present.
"""
import hashlib
import os
import sys

import numpy as np
import pytest

from orbx import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_ref_golden import CASES  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def ref_digests():
    return dict(line.split() for line in open(os.path.join(GOLD, "ref_extract_digests.txt")))


def have_ref():
    from oracle import oracle_py as O
    return O.ref_extractor_lib() is not None


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liborbextractor_ref.so")),
                               reason="oracle/_ref/liborbextractor_ref.so is built from the reference tree (make -C oracle ref)")


def test_oracle_reproduces_the_reference_digests():
    from oracle import oracle_py as O
    d = ref_digests()
    assert len(d) == len(CASES)
    for name, kind, seed, w, h, nf, sf, nl, ith, mth in CASES:
        kp, de = O.Extractor(nf, sf, nl, ith, mth)(synth.frame(kind, seed, w, h))
        assert d[name] == "%d:%s" % (len(kp), sha(kp, de)), name
    g = np.load(os.path.join(GOLD, "ref_extract_kitti_rect_seed1.npz"))
    kp, de = O.Extractor(2000, 1.2, 8, 20, 7)(synth.frame("rect", 1, 1241, 376))
    assert kp.tobytes() == g["kps"].tobytes() and np.array_equal(de, g["desc"])


@needs_ref
def test_reference_still_produces_its_digests():
    """the committed fixture is what the reference build gives (guards the stand-in and the shim)"""
    from oracle import oracle_py as O
    d = ref_digests()
    for name, kind, seed, w, h, nf, sf, nl, ith, mth in CASES[:3] + CASES[12:]:
        ex = O.RefExtractor(nf, sf, nl, ith, mth)
        kp, de = ex(synth.frame(kind, seed, w, h))
        ex.close()
        assert d[name] == "%d:%s" % (len(kp), sha(kp, de)), name


@needs_ref
@pytest.mark.parametrize("params", [(1000, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (500, 1.3, 5, 25, 9), (1500, 1.1, 10, 12, 4)])
def test_tables_equal_the_reference(params):
    from oracle import oracle_py as O
    r, o = O.RefExtractor(*params), O.Extractor(*params)
    tr, to = r.tables(), o.tables()
    r.close()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
        assert tr[k].tobytes() == to[k].tobytes(), k


@needs_ref
@pytest.mark.parametrize("case", [
    ("rect", 21, 640, 480, 1000, 1.2, 8, 20, 7), ("noise", 22, 640, 480, 1000, 1.2, 8, 20, 7), ("sparse", 23, 640, 480, 1000, 1.2, 8, 20, 7),
    ("rect", 24, 1241, 376, 2000, 1.2, 8, 20, 7), ("noise", 25, 1241, 376, 2000, 1.2, 8, 20, 7), ("rect", 26, 752, 480, 1200, 1.2, 8, 20, 7),
    ("rect", 27, 333, 251, 300, 1.2, 8, 20, 7), ("noise", 28, 640, 480, 3000, 1.2, 8, 20, 7), ("rect", 29, 640, 480, 100, 1.2, 8, 20, 7),
    ("sparse", 30, 800, 600, 1000, 1.25, 7, 30, 10), ("rect", 31, 512, 512, 1000, 1.2, 8, 20, 7), ("noise", 32, 200, 150, 400, 1.2, 6, 20, 7),
])
def test_oracle_equals_the_reference_on_fresh_inputs(case):
    """every field of every keypoint, every descriptor byte, and the padded pyramid levels"""
    from oracle import oracle_py as O
    kind, seed, w, h, nf, sf, nl, ith, mth = case
    img = synth.frame(kind, seed, w, h)
    r, o = O.RefExtractor(nf, sf, nl, ith, mth), O.Extractor(nf, sf, nl, ith, mth)
    k1, d1 = r(img)
    k2, d2 = o(img)
    assert len(k1) == len(k2) and len(k1) > 0
    for f in k1.dtype.names:
        assert k1[f].tobytes() == k2[f].tobytes(), f
    assert np.array_equal(d1, d2)
    for l in range(nl):
        assert np.array_equal(r.level(l), o.level_padded(l)), l
    r.close()


@needs_ref
def test_reference_edge_cases():
    from oracle import oracle_py as O
    r, o = O.RefExtractor(), O.Extractor()
    # empty image: silent return (ORBextractor.cc:1046)
    k, d = r(np.zeros((0, 0), np.uint8))
    assert len(k) == 0
    # no corners anywhere: zero keypoints, descriptors released (:1064-1065)
    k, d = r(synth.frame("flat", 0))
    k2, d2 = o(synth.frame("flat", 0))
    assert len(k) == 0 and len(k2) == 0
    # a view of a wider buffer (row stride > width)
    big = synth.frame("rect", 40, 800, 480)
    view = big[:, 64:704]
    k1, d1 = r(view)
    k2, d2 = o(np.ascontiguousarray(view))
    assert k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2) and len(k1) > 900
    # the same handle again (mvImagePyramid is reused between calls)
    k3, d3 = r(synth.frame("noise", 41))
    k4, d4 = o(synth.frame("noise", 41))
    assert k3.tobytes() == k4.tobytes() and np.array_equal(d3, d4)
    r.close()


@pytest.mark.gpu
def test_cuda_extractor_reproduces_the_reference_digests():
    """no oracle and no reference at run time: the CUDA path against the vectors written from the reference build"""
    from orbx.extractor import ORBextractor
    d = ref_digests()
    for name, kind, seed, w, h, nf, sf, nl, ith, mth in CASES:
        ex = ORBextractor(nf, sf, nl, ith, mth, max_width=w, max_height=h, max_batch=1)
        try:
            kp, de = ex(synth.frame(kind, seed, w, h))
            assert d[name] == "%d:%s" % (len(kp), sha(kp, de)), name
        finally:
            ex.close()

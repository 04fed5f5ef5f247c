"""THE DROP-IN, RUN: the product's C++ adapters (active-orb-slam2_b200/adapter/*.cc — ORB_SLAM2::ORBextractor, ORB_SLAM2::ORBmatcher,
Frame::ComputeStereoMatches, Frame::ComputeBoW on top of the C ABI of liborbx.so) linked with the reference's own, unmodified
Frame.cc / MapPoint.cc / KeyFrame.cc / DBoW2 into oracle/_ref/liborbmatcher_adapter.so (make -C oracle ref).  The reference's
objects (Frame, KeyFrame, MapPoint, ORBVocabulary) are built exactly as for oracle/_ref/liborbmatcher_ref.so, and the class members
the reference's Tracking / LocalMapping / LoopClosing would call are called on them — except that they now run on the GPU.
Every result is compared with the CPU oracle (which tests/test_oracle_ref_matcher.py shows equal to the reference's own classes)
and, where both libraries are present, with the reference library directly.  Needs a B200 and the prebuilt library."""
import os

import numpy as np
import pytest

from orbx import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liborbmatcher_adapter.so")),
                                 reason="oracle/_ref/liborbmatcher_adapter.so is built where the reference tree is mounted (make -C oracle ref)")]


@pytest.fixture
def adapter():
    from oracle import oracle_py as O
    O.USE_ADAPTER = True
    try:
        assert O.ref_matcher_lib()._name.endswith("liborbmatcher_adapter.so")     # the drop-in build, not the reference library
        yield O
    finally:
        O.USE_ADAPTER = False


def test_adapter_extractor_is_the_reference_class(adapter):
    """ORB_SLAM2::ORBextractor (adapter/ORBextractor_orbx.cc): same keypoints and descriptors as the reference's class"""
    O = adapter
    for kind, seed, w, h, nf in (("rect", 1, 640, 480, 1000), ("sparse", 2, 640, 480, 1000), ("rect", 3, 1241, 376, 2000), ("noise", 4, 752, 480, 1200)):
        img = synth.frame(kind, seed, w, h)
        k1, d1 = O.ref_extract(img, nf)
        k2, d2 = O.Extractor(nf)(img)
        assert len(k1) == len(k2) > 0 and k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2), (kind, seed)


def test_adapter_descriptor_distance(adapter):
    O = adapter
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert O.ref_hamming256(a, b) == O.hamming256(a, b)


def test_adapter_search_by_projection_last_frame(adapter):
    import test_oracle_ref_matcher as T
    for args in ((3, 0.0, 0, 7.0, 1), (4, 1.0, 0, 7.0, 1), (5, -1.0, 0, 15.0, 1), (6, 1.0, 1, 7.0, 1), (7, 0.0, 0, 15.0, 0)):
        T.test_search_by_projection_last_frame_equals_the_reference(*args)


def test_adapter_search_by_projection_map_points(adapter):
    import test_oracle_ref_matcher as T
    for args in ((10, 1.0, 0.8), (11, 3.0, 0.8), (12, 5.0, 0.8), (13, 3.0, 0.6)):
        T.test_search_by_projection_map_points_equals_the_reference(*args)


def test_adapter_relocalisation_search(adapter):
    import test_oracle_ref_matcher as T
    for args in ((1, 10.0, 100, True), (2, 3.0, 64, True), (3, 10.0, 100, False)):
        T.test_relocalisation_search_equals_the_reference(*args)


def test_adapter_vocabulary_node_matchers(adapter):
    import test_oracle_ref_matcher as T
    for seed in (1, 2):
        T.test_vocabulary_node_matchers_equal_the_reference(seed)


def test_adapter_search_for_initialization(adapter):
    import test_oracle_ref_matcher as T
    T.test_search_for_initialization_equals_the_reference(0)


def test_adapter_compute_stereo_matches(adapter):
    """Frame::ComputeStereoMatches (adapter/Frame_orbx.cc) after the adapter's two extractors: mvuRight / mvDepth bit for bit"""
    import test_oracle_ref_matcher as T
    for args in ((0, 640, 480, 1000), (2, 1241, 376, 2000)):
        T.test_compute_stereo_matches_equals_the_reference(*args)


@pytest.mark.parametrize("which,scale,th", [(0, 1.0, 3.0), (0, 1.0, 6.0), (1, 1.0, 4.0), (1, 1.3, 4.0), (2, 1.0, 10.0), (2, 0.8, 10.0)])
def test_adapter_window_searches_equal_the_reference_library(which, scale, th):
    """Fuse(pKF, vpMapPoints, th), Fuse(pKF, Scw, ...) and SearchByProjection(pKF, Scw, ...): the adapter (projection and gates on the host,
    window + Hamming on the device, map mutations on the host in reference order) against the reference's own functions on the
    same KeyFrame / MapPoint graph: return value, every keypoint's map point, every point's bad flag / observation count /
    replacement, vpReplacePoint / vpMatched"""
    from oracle import oracle_py as O
    import test_oracle_ref_matcher as T
    if O.ref_matcher_lib() is None:
        pytest.skip("needs oracle/_ref/liborbmatcher_ref.so as well")
    for seed in (7, 8):
        scene = T.window_scene(seed)
        ref = O.ref_window(which, *scene[:2], scene[2], scale, *scene[3:], th)
        O.USE_ADAPTER = True
        try:
            got = O.ref_window(which, *scene[:2], scene[2], scale, *scene[3:], th)
        finally:
            O.USE_ADAPTER = False
        assert got["ret"] == ref["ret"] and ref["ret"] > 100, (got["ret"], ref["ret"])
        for k in ("kf_slot", "pt_bad", "pt_obs", "pt_replaced", "kfmp_bad", "kfmp_obs", "kfmp_replaced", "aux"):
            assert np.array_equal(got[k], ref[k]), (k, int((got[k] != ref[k]).sum()))


def test_adapter_search_by_sim3_equals_the_reference_library():
    """SearchBySim3 (both projection directions, preset matches, mutual-consistency check) on two KeyFrames of two maps"""
    from oracle import oracle_py as O
    import test_oracle_ref_matcher as T
    if O.ref_matcher_lib() is None:
        pytest.skip("needs oracle/_ref/liborbmatcher_ref.so as well")
    for seed, s12, th in ((11, 1.15, 7.5), (12, 0.9, 7.5), (13, 1.0, 3.0)):
        scene = T.sim3_scene(seed, s12=s12)
        n_ref, m_ref = O.ref_search_by_sim3(*scene, th)
        O.USE_ADAPTER = True
        try:
            n_got, m_got = O.ref_search_by_sim3(*scene, th)
        finally:
            O.USE_ADAPTER = False
        assert n_got == n_ref and np.array_equal(m_got, m_ref) and n_ref > 40, (n_got, n_ref)


def test_adapter_compute_bow_equals_the_reference_library(tmp_path):
    """Frame::ComputeBoW: the adapter (device tree descent + the reference's BowVector / FeatureVector bookkeeping) against the
    reference's own Frame::ComputeBoW on the same ORBVocabulary object file"""
    from oracle import oracle_py as O
    if O.ref_matcher_lib() is None:
        pytest.skip("needs oracle/_ref/liborbmatcher_ref.so as well")
    args = synth.random_vocabulary(6, k=4, L=6, shuffle=False, prune=0.0)
    path = str(tmp_path / "voc.txt")
    O.write_vocabulary_text(path, *args)
    from orbx.vocabulary import tree_from_parents
    tree = tree_from_parents(*args)
    rng = np.random.default_rng(6)
    feats = np.concatenate([synth.descriptors_near_words(rng, tree, 900), rng.integers(0, 256, (300, 32)).astype(np.uint8)])
    rv = O.RefVocabulary(path)
    bow_ref, fv_ref = rv.compute_bow(feats)
    rv.close()
    O.USE_ADAPTER = True
    try:
        av = O.RefVocabulary(path)
        bow_a, fv_a = av.compute_bow(feats)
        av.close()
    finally:
        O.USE_ADAPTER = False
    assert list(bow_a.keys()) == list(bow_ref.keys()) and bow_a == bow_ref and fv_a == fv_ref and len(bow_ref) > 100

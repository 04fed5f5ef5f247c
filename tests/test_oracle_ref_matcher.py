"""Pins the matcher oracle (and the CUDA matcher) against the REFERENCE'S OWN ORBmatcher / Frame / MapPoint.

oracle/_ref/liborbmatcher_ref.so is the reference's src/ORBmatcher.cc, Frame.cc, MapPoint.cc and KeyFrame.cc compiled unmodified
(make -C oracle ref) against the OpenCV stand-in oracle/cvmini; the shim builds the reference's Frame and MapPoint objects from
the oracle's input records, and the reference's code does the rest: AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea, the pose
algebra on cv::Mat, the search loops, DescriptorDistance, the rotation histogram and ComputeThreeMaxima.

 - live comparison (needs the .so, i.e. a snapshot taken from the build container): oracle == reference;
 - tests/golden/ref_match.npz (tools/make_ref_golden.py, written from the reference build): the oracle on CPU and the CUDA
   matcher on the GPU reproduce the reference's match arrays without the reference being present.
"""
import os
import sys

import numpy as np
import pytest

from orbx import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_ref_golden import MATCH_CASES, POINT_CASES, match_inputs, point_inputs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "ref_match.npz")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liborbmatcher_ref.so")),
                               reason="oracle/_ref/liborbmatcher_ref.so is built from the reference tree (make -C oracle ref)")


def flags(cur, R, t, Rlw, tlw, mono):
    """bForward / bBackward as ORBmatcher.cc:1340-1351 computes them, in float"""
    f = np.float32
    R, t, Rlw, tlw = (np.asarray(v, f) for v in (R, t, Rlw, tlw))
    twc = np.array([-(f(f(R[0, i] * t[0]) + f(R[1, i] * t[1])) + f(R[2, i] * t[2])) for i in range(3)], f)
    tlc2 = f(f(f(f(Rlw[2, 0] * twc[0]) + f(Rlw[2, 1] * twc[1])) + f(Rlw[2, 2] * twc[2])) + tlw[2])
    b = f(cur["K"][5])
    return bool(tlc2 > b and not mono), bool(-tlc2 > b and not mono)


def test_oracle_reproduces_the_reference_matches():
    from oracle import oracle_py as O
    g = np.load(GOLD)
    for case in MATCH_CASES:
        cur, pts, desc, R, t, Rlw, tlw = match_inputs(case)
        fwd, bwd = flags(cur, R, t, Rlw, tlw, case[5])
        n, m = O.search_by_projection_frame(cur, pts, desc, R, t, fwd, bwd, case[6], True)
        assert n == int(g[case[0] + "/n"]) and np.array_equal(m, g[case[0] + "/match"]) and n > 200, case[0]
    for case in POINT_CASES:
        cur, tp, tdesc = point_inputs(case)
        n, m = O.search_by_projection_points(cur, tp, tdesc, case[4], 0.8)
        assert n == int(g[case[0] + "/n"]) and np.array_equal(m, g[case[0] + "/match"]) and n > 200, case[0]
    for case, want in ((MATCH_CASES[1], (True, False)), (MATCH_CASES[2], (False, True)), (MATCH_CASES[3], (False, False))):
        cur, pts, desc, R, t, Rlw, tlw = match_inputs(case)
        assert flags(cur, R, t, Rlw, tlw, case[5]) == want


@needs_ref
def test_descriptor_distance_equals_the_reference():
    from oracle import oracle_py as O
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert O.ref_hamming256(a, b) == O.hamming256(a, b)
    z = np.zeros(32, np.uint8)
    assert O.ref_hamming256(z, ~z) == 256 and O.ref_hamming256(z, z) == 0


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_features_in_area_equals_the_reference(seed):
    """Frame::AssignFeaturesToGrid + GetFeaturesInArea: same members in the same order, incl. windows that leave the image"""
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    fr = synth.random_frame(rng, 1500 if seed else 300)
    if seed == 2:                                  # undistorted bounds that do not start at 0 (Frame::ComputeImageBounds)
        fr["bounds"] = (-12.5, -7.25, 655.0, 490.5)
    nonempty = 0
    for i in range(400):
        x, y, r = rng.uniform(-30, 670), rng.uniform(-30, 510), rng.uniform(0.5, 80)
        lv = [(-1, -1), (0, 3), (2, -1), (1, 2), (3, 3), (0, 0), (7, 7)][i % 7]
        a, b = O.ref_features_in_area(fr, x, y, r, *lv), O.features_in_area(fr, x, y, r, *lv)
        assert np.array_equal(a, b), (x, y, r, lv)
        nonempty += len(a) > 0
    assert nonempty > 100


@needs_ref
@pytest.mark.parametrize("seed,dz,mono,th,check_ori", [(3, 0.0, 0, 7.0, 1), (4, 1.0, 0, 7.0, 1), (5, -1.0, 0, 15.0, 1), (6, 1.0, 1, 7.0, 1),
                                                       (7, 0.0, 0, 15.0, 0), (8, 0.02, 0, 3.0, 1), (9, -0.3, 0, 7.0, 0)])
def test_search_by_projection_last_frame_equals_the_reference(seed, dz, mono, th, check_ori):
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    cur = synth.random_frame(rng, int(rng.integers(500, 1600)))
    pts, desc, R, t = synth.last_frame_points(rng, cur, int(rng.integers(400, 1400)))
    Rlw, tlw = R.copy(), (t + np.array([0, 0, dz], np.float32)).astype(np.float32)
    fwd, bwd = flags(cur, R, t, Rlw, tlw, mono)
    n1, m1 = O.ref_search_by_projection_frame(cur, pts, desc, R, t, Rlw, tlw, mono, th, 0.9, check_ori)
    n2, m2 = O.search_by_projection_frame(cur, pts, desc, R, t, fwd, bwd, th, check_ori)
    assert n1 == n2 and np.array_equal(m1, m2) and n1 > 100, (n1, n2)


@needs_ref
@pytest.mark.parametrize("seed,th,nnratio", [(10, 1.0, 0.8), (11, 3.0, 0.8), (12, 5.0, 0.8), (13, 3.0, 0.6), (14, 1.0, 0.9)])
def test_search_by_projection_map_points_equals_the_reference(seed, th, nnratio):
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    cur = synth.random_frame(rng, int(rng.integers(500, 1600)))
    tp, tdesc = synth.track_points(rng, cur, int(rng.integers(400, 1800)))
    n1, m1 = O.ref_search_by_projection_points(cur, tp, tdesc, th, nnratio)
    n2, m2 = O.search_by_projection_points(cur, tp, tdesc, th, nnratio)
    assert n1 == n2 and np.array_equal(m1, m2) and n1 > 100, (n1, n2)


@needs_ref
def test_reference_still_produces_its_match_vectors():
    from oracle import oracle_py as O
    g = np.load(GOLD)
    cur, pts, desc, R, t, Rlw, tlw = match_inputs(MATCH_CASES[1])
    n, m = O.ref_search_by_projection_frame(cur, pts, desc, R, t, Rlw, tlw, 0, 7.0)
    assert n == int(g["frame/forward/n"]) and np.array_equal(m, g["frame/forward/match"])


@needs_ref
@pytest.mark.parametrize("seed,w,h,nf", [(0, 640, 480, 1000), (1, 640, 480, 1000), (2, 1241, 376, 2000), (3, 752, 480, 1200)])
def test_compute_stereo_matches_equals_the_reference(seed, w, h, nf):
    """the reference's two extractors + Frame::ComputeStereoMatches (Frame.cc:495-669: row table, Hamming search, 11x11 SAD over
    +-5 px on the pyramid level, parabola fit, median cut) against the oracle chain: mvuRight / mvDepth bit for bit"""
    from oracle import oracle_py as O
    world = synth.stereo_world(seed, w, h)
    left, right = world.render(0.0, 0.01, 0.0), world.render(0.0, 0.01, 0.0, right=True)
    b = world.bf / world.fx
    r = O.ref_stereo(left, right, world.bf, b, nf)
    exl, exr = O.Extractor(nf), O.Extractor(nf)
    kl, dl = exl(left)
    kr, dr = exr(right)
    t = exl.tables()
    o = O.stereo_matches(kl, dl, kr, dr, [exl.level(l) for l in range(8)], [exr.level(l) for l in range(8)], t["scale"], t["inv_scale"],
                         world.bf, b)
    assert r["keys"].tobytes() == kl.tobytes() and np.array_equal(r["desc"], dl)
    assert r["u_right"].tobytes() == o["u_right"].tobytes() and r["depth"].tobytes() == o["depth"].tobytes()
    assert int((r["depth"] > 0).sum()) == o["kept"] and o["kept"] > nf // 4
    if seed == 0:                                  # the committed vector that the -m gpu stereo test reproduces is the reference's result
        g = np.load(os.path.join(ROOT, "tests", "golden", "stereo_seed0.npz"))
        assert r["u_right"].tobytes() == g["u_right"].tobytes() and r["depth"].tobytes() == g["depth"].tobytes()


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_is_in_frustum_equals_the_reference(seed):
    """Frame::isInFrustum + MapPoint::PredictScale (Frame.cc:298-354, MapPoint.cc:427-459) on 4000 map points around the camera:
    every record bit-equal (projection, stereo coordinate, viewing cosine, predicted level, all five gates)"""
    from oracle import oracle_py as O
    frame, pts = synth.frustum_scene(seed)
    fr = np.zeros(1, O.FRUSTUM_FRAME_DTYPE)
    for k, v in frame.items():
        fr[k] = v
    p = np.zeros(len(pts["x"]), O.FRUSTUM_POINT_DTYPE)
    for k, v in pts.items():
        p[k] = v
    ref, ow = O.ref_is_in_frustum(fr[0], p)
    assert np.allclose(ow, fr["Ow"][0], atol=1e-6)
    fr["Ow"][0] = ow                               # mOw as the reference computes it (float; the scene generator used double)
    got = O.is_in_frustum(fr[0], p)
    assert got.tobytes() == ref.tobytes()
    assert 500 < got["in_view"].sum() < 3000 and len(set(got["level"][got["in_view"] == 1].tolist())) >= 6


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_vocabulary_node_matchers_equal_the_reference(seed):
    """SearchByBoW(KF, F), SearchByBoW(KF, KF) and SearchForTriangulation (+ CheckDistEpipolarLine) on KeyFrames that the reference's
    own KeyFrame constructor built from the records; the epipole is the one the reference derives from its two poses"""
    from oracle import oracle_py as O
    A, B, F12, epi, s2, sc = synth.bow_pair(seed, 900, 1000, 500)
    total = 0
    for mode, kw in [(0, {}), (1, {}), (2, dict(only_stereo=False)), (2, dict(only_stereo=True)), (0, dict(check_ori=False, nnratio=0.9)),
                     (1, dict(nnratio=0.6)), (2, dict(check_ori=False))]:
        for e_in in ((epi,) if mode != 2 else (epi, (300.0, 200.0))):       # second epipole: inside the image, so the :743-749 gate bites
            n1, m1, e = O.match_buckets(mode, A, B, F12=F12, epipole=e_in, sigma2=s2, scale=sc, use_reference=True, **kw)
            n2, m2 = O.match_buckets(mode, A, B, F12=F12, epipole=e if mode == 2 else e_in, sigma2=s2, scale=sc, **kw)
            assert n1 == n2 and np.array_equal(m1, m2), (mode, kw, e_in)
            total += n1
    assert total > 500


@needs_ref
@pytest.mark.parametrize("seed", [0, 1])
def test_search_for_initialization_equals_the_reference(seed):
    from oracle import oracle_py as O
    from test_match_oracle import init_pair
    F1, F2, prev = init_pair(80 + seed, 700)
    for win, ratio, ori in ((30, 0.9, True), (100, 0.9, True), (100, 0.7, False)):
        n1, m1 = O.ref_search_for_initialization(F1, F2, prev, win, ratio, ori)
        n2, m2 = O.search_for_initialization(F1, F2, prev, win, ratio, ori)
        assert n1 == n2 and np.array_equal(m1, m2) and n1 > 20


@needs_ref
def test_distinctive_descriptors_equal_the_reference():
    """MapPoint::ComputeDistinctiveDescriptors over observations held by the reference's KeyFrames: the chosen descriptor"""
    from oracle import oracle_py as O
    from test_mappoint_oracle import observation_sets, to_csr
    for seed in range(2):
        sets = observation_sets(seed, n_points=150, max_obs=30)
        start, desc = to_csr(sets)
        bi, bm = O.distinctive_descriptors(start, desc)
        ref = O.ref_distinctive_descriptors(start, desc)
        for i in range(len(sets)):
            if bi[i] >= 0:
                assert np.array_equal(ref[i], desc[start[i] + bi[i]]), i
        assert (bi > 0).sum() > 30


@needs_ref
@pytest.mark.parametrize("seed,k,L,levelsup,scoring,weighting", [(0, 10, 4, 2, 0, 0), (1, 10, 3, 4, 0, 0), (2, 4, 6, 4, 0, 0), (3, 10, 5, 1, 1, 1),
                                                                 (4, 2, 7, 3, 0, 2), (5, 10, 4, 0, 4, 3), (6, 10, 6, 4, 0, 0)])
def test_vocabulary_transform_equals_the_reference(seed, k, L, levelsup, scoring, weighting, tmp_path):
    """ORBVocabulary (DBoW2 TemplatedVocabulary<FORB>) loaded by the reference's loadFromTextFile; transform() as Frame::ComputeBoW
    calls it.  Oracle descent + orbx.vocabulary.bow_maps give the same word per feature, the same BowVector (keys and bit-identical
    doubles, every weighting / scoring pair tried) and the same FeatureVector.  Trees list parents before children (the text loader
    indexes m_nodes[parent] while it grows it) and have no leaf above level L - levelsup (the reference leaves `nid` uninitialised
    for such a branch, TemplatedVocabulary.h:1160-1170 + :1231-1262; the oracle and the kernel define it as 0)."""
    from oracle import oracle_py as O
    from orbx.vocabulary import bow_maps, tree_from_parents
    args = synth.random_vocabulary(seed, k=k if L < 6 or k < 10 else 4, L=L, shuffle=False, prune=0.0)
    tree = tree_from_parents(*args, scoring=scoring, weighting=weighting)
    rng = np.random.default_rng(seed)
    feats = np.concatenate([synth.descriptors_near_words(rng, tree, 700), rng.integers(0, 256, (300, 32)).astype(np.uint8)])
    path = str(tmp_path / "voc.txt")
    O.write_vocabulary_text(path, *args, scoring=scoring, weighting=weighting)
    rv = O.RefVocabulary(path)
    w_ref, bow_ref, fv_ref = rv.transform(feats, levelsup)
    rv.close()
    word, node, wt = O.bow_transform(tree, feats, levelsup)
    v, fv = bow_maps(word, node, wt, weighting, scoring)
    assert np.array_equal(w_ref, word) and len(set(word.tolist())) > 20
    assert list(v.keys()) == list(bow_ref.keys()) and v == bow_ref
    assert fv == fv_ref


@needs_ref
@pytest.mark.parametrize("seed,th,orb_dist,check_ori", [(1, 10.0, 100, True), (2, 3.0, 64, True), (3, 10.0, 100, False)])
def test_relocalisation_search_equals_the_reference(seed, th, orb_dist, check_ori):
    """SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1472-1599).  The oracle's entry point
    takes the distance-invariance gate and the predicted level from its caller (the adapter); here both come from the reference's
    own MapPoint::GetMin/MaxDistanceInvariance and PredictScale, so the whole function is compared, NULL / bad / already-found
    map points included."""
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    cur = synth.random_frame(rng, 1200)
    pts, desc, R, t = synth.last_frame_points(rng, cur, 1000)
    Ow = -(R.T.astype(np.float64) @ t.astype(np.float64))
    d = np.linalg.norm(np.stack([pts["x"], pts["y"], pts["z"]], 1).astype(np.float64) - Ow, axis=1)
    sf = np.float32(1.2)
    max_d = (d * rng.uniform(0.6, 1.6, len(d)) * sf ** np.clip(pts["octave"], 0, 7)).astype(np.float32)
    min_d = (max_d / sf ** 7).astype(np.float32)
    n1, m1, gate, level = O.ref_search_by_projection_kf(cur, pts, desc, np.stack([min_d, max_d], 1), R, t, th, orb_dist, 0.9, check_ori)
    p2 = pts.copy()
    p2["valid"] = (pts["valid"] != 0) & (gate != 0)
    p2["octave"] = level
    n2, m2 = O.search_by_projection_kf(cur, p2, desc, R, t, th, orb_dist, check_ori)
    assert n1 == n2 and np.array_equal(m1, m2) and n1 > 100
    assert 600 < gate.sum() < 1000 and len(set(level[gate != 0].tolist())) == 8


@needs_ref
def test_every_ref_library_loads_and_exports_its_entry_points():
    """no compute: dlopen with RTLD_NOW (ctypes) fails on any undefined symbol, which is how a broken drop-in build would show up
    on the GPU box; the drop-in library must load here too (liborbx.so loads without a device)"""
    import ctypes
    from oracle import oracle_py as O
    O.lib()
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    want = {"liborbextractor_ref.so": ["orbref_extractor_create", "orbref_extract", "orbref_level"],
            "liborbmatcher_ref.so": ["orbmref_search_by_projection_frame", "orbmref_window", "orbmref_search_by_sim3", "orbmref_stereo",
                                     "orbvref_compute_bow", "orbmref_three_maxima"],
            "liborbmatcher_adapter.so": ["orbmref_search_by_projection_frame", "orbmref_window", "orbmref_search_by_sim3", "orbmref_stereo",
                                         "orbvref_compute_bow", "orbmref_extract"]}
    for name, syms in want.items():
        path = os.path.join(ref_dir, name)
        if not os.path.exists(path):
            assert name == "liborbmatcher_adapter.so"       # only built where liborbx.so exists
            continue
        L = ctypes.CDLL(path)
        for s_ in syms:
            assert hasattr(L, s_), (name, s_)


@needs_ref
def test_three_maxima_equal_the_reference():
    """ORBmatcher::ComputeThreeMaxima incl. ties, empty histograms and the 10 % rule"""
    from oracle import oracle_py as O
    rng = np.random.default_rng(0)
    cases = [np.zeros(30, np.int32), np.full(30, 5, np.int32), np.array([10] + [0] * 29, np.int32), np.array([100, 9, 10, 11] + [0] * 26, np.int32)]
    cases += [rng.integers(0, 6, 30).astype(np.int32) for _ in range(100)] + [rng.integers(0, 200, 30).astype(np.int32) for _ in range(100)]
    for h in cases:
        assert O.ref_three_maxima(h) == O.three_maxima(h), h.tolist()


@needs_ref
def test_keyframe_features_in_area_equals_the_reference():
    """KeyFrame::GetFeaturesInArea (KeyFrame.cc:630-669) = the frame window search without a level filter"""
    from oracle import oracle_py as O
    rng = np.random.default_rng(4)
    fr = synth.random_frame(rng, 1500)
    hits = 0
    for _ in range(300):
        x, y, r = rng.uniform(-30, 670), rng.uniform(-30, 510), rng.uniform(0.5, 80)
        a, b = O.ref_keyframe_features_in_area(fr, x, y, r), O.features_in_area(fr, x, y, r, -1, -1)
        assert np.array_equal(a, b), (x, y, r)
        hits += len(a) > 0
    assert hits > 100


@needs_ref
@pytest.mark.parametrize("params", [(1000, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (500, 1.3, 5, 25, 9), (1500, 1.1, 10, 12, 4), (123, 1.5, 3, 20, 7)])
def test_extractor_quota_and_umax_equal_the_reference(params):
    """mnFeaturesPerLevel (ORBextractor.cc:436-446) and umax (:454-469), protected members of the reference's class"""
    from oracle import oracle_py as O
    q, u = O.ref_extractor_quota_umax(*params)
    t = O.Extractor(*params).tables()
    assert q.tolist() == t["quota"][:params[2]].tolist() and u.tolist() == t["umax"].tolist()
    assert q.sum() == params[0] or q[-1] == 0


def window_scene(seed, n=1200, n_pts=900):
    """a keyframe, its own map points, and candidate map points that project near its keypoints (position, normal, distance
    range, descriptor), some NULL / bad / already observed by the keyframe, with varying observation counts"""
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    kf = synth.random_frame(rng, n, claimed_frac=0.0)
    fx, fy, cx, cy, bf, b = kf["K"]
    sf = kf["scale_factors"]
    yaw = rng.uniform(-0.05, 0.05)
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = rng.uniform(-0.1, 0.1, 3)
    tgt = rng.integers(0, n, n_pts)
    tgt[-n_pts // 6:] = tgt[:n_pts // 6]                      # several candidates aim at the same keypoint
    k = kf["keys_un"][tgt]
    z = rng.uniform(1.0, 8.0, n_pts)
    stereo = kf["u_right"][tgt] > 0
    kf["u_right"][tgt[stereo]] = (k["x"][stereo] - bf / z[stereo] + rng.normal(0, 0.5, stereo.sum())).astype(np.float32)
    z = np.where(kf["u_right"][tgt] > 0, bf / np.maximum(k["x"] - kf["u_right"][tgt], 1e-3), z)      # duplicates: the depth that was kept
    z[rng.random(n_pts) < 0.03] *= -1
    s = sf[k["octave"]]
    u = k["x"] + rng.normal(0, 0.8, n_pts) * s
    v = k["y"] + rng.normal(0, 0.8, n_pts) * s
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    pw = (pc - t) @ R
    Ow = -(R.T @ t)
    d = np.linalg.norm(pw - Ow, axis=1)
    p = np.zeros(n_pts, O.FRUSTUM_POINT_DTYPE)
    p["x"], p["y"], p["z"] = pw[:, 0], pw[:, 1], pw[:, 2]
    nrm = (pw - Ow) / d[:, None] + rng.normal(0, 0.3, (n_pts, 3))
    nrm[rng.random(n_pts) < 0.08] *= -1
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    p["nx"], p["ny"], p["nz"] = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    p["max_distance"] = d * np.float32(1.2) ** k["octave"] * rng.uniform(0.8, 1.2, n_pts)
    p["min_distance"] = p["max_distance"] / np.float32(1.2) ** 7
    p["skip"] = rng.choice(4, n_pts, p=[0.8, 0.05, 0.05, 0.1])
    p["blocks"] = rng.integers(0, 4, n_pts)
    desc = synth.flip_bits(rng, kf["desc"][tgt], rng.integers(0, 70, n_pts))
    has = (rng.random(n) < 0.5).astype(np.uint8)
    has[(has == 1) & (rng.random(n) < 0.2)] = 2
    extra = rng.integers(0, 4, n).astype(np.int32)
    return kf, R.astype(np.float32), t.astype(np.float32), has, extra, p, desc


@needs_ref
@pytest.mark.parametrize("which,scale,th", [(0, 1.0, 3.0), (1, 1.0, 4.0), (1, 1.3, 4.0), (2, 1.0, 10.0), (2, 0.8, 10.0)])
def test_reference_window_searches_run_on_the_scene(which, scale, th):
    """Fuse / Fuse(Scw) / SearchByProjection(KF, Scw) of the reference on the scene that tests/test_adapter_dropin_gpu.py compares the
    drop-in build on: the scene exercises every outcome (new observation, replacement in both directions, skips)"""
    from oracle import oracle_py as O
    kf, R, t, has, extra, p, desc = window_scene(7)
    o = O.ref_window(which, kf, R, t, scale, has, extra, p, desc, th)
    assert o["ret"] > 150
    if which == 0:
        assert o["pt_bad"][p["skip"] != 2].sum() > 10 and o["kfmp_bad"].sum() > 10          # replaced in both directions
        assert ((o["kf_slot"] >= 0) & (o["kf_slot"] < 1000000)).sum() > 50                 # candidates now held by the keyframe
    elif which == 1:
        assert (o["aux"][:len(p)] >= 1000000).sum() > 20 and ((o["kf_slot"] >= 0) & (o["kf_slot"] < 1000000)).sum() > 50
        assert o["pt_bad"][p["skip"] != 2].sum() == 0                                       # this overload leaves the replacing to its caller
    else:
        assert ((o["aux"][:len(has)] >= 0) & (o["aux"][:len(has)] < 1000000)).sum() == o["ret"]


def sim3_scene(seed, n=1000, n_common=600, s12=1.15):
    """two keyframes of two maps related by a Sim3: keyframe 2's points, moved by (s12, R12, t12), project near keyframe 1's keypoints"""
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    k2 = synth.random_frame(rng, n, claimed_frac=0.0)
    k1 = synth.random_frame(rng, n, claimed_frac=0.0)
    fx, fy, cx, cy, bf, b = k2["K"]
    sf = k2["scale_factors"]

    def rot(yaw, pitch):
        Ry = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(pitch), -np.sin(pitch)], [0, np.sin(pitch), np.cos(pitch)]])
        return Ry @ Rx
    R1, t1 = rot(rng.uniform(-0.1, 0.1), rng.uniform(-0.05, 0.05)), rng.uniform(-0.3, 0.3, 3)
    R2, t2 = rot(rng.uniform(-0.1, 0.1), rng.uniform(-0.05, 0.05)), rng.uniform(-0.3, 0.3, 3)
    R12, t12 = rot(rng.uniform(-0.05, 0.05), rng.uniform(-0.03, 0.03)), rng.uniform(-0.1, 0.1, 3)
    a = rng.permutation(n)[:n_common]                       # keypoints of keyframe 2 with a counterpart ...
    c = rng.permutation(n)[:n_common]                       # ... at these keypoints of keyframe 1
    z = rng.uniform(1.5, 8.0, n)
    Xc2 = np.stack([(k2["keys_un"]["x"] - cx) / fx * z, (k2["keys_un"]["y"] - cy) / fy * z, z], 1)
    Xc1 = s12 * (Xc2[a] @ R12.T) + t12                      # camera-1 coordinates of keyframe 2's points
    ok = Xc1[:, 2] > 0.5
    u1, v1 = fx * Xc1[:, 0] / Xc1[:, 2] + cx, fy * Xc1[:, 1] / Xc1[:, 2] + cy
    ok &= (u1 > 20) & (u1 < 620) & (v1 > 20) & (v1 < 460)
    a, c, Xc1, u1, v1 = a[ok], c[ok], Xc1[ok], u1[ok], v1[ok]
    k1["keys_un"]["x"][c] = u1 + rng.normal(0, 1.0, len(c))
    k1["keys_un"]["y"][c] = v1 + rng.normal(0, 1.0, len(c))
    k1["keys_un"]["octave"][c] = np.clip(k2["keys_un"]["octave"][a] + rng.integers(-1, 2, len(c)), 0, 7)
    k1["desc"][c] = synth.flip_bits(rng, k2["desc"][a], rng.integers(0, 80, len(c)))
    z1 = rng.uniform(1.5, 8.0, n)
    X1c = np.stack([(k1["keys_un"]["x"] - cx) / fx * z1, (k1["keys_un"]["y"] - cy) / fy * z1, z1], 1)
    X1c[c] = Xc1 + rng.normal(0, 0.01, Xc1.shape)
    pts = []
    for Xc, R, t, kk in ((X1c, R1, t1, k1), (Xc2, R2, t2, k2)):
        Xw = (Xc - t) @ R
        d = np.linalg.norm(Xc, axis=1)
        p = np.zeros(n, O.FRUSTUM_POINT_DTYPE)
        p["x"], p["y"], p["z"] = Xw[:, 0], Xw[:, 1], Xw[:, 2]
        p["max_distance"] = d * np.float32(1.2) ** kk["keys_un"]["octave"] * rng.uniform(0.8, 1.25, n)
        p["min_distance"] = p["max_distance"] / np.float32(1.2) ** 7
        p["skip"] = rng.choice(3, n, p=[0.85, 0.1, 0.05])
        pts.append(p)
    preset = np.full(n, -1, np.int32)
    pick = rng.random(len(c)) < 0.1
    preset[c[pick]] = a[pick]
    preset[c[pick][pts[1]["skip"][a[pick]] == 1]] = -1       # no point there to preset
    f = np.float32
    return (k1, R1.astype(f), t1.astype(f), pts[0], k2, R2.astype(f), t2.astype(f), pts[1], preset, float(s12), R12.astype(f), t12.astype(f))


@needs_ref
def test_reference_search_by_sim3_runs_on_the_scene():
    from oracle import oracle_py as O
    for seed, s12 in ((11, 1.15), (12, 0.9)):
        scene = sim3_scene(seed, s12=s12)
        n, m = O.ref_search_by_sim3(*scene, 7.5)
        assert n > 60 and (m >= 0).sum() >= n


@pytest.mark.gpu
def test_cuda_matcher_reproduces_the_reference_matches():
    """no oracle and no reference at run time: the CUDA matcher against the vectors written from the reference build"""
    from orbx.matcher import ORBmatcher
    g = np.load(GOLD)
    m = ORBmatcher(0.9, True, max_keypoints=2048, max_points=2048)
    try:
        for case in MATCH_CASES:
            cur, pts, desc, R, t, Rlw, tlw = match_inputs(case)
            fwd, bwd = flags(cur, R, t, Rlw, tlw, case[5])
            n, mm = m.SearchByProjectionLast(cur, pts, desc, R, t, fwd, bwd, case[6])
            assert n == int(g[case[0] + "/n"]) and np.array_equal(mm, g[case[0] + "/match"]), case[0]
        m.mfNNratio = 0.8
        for case in POINT_CASES:
            cur, tp, tdesc = point_inputs(case)
            n, mm = m.SearchByProjection(cur, tp, tdesc, case[4])
            assert n == int(g[case[0] + "/n"]) and np.array_equal(mm, g[case[0] + "/match"]), case[0]
    finally:
        m.close()

"""Golden fixtures (tests/golden, written by tools/make_golden.py from the CPU oracle on seeded inputs).
CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them without consulting the oracle."""
import hashlib
import os

import numpy as np
import pytest

from orbx import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def digests():
    return dict(line.split() for line in open(os.path.join(GOLD, "digests.txt")))


def lba_rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def test_oracle_reproduces_golden():
    from oracle import oracle_py as O
    ex = O.Extractor(1000, 1.2, 8, 20, 7)
    d = digests()
    for kind in ("rect", "noise", "sparse"):
        for seed in range(4):
            kp, de = ex(synth.frame(kind, seed))
            assert d["extract/%s/%d" % (kind, seed)] == "%d:%s" % (len(kp), sha(kp, de))
    g = np.load(os.path.join(GOLD, "extract_vga_rect_seed0.npz"))
    kp, de = ex(synth.frame("rect", 0))
    assert kp.tobytes() == g["kps"].tobytes() and np.array_equal(de, g["desc"])
    p = synth.lba_problem(5, n_kf=8, n_pts=400, n_fixed=1)
    r = O.lba_solve(p, 5, 10)
    g = np.load(os.path.join(GOLD, "lba_seed5.npz"))
    assert r["trials"] == int(g["trials"]) and np.array_equal(r["erase"], g["erase"]) and lba_rel(r["pts"], g["pts"]) < 1e-12
    p = synth.lba_problem(6, n_kf=6, n_pts=300, stereo=True, n_fixed=1)
    r = O.lba_solve(p, 5, 10)
    g = np.load(os.path.join(GOLD, "lba_stereo_seed6.npz"))
    assert np.array_equal(r["erase"], g["erase"]) and lba_rel(r["pts"], g["pts"]) < 1e-12 and lba_rel(r["kf"], g["kf"]) < 1e-12
    st = oracle_stereo_seed0(O)
    g = np.load(os.path.join(GOLD, "stereo_seed0.npz"))
    assert st["u_right"].tobytes() == g["u_right"].tobytes() and st["depth"].tobytes() == g["depth"].tobytes() and st["kept"] == int(g["kept"])
    r = O.pose_optimize(synth.pose_problem(3, n=300))
    g = np.load(os.path.join(GOLD, "pose_seed3.npz"))
    assert np.array_equal(r["outlier"], g["outlier"]) and r["n_inliers"] == int(g["n_inliers"]) and lba_rel(r["pose"], g["pose"]) < 1e-12


def oracle_stereo_seed0(O):
    world = synth.stereo_world(0)
    exl, exr = O.Extractor(1000, 1.2, 8, 20, 7), O.Extractor(1000, 1.2, 8, 20, 7)
    kl, dl = exl(world.render(0.0, 0.01, 0.0))
    kr, dr = exr(world.render(0.0, 0.01, 0.0, right=True))
    t = exl.tables()
    return O.stereo_matches(kl, dl, kr, dr, [exl.level(l) for l in range(8)], [exr.level(l) for l in range(8)], t["scale"], t["inv_scale"],
                            world.bf, world.bf / world.fx)


@pytest.mark.gpu
def test_cuda_stereo_and_pose_reproduce_golden():
    """no oracle at run time: the CUDA path against the committed vectors"""
    from orbx.extractor import ORBextractor
    from orbx.optimizer import Optimizer, PoseOptimizer
    from orbx.stereo import StereoMatcher
    world = synth.stereo_world(0)
    el, er, sm = ORBextractor(1000, 1.2, 8, 20, 7), ORBextractor(1000, 1.2, 8, 20, 7), StereoMatcher(max_keypoints=2048)
    po, o = PoseOptimizer(max_observations=1000, max_frames=1), Optimizer(max_keyframes=16, max_points=1000, max_edges=5000)
    try:
        kl, dl = el(world.render(0.0, 0.01, 0.0))
        kr, dr = er(world.render(0.0, 0.01, 0.0, right=True))
        ur, dp, kept = sm.ComputeStereoMatches(el, er, kl, dl, kr, dr, world.bf, world.bf / world.fx)
        g = np.load(os.path.join(GOLD, "stereo_seed0.npz"))
        assert ur.tobytes() == g["u_right"].tobytes() and dp.tobytes() == g["depth"].tobytes() and kept == int(g["kept"])
        p = synth.pose_problem(3, n=300)
        r = po.PoseOptimization(p)
        g = np.load(os.path.join(GOLD, "pose_seed3.npz"))
        assert np.array_equal(r["outlier"], g["outlier"]) and r["n_inliers"] == int(g["n_inliers"])
        assert lba_rel(r["pose"] - p["pose"], g["pose"] - p["pose"]) < 1e-4
        p = synth.lba_problem(6, n_kf=6, n_pts=300, stereo=True, n_fixed=1)
        g = np.load(os.path.join(GOLD, "lba_stereo_seed6.npz"))
        r = o.LocalBundleAdjustment(p, 5, 10)
        assert np.array_equal(r["erase"], g["erase"])
        assert lba_rel(r["kf"] - p["kf_pose"], g["kf"] - p["kf_pose"]) < 1e-4 and lba_rel(r["pts"] - p["pts"], g["pts"] - p["pts"]) < 1e-4
    finally:
        for h in (el, er, sm, po, o):
            h.close()


@pytest.mark.gpu
def test_cuda_extractor_reproduces_golden():
    from orbx.extractor import ORBextractor
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=4)
    d = digests()
    try:
        for kind in ("rect", "noise", "sparse"):
            kps, des = ex.extract_batch([synth.frame(kind, s) for s in range(4)])
            for seed in range(4):
                assert d["extract/%s/%d" % (kind, seed)] == "%d:%s" % (len(kps[seed]), sha(kps[seed], des[seed])), (kind, seed)
            g = np.load(os.path.join(GOLD, "extract_vga_%s_seed0.npz" % kind))
            assert kps[0].tobytes() == g["kps"].tobytes() and np.array_equal(des[0], g["desc"])
    finally:
        ex.close()


@pytest.mark.gpu
def test_cuda_matchers_reproduce_golden():
    from orbx.matcher import ORBmatcher
    m = ORBmatcher(0.8, True, max_keypoints=2048, max_points=2048)
    try:
        rng = np.random.default_rng(1234)
        cur = synth.random_frame(rng, 800)
        pts, desc, R, t = synth.last_frame_points(rng, cur, 700)
        g = np.load(os.path.join(GOLD, "match_projection_seed1234.npz"))
        n, mm = m.SearchByProjectionLast(cur, pts, desc, R, t, False, False, 7.0)
        assert n == int(g["n_frame"]) and np.array_equal(mm, g["match_frame"])
        tp, tdesc = synth.track_points(rng, cur, 700)
        n, mm = m.SearchByProjection(cur, tp, tdesc, 3.0)
        assert n == int(g["n_points"]) and np.array_equal(mm, g["match_points"])
        A, B, F12, epi, s2, sc = synth.bow_pair(77, 600, 650, 350, n_nodes=40)
        g = np.load(os.path.join(GOLD, "match_buckets_seed77.npz"))
        m.mfNNratio = 0.75
        n, ma = m.SearchByBoWKF(A, B)
        assert n == int(g["n1"]) and np.array_equal(ma, g["m1"])
        n, pairs = m.SearchForTriangulation(A, B, F12, epi, s2, sc, False)
        idx1 = np.nonzero(g["m2"] >= 0)[0]
        assert n == int(g["n2"]) and np.array_equal(pairs, np.stack([idx1, g["m2"][idx1]], 1))
    finally:
        m.close()


@pytest.mark.gpu
def test_cuda_lba_reproduces_golden():
    from orbx.optimizer import Optimizer
    o = Optimizer(max_keyframes=16, max_points=1000, max_edges=5000)
    try:
        p = synth.lba_problem(5, n_kf=8, n_pts=400, n_fixed=1)
        g = np.load(os.path.join(GOLD, "lba_seed5.npz"))
        r = o.LocalBundleAdjustment(p, 5, 10, want_system=True)
        assert r["trials"] == int(g["trials"]) and np.array_equal(r["erase"], g["erase"])
        assert lba_rel(r["Hschur"], g["Hschur"]) < 1e-9 and lba_rel(r["bschur"], g["bschur"]) < 1e-9
        assert lba_rel(r["kf"] - p["kf_pose"], g["kf"] - p["kf_pose"]) < 1e-4       # north_star tolerance on the updates
        assert lba_rel(r["pts"] - p["pts"], g["pts"] - p["pts"]) < 1e-4
    finally:
        o.close()

"""Pins the two optimiser oracles (oracle/lba_oracle.c, oracle/pose_oracle.c) against the REFERENCE'S OWN Optimizer + g2o.

oracle/_ref/liboptimizer_ref.so is the reference's src/Optimizer.cc, src/Converter.cc, Frame / KeyFrame / MapPoint and the whole vendored
Thirdparty/g2o (core, types, solvers, stuff) compiled unmodified (make -C oracle ref_opt) against two stand-ins: oracle/cvmini for
OpenCV and oracle/eigenmini for Eigen (this image has neither).  Two layers:

 - leaves: every g2o edge type (EdgeSE3ProjectXYZ, EdgeStereoSE3ProjectXYZ and the two OnlyPose edges: computeError, chi2,
   isDepthPositive, linearizeOplus), VertexSE3Expmap::oplusImpl / SE3Quat::exp, SE3Quat::map, RobustKernelHuber::robustify and
   Converter::toSE3Quat / toCvMat evaluated by the reference's classes on 10^4 random inputs: the oracle's functions (the ones its
   solvers call) return the SAME BITS;
 - whole functions: Optimizer::LocalBundleAdjustment on KeyFrame / MapPoint / Map graphs built with the reference's constructors,
   AddObservation, AddMapPoint and UpdateConnections, and Optimizer::PoseOptimization on a Frame: same erased observations / bad
   points / outlier flags / return values, float poses and points equal to within float rounding of the result.
"""
import os

import numpy as np
import pytest

from orbx import synth

from lba_graph import expected_kept, graph_from_problem, inv_sigma2_table, problem_from_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liboptimizer_ref.so")),
                               reason="oracle/_ref/liboptimizer_ref.so is built from the reference tree (make -C oracle ref_opt)")
K = np.array([517.3, 516.5, 318.6, 255.3, 40.0], np.float32).astype(np.float64)


def _cases(n):
    """n random (pose, point, observation, inv_sigma2) tuples incl. points behind the camera and near-zero depth"""
    out = []
    for i in range(n // 5):
        p = synth.pose_problem(i, n=5)
        for j in range(5):
            X = p["Xw"][j].copy()
            if (5 * i + j) % 97 == 0:
                X = -X                                     # behind the camera: isDepthPositive false
            out.append((p["pose"], X, p["obs"][j], p["inv_sigma2"][j]))
    return out


@needs_ref
def test_binary_edges_bit_equal_to_g2o():
    """EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ (types_six_dof_expmap.cpp:103-157, 188-234): error, chi2, depth test, both
    Jacobian blocks; the stereo edge's float `invz` / float `bf` quirk included"""
    from oracle import oracle_py as O
    n_neg = 0
    for pose, X, obs, is2 in _cases(10000):
        obs = obs.copy()
        if obs[2] < 0:
            obs[2] = obs[0] - 3.0
        for st in (0, 1):
            a, b = O.edge_binary(st, pose, X, obs, K, is2), O.edge_binary(st, pose, X, obs, K, is2, ref=True)
            for k in ("err", "JX", "Jxi"):
                assert np.array_equal(a[k], b[k]), (k, a[k], b[k])
            assert a["chi2"] == b["chi2"] and a["depth_positive"] == b["depth_positive"]
            n_neg += not b["depth_positive"]
    assert n_neg > 50


@needs_ref
def test_pose_only_edges_bit_equal_to_g2o():
    """EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose (types_six_dof_expmap.cpp:266-364)"""
    from oracle import oracle_py as O
    n_st = 0
    for pose, X, obs, is2 in _cases(10000):
        a, b = O.edge_pose_only(pose, X, obs, K, is2), O.edge_pose_only(pose, X, obs, K, is2, ref=True)
        assert np.array_equal(a["err"], b["err"]) and np.array_equal(a["Jxi"], b["Jxi"])
        assert a["chi2"] == b["chi2"] and a["depth_positive"] == b["depth_positive"]
        n_st += len(a["err"]) == 3
    assert 2000 < n_st < 9000


@needs_ref
def test_exp_map_casts_and_huber_bit_equal_to_g2o():
    """VertexSE3Expmap::oplusImpl = SE3Quat::exp(update) * estimate incl. the theta < 1e-5 branch (se3quat.h:223-257, 280-285),
    SE3Quat::map, Converter::toSE3Quat / toCvMat (Converter.cc:41-75), RobustKernelHuber::robustify with its float dsqr member
    (robust_kernel_impl.h:84, .cpp:65-91)"""
    from oracle import oracle_py as O
    rng = np.random.default_rng(1)
    small = 0
    for i in range(10000):
        pose = synth.pose_problem(i % 200, n=1)["pose"]
        u = rng.normal(0, [1e-7, 1e-3, 0.1, 1.0][i % 4], 6)
        small += np.linalg.norm(u[:3]) < 1e-5
        assert np.array_equal(O.se3_oplus(pose, u), O.se3_oplus(pose, u, ref=True))
        X = rng.uniform(-5, 5, 3)
        assert np.array_equal(O.se3_map(pose, X), O.se3_map(pose, X, ref=True))
        T = O.to_cvmat(pose)
        assert np.array_equal(T, O.to_cvmat(pose, ref=True))
        assert np.array_equal(O.to_se3quat(T), O.to_se3quat(T, ref=True))
        e2 = rng.uniform(0, 30)
        for d in (np.float32(np.sqrt(5.991)), np.float32(np.sqrt(7.815))):
            assert np.array_equal(O.huber(e2, d), O.huber(e2, d, ref=True))
    assert small > 1000
    # rotations whose trace is <= 0 take the other branches of Quaterniond(Matrix3d)
    for ax in range(3):
        for ang in (np.pi, 3.0, -3.1, 2.5):
            T = np.eye(4, dtype=np.float32)
            T[:3, :3] = synth._rot(ax, ang).astype(np.float32)
            assert np.array_equal(O.to_se3quat(T), O.to_se3quat(T, ref=True))
    # the float dsqr is visible: rho[0] differs from the double-precision formula
    d = float(np.float32(np.sqrt(5.991)))
    assert O.huber(20.0, d)[0] != 2 * np.sqrt(20.0) * d - d * d


LBA_CASES = [  # n_kf, n_pts, stereo, first keyframe id, centre keyframe, obs per point
    (6, 300, False, 1, 0, 4), (8, 500, True, 1, 2, 4), (20, 3000, False, 1, 0, 4), (20, 3000, True, 5, 7, 4), (6, 300, False, 0, 3, 4),
    (10, 800, True, 0, 1, 3), (12, 600, False, 3, 11, 5)]


@needs_ref
@pytest.mark.parametrize("case", range(len(LBA_CASES)))
def test_local_bundle_adjustment_equals_the_reference(case):
    """the reference's Optimizer::LocalBundleAdjustment (Optimizer.cc:454-779) + g2o (BlockSolver_6_3 with Schur complement,
    LinearSolverEigen, Levenberg, two rounds) against oracle/lba_oracle.c on the problem that graph contains"""
    from oracle import oracle_py as O
    nk, npnt, st, fid, center, opp = LBA_CASES[case]
    p = synth.lba_problem(case, n_kf=nk, n_pts=npnt, stereo=st, obs_per_pt=opp)
    g = graph_from_problem(p, center=center, first_id=fid, to_cvmat=O.to_cvmat)
    r = O.ref_local_ba(g)
    assert (r["kf_role"] == 1).sum() >= nk // 2
    prob, kfs, pts_local, e_kp = problem_from_graph(g, r["kf_role"], O.to_se3quat)
    if fid == 0:
        assert prob["kf_fixed"][kfs.index(0)] == 1            # vSE3->setFixed(pKFi->mnId == 0), Optimizer.cc:530
    o = O.lba_solve(prob)
    local = [i for i, k in enumerate(kfs) if r["kf_role"][k] == 1]
    T_or = np.stack([O.to_cvmat(o["kf"][i]) for i in local])
    T_ref = r["Tcw"][[kfs[i] for i in local]]
    upd = np.abs(T_ref - g["kf_Tcw"][[kfs[i] for i in local]]).max()
    assert upd > 1e-3                                          # the window moved
    assert np.abs(T_or - T_ref).max() <= 2e-7                  # float 4x4 entries: equal to float rounding
    # points: BASELINE.json's tolerance is 1e-4 relative on the update; the two agree 10x tighter (an ill-conditioned point seen
    # under a small baseline moves by a few float ulps more)
    p0 = g["pts"][pts_local].astype(np.float64)
    rel = np.linalg.norm(o["pts"] - r["pts"][pts_local]) / np.linalg.norm(r["pts"][pts_local] - p0)
    assert rel <= 1e-5 and np.abs(o["pts"].astype(np.float32) - r["pts"][pts_local]).max() <= 2e-5, rel
    kept, bad = expected_kept(g, prob, e_kp, o["erase"])
    assert o["erase"].sum() > 10
    assert np.array_equal(kept, r["kp_kept"]) and np.array_equal(bad, r["pt_bad"] == 1)      # pt_bad == 2: a point no keyframe observes
    # fixed keyframes and keyframes outside the window keep their pose bit for bit
    others = [k for k in range(nk) if r["kf_role"][k] != 1]
    assert np.array_equal(r["Tcw"][others], g["kf_Tcw"][others])


@needs_ref
def test_local_bundle_adjustment_stop_flag_and_keyframe_zero():
    from oracle import oracle_py as O
    p = synth.lba_problem(3, n_kf=6, n_pts=300)
    g = graph_from_problem(p, center=0, first_id=1, to_cvmat=O.to_cvmat)
    g["stop_before"] = 1                                       # Optimizer.cc:656-658: return before optimising
    r = O.ref_local_ba(g)
    assert np.array_equal(r["Tcw"], g["kf_Tcw"]) and np.array_equal(r["pts"], g["pts"]) and r["kp_kept"].all()
    # pKF->mnId == 0: every map point already carries mnBALocalForKF == 0, so the reference optimises an empty graph
    g = graph_from_problem(p, center=0, first_id=0, to_cvmat=O.to_cvmat)
    r = O.ref_local_ba(g)
    assert np.array_equal(r["pts"], g["pts"]) and r["kp_kept"].all() and np.abs(r["Tcw"] - g["kf_Tcw"]).max() < 1e-6


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_pose_optimization_equals_the_reference(seed):
    """the reference's Optimizer::PoseOptimization (Optimizer.cc:239-452) + g2o (unary edges, LinearSolverDense, 4 x optimize(10))
    against oracle/pose_oracle.c: same outlier flags, return value and nBadPoseOpt, float pose identical"""
    from oracle import oracle_py as O
    inv = inv_sigma2_table()
    n = [400, 50, 8, 1000, 5, 2][seed % 6]
    p = synth.pose_problem(seed, n=n, stereo_frac=[0.6, 0.0, 1.0][seed % 3])
    Tcw = O.to_cvmat(p["pose"])
    octave = np.array([int(np.argmin(np.abs(inv - v))) for v in p["inv_sigma2"]], np.int32)
    m = int(n * 1.3) + 1                                       # the frame also has keypoints without a map point
    rng = np.random.default_rng(seed)
    sel = np.sort(rng.choice(m, n, replace=False))
    has = np.zeros(m, np.uint8); has[sel] = 1
    kp = np.zeros((m, 3), np.float32); kp[:, 2] = -1; kp[sel] = p["obs"].astype(np.float32)
    oc = np.zeros(m, np.int32); oc[sel] = octave
    Xw = np.zeros((m, 3), np.float32); Xw[sel] = p["Xw"].astype(np.float32)
    r = O.ref_pose_optimization(dict(kp_xy_ur=kp, kp_octave=oc, Xw=Xw, has_point=has, Tcw=Tcw, K=p["K"]))
    q = dict(p); q["pose"] = O.to_se3quat(Tcw)
    o = O.pose_optimize(q)
    assert r["ret"] == o["n_inliers"] and r["n_bad"] == o["n_bad"]
    assert np.array_equal(r["outlier"][sel], o["outlier"]) and not r["outlier"][has == 0].any()
    if n >= 3:
        assert np.abs(O.to_cvmat(o["pose"]) - r["Tcw"]).max() <= 1e-7
    else:
        assert np.array_equal(r["Tcw"], Tcw)                   # fewer than 3 correspondences: returns before SetPose (:355-356)

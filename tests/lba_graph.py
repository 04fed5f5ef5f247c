"""Helpers shared by tests/test_oracle_ref_optimizer.py (CPU) and tests/test_adapter_optimizer_gpu.py (GPU): turn a synthetic local-BA
problem (orbx.synth.lba_problem) into the keyframe / keypoint / map point arrays from which oracle/optimizer_ref_shim.cpp builds the
reference's KeyFrame / MapPoint / Map objects, and back into the POD problem that Optimizer.cc:456-655 builds from such a graph."""
import numpy as np

from orbx import synth


def inv_sigma2_table(nlevels=8):
    sf = synth.scale_factors(nlevels)
    return (np.float32(1.0) / (sf * sf)).astype(np.float32)


def graph_from_problem(p, center=0, first_id=1, to_cvmat=None):
    n_kf = len(p["kf_pose"])
    Tcw = np.stack([to_cvmat(p["kf_pose"][k]) for k in range(n_kf)])
    order = np.argsort(p["e_kf"], kind="stable")
    kf_start = np.zeros(n_kf + 1, np.int64)
    np.add.at(kf_start, np.asarray(p["e_kf"]) + 1, 1)
    kf_start = np.cumsum(kf_start).astype(np.int32)
    obs = p["e_obs"][order].astype(np.float32).copy()
    obs[p["e_stereo"][order] == 0, 2] = -1.0
    inv = inv_sigma2_table()
    octave = np.array([int(np.argmin(np.abs(inv - v))) for v in p["e_inv_sigma2"][order]], np.int32)
    return dict(kf_Tcw=Tcw, kf_start=kf_start, kp_xy_ur=obs, kp_octave=octave, kp_point=p["e_pt"][order].astype(np.int32),
                pts=p["pts"].astype(np.float32), K=p["K"], center_kf=center, first_kf_id=first_id)


def problem_from_graph(g, role, to_se3quat):
    """role[k]: 1 = local keyframe, 2 = fixed keyframe, 0 = outside the window (as the reference marked them).
    -> (problem dict, keyframe indices, point indices, keypoint row of every edge)"""
    first_id = g["first_kf_id"]
    n_kf = len(g["kf_Tcw"])
    kfs = [k for k in range(n_kf) if role[k]]
    idx = {k: i for i, k in enumerate(kfs)}
    local = [k for k in kfs if role[k] == 1]
    pts_local = sorted(set(int(g["kp_point"][j]) for k in local for j in range(g["kf_start"][k], g["kf_start"][k + 1]) if g["kp_point"][j] >= 0))
    pidx = {q: i for i, q in enumerate(pts_local)}
    inv = inv_sigma2_table()
    e_kf, e_pt, e_obs, e_is2, e_st, e_kp = [], [], [], [], [], []
    for k in kfs:
        for j in range(g["kf_start"][k], g["kf_start"][k + 1]):
            q = int(g["kp_point"][j])
            if q in pidx:
                e_kf.append(idx[k]); e_pt.append(pidx[q]); e_obs.append(g["kp_xy_ur"][j].astype(np.float64))
                e_is2.append(inv[g["kp_octave"][j]]); e_st.append(0 if g["kp_xy_ur"][j][2] < 0 else 1); e_kp.append(j)
    kf_pose = np.stack([to_se3quat(g["kf_Tcw"][k]) for k in kfs]) if kfs else np.zeros((0, 7))
    fixed = np.array([1 if (role[k] == 2 or k + first_id == 0) else 0 for k in kfs], np.uint8)
    prob = dict(kf_pose=kf_pose, kf_fixed=fixed, pts=g["pts"][pts_local].astype(np.float64), e_kf=np.array(e_kf, np.int32),
                e_pt=np.array(e_pt, np.int32), e_obs=np.array(e_obs, np.float64).reshape(-1, 3), e_inv_sigma2=np.array(e_is2, np.float32),
                e_stereo=np.array(e_st, np.uint8), K=g["K"])
    return prob, kfs, pts_local, np.array(e_kp, np.int64)


def expected_kept(g, prob, e_kp, erase):
    """What Optimizer.cc:745-756 + MapPoint::EraseObservation (MapPoint.cc:117-147) leave of the graph for a given erase list: an
    erased observation is removed; a point left with <= 2 observations (a stereo observation counts twice) goes bad and loses all."""
    kept = (g["kp_point"] >= 0).astype(np.uint8)
    n_obs = np.zeros(len(g["pts"]), np.int64)
    w = np.where(g["kp_xy_ur"][:, 2] >= 0, 2, 1)
    np.add.at(n_obs, g["kp_point"][g["kp_point"] >= 0], w[g["kp_point"] >= 0])
    bad = np.zeros(len(g["pts"]), np.uint8)
    for e in np.nonzero(erase)[0]:
        j = e_kp[e]
        q = g["kp_point"][j]
        if bad[q]:
            continue
        kept[j] = 0
        n_obs[q] -= w[j]
        if n_obs[q] <= 2:
            bad[q] = 1
            kept[g["kp_point"] == q] = 0
    return kept, bad

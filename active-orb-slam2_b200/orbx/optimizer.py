"""Host-side mirror of Optimizer::LocalBundleAdjustment (reference include/Optimizer.h:45, src/Optimizer.cc:454-779)
over the orbx C ABI.  The problem is the POD form of the g2o graph the reference builds (Optimizer.cc:486-655):
a dict with kf_pose (n x 7: quaternion x,y,z,w + translation), kf_fixed, pts (n x 3), e_kf, e_pt, e_obs (n x 3),
e_inv_sigma2, e_stereo, K = (fx, fy, cx, cy, bf) and optionally stop_flag (uint8[1], the reference's pbStopFlag)."""
import ctypes as C

import numpy as np

from ._lib import check, lib


class LbaProblem(C.Structure):
    """orbx_lba_problem (include/orbx.h)"""
    _fields_ = [("n_kf", C.c_int32), ("kf_pose", C.c_void_p), ("kf_fixed", C.c_void_p), ("n_pts", C.c_int32), ("pts", C.c_void_p),
                ("n_edges", C.c_int32), ("e_kf", C.c_void_p), ("e_pt", C.c_void_p), ("e_obs", C.c_void_p),
                ("e_inv_sigma2", C.c_void_p), ("e_stereo", C.c_void_p),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double),
                ("stop_flag", C.c_void_p)]


class LbaResult(C.Structure):
    """orbx_lba_result (include/orbx.h)"""
    _fields_ = [("kf_pose", C.c_void_p), ("pts", C.c_void_p), ("chi2", C.c_void_p), ("erase", C.c_void_p), ("lm_trials", C.c_int32),
                ("stopped", C.c_int32), ("first_Hschur", C.c_void_p), ("first_bschur", C.c_void_p), ("first_xp", C.c_void_p),
                ("first_lambda", C.c_double)]


def pack_problem(prob):
    keep = dict(kf_pose=np.ascontiguousarray(prob["kf_pose"], np.float64), kf_fixed=np.ascontiguousarray(prob["kf_fixed"], np.uint8),
                pts=np.ascontiguousarray(prob["pts"], np.float64), e_kf=np.ascontiguousarray(prob["e_kf"], np.int32),
                e_pt=np.ascontiguousarray(prob["e_pt"], np.int32), e_obs=np.ascontiguousarray(prob["e_obs"], np.float64),
                e_inv_sigma2=np.ascontiguousarray(prob["e_inv_sigma2"], np.float32), e_stereo=np.ascontiguousarray(prob["e_stereo"], np.uint8))
    P = LbaProblem()
    P.n_kf, P.n_pts, P.n_edges = len(keep["kf_pose"]), len(keep["pts"]), len(keep["e_kf"])
    for k, v in keep.items():
        setattr(P, k, v.ctypes.data)
    P.fx, P.fy, P.cx, P.cy, P.bf = prob["K"]
    if prob.get("stop_flag") is not None:
        keep["stop_flag"] = prob["stop_flag"]
        P.stop_flag = prob["stop_flag"].ctypes.data
    return P, keep


class Optimizer:
    """holds the device buffers; LocalBundleAdjustment() is the reference's static method"""

    def __init__(self, max_keyframes=64, max_points=8192, max_edges=65536, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.orbx_lba_create(C.byref(self._h), max_keyframes, max_points, max_edges, device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_lba_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def LocalBundleAdjustment(self, prob, its1=5, its2=10, want_system=False):
        """-> dict(kf, pts, chi2, erase, trials, stopped[, Hschur, bschur, xp, lambda0]); prob: the problem dict, or what
        pack_problem(prob) returned (a caller that times the C call packs once)"""
        P, keep = prob if isinstance(prob, tuple) else pack_problem(prob)
        kf, pt = np.zeros((P.n_kf, 7)), np.zeros((P.n_pts, 3))
        chi2, erase = np.zeros(max(P.n_edges, 1)), np.zeros(max(P.n_edges, 1), np.uint8)
        R = LbaResult()
        R.kf_pose, R.pts, R.chi2, R.erase = kf.ctypes.data, pt.ctypes.data, chi2.ctypes.data, erase.ctypes.data
        dim = 6 * int((keep["kf_fixed"] == 0).sum())
        if want_system:
            Hs, bs, xp = np.zeros((dim, dim)), np.zeros(dim), np.zeros(dim)
            R.first_Hschur, R.first_bschur, R.first_xp = Hs.ctypes.data, bs.ctypes.data, xp.ctypes.data
        check(self._L.orbx_lba_solve_host(self._h, C.byref(P), its1, its2, C.byref(R)))
        out = dict(kf=kf, pts=pt, chi2=chi2[:P.n_edges], erase=erase[:P.n_edges], trials=R.lm_trials, stopped=R.stopped)
        if want_system:
            out.update(Hschur=Hs, bschur=bs, xp=xp, lambda0=R.first_lambda)
        return out

    def prepare(self, prob):
        """argument and result structs of orbx_lba_solve_host with their arrays, built once (a caller that times the C call alone)"""
        P, keep = prob if isinstance(prob, tuple) else pack_problem(prob)
        arrays = (np.zeros((P.n_kf, 7)), np.zeros((P.n_pts, 3)), np.zeros(max(P.n_edges, 1)), np.zeros(max(P.n_edges, 1), np.uint8))
        R = LbaResult()
        R.kf_pose, R.pts, R.chi2, R.erase = (a.ctypes.data for a in arrays)
        return P, keep, R, arrays

    def call(self, prepared, its1=5, its2=10):
        """orbx_lba_solve_host on prepare()'s structs; the results stay in its arrays -> Levenberg trials"""
        check(self._L.orbx_lba_solve_host(self._h, C.byref(prepared[0]), its1, its2, C.byref(prepared[2])))
        return prepared[2].lm_trials

    def begin(self, prob, its1=5, its2=10):
        """asynchronous LocalBundleAdjustment: enqueue the whole window on this handle's stream and return.  prob: the problem dict,
        or the (struct, keepalive) pair pack_problem(prob) returned (a caller that submits the same arrays again packs once)"""
        self._pending = prob if isinstance(prob, tuple) else pack_problem(prob)
        check(self._L.orbx_lba_solve_begin(self._h, C.byref(self._pending[0]), its1, its2))

    def end(self):
        """wait for begin() and return the same dict as LocalBundleAdjustment()"""
        P, keep = self._pending
        if getattr(self, "_out", None) is None or self._out[0].shape[0] != P.n_kf or self._out[1].shape[0] != P.n_pts or len(self._out[2]) != max(P.n_edges, 1):
            self._out = (np.zeros((P.n_kf, 7)), np.zeros((P.n_pts, 3)), np.zeros(max(P.n_edges, 1)), np.zeros(max(P.n_edges, 1), np.uint8))
        kf, pt, chi2, erase = self._out                  # reused between calls of the same shape: the caller copies what it keeps
        R = LbaResult()
        R.kf_pose, R.pts, R.chi2, R.erase = kf.ctypes.data, pt.ctypes.data, chi2.ctypes.data, erase.ctypes.data
        check(self._L.orbx_lba_solve_end(self._h, C.byref(P), C.byref(R)))
        self._pending = None
        return dict(kf=kf, pts=pt, chi2=chi2[:P.n_edges], erase=erase[:P.n_edges], trials=R.lm_trials, stopped=R.stopped)

    def build_schur_timed(self, prob, lam, reps=1, want_system=False):
        """one Levenberg trial's system build (residuals + Jacobians + quadratic form + Schur complement), `reps` times;
        -> (milliseconds for all reps, Hschur, bschur)"""
        P, keep = (pack_problem(prob) if prob is not None else (None, None))
        ms = C.c_float()
        Hs = bs = None
        dim = 6 * int((keep["kf_fixed"] == 0).sum()) if keep else 0
        if want_system:
            Hs, bs = np.zeros((dim, dim)), np.zeros(dim)
        check(self._L.orbx_lba_build_schur_timed(self._h, C.byref(P) if P is not None else None, lam, reps, C.byref(ms),
                                                 Hs.ctypes.data if want_system else None, bs.ctypes.data if want_system else None))
        return ms.value, Hs, bs

    def last_launches(self):
        return self._L.orbx_lba_last_launches(self._h)

    def phase_us(self):
        """microseconds per phase of the cluster kernel in the last LocalBundleAdjustment call"""
        ns = np.zeros(6)
        check(self._L.orbx_lba_phase_ns(self._h, ns.ctypes.data))
        return dict(zip(("build", "schur", "reduce", "solve", "update", "err"), (ns / 1e3).tolist()))


# ---- Optimizer::PoseOptimization (reference include/Optimizer.h:49, src/Optimizer.cc:239-452) ----------------------
class PoseProblem(C.Structure):
    """orbx_pose_problem (include/orbx.h)"""
    _fields_ = [("n", C.c_int32), ("Xw", C.c_void_p), ("obs", C.c_void_p), ("inv_sigma2", C.c_void_p), ("pose", C.c_double * 7),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bf", C.c_double)]


class PoseResult(C.Structure):
    """orbx_pose_result (include/orbx.h)"""
    _fields_ = [("pose", C.c_double * 7), ("outlier", C.c_void_p), ("n_inliers", C.c_int32), ("n_bad", C.c_int32),
                ("lm_trials", C.c_int32)]


class PoseOptimizer:
    """holds the device buffers; PoseOptimization() is the reference's static method, batched over frames"""

    def __init__(self, max_observations=65536, max_frames=64, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.orbx_pose_create(C.byref(self._h), max_observations, max_frames, device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_pose_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def pack(self, probs):
        """the argument arrays of orbx_pose_optimize_host for a list of problem dicts (reusable: a caller that times the C call alone
        packs once)"""
        plist = list(probs)
        nf = len(plist)
        P, R = (PoseProblem * max(nf, 1))(), (PoseResult * max(nf, 1))()
        keep, outs = [], []
        for f, p in enumerate(plist):
            Xw = np.ascontiguousarray(p["Xw"], np.float64).reshape(-1, 3)
            obs = np.ascontiguousarray(p["obs"], np.float64).reshape(-1, 3)
            s2 = np.ascontiguousarray(p["inv_sigma2"], np.float32)
            out = np.zeros(max(len(Xw), 1), np.uint8)
            keep += [Xw, obs, s2]
            outs.append(out)
            P[f].n, P[f].Xw, P[f].obs, P[f].inv_sigma2 = len(Xw), Xw.ctypes.data, obs.ctypes.data, s2.ctypes.data
            for i, v in enumerate(np.asarray(p["pose"], np.float64)):
                P[f].pose[i] = v
            P[f].fx, P[f].fy, P[f].cx, P[f].cy, P[f].bf = (float(v) for v in p["K"][:5])
            R[f].outlier = out.ctypes.data
        return P, R, nf, keep, outs

    def run(self, packed):
        """orbx_pose_optimize_host on pack()'s arrays -> dict(pose, outlier, n_inliers, n_bad, trials) per problem"""
        P, R, nf, _, outs = packed
        check(self._L.orbx_pose_optimize_host(self._h, P, nf, R))
        return [dict(pose=np.array(list(R[f].pose)), outlier=outs[f][:P[f].n].copy(), n_inliers=R[f].n_inliers, n_bad=R[f].n_bad,
                     trials=R[f].lm_trials) for f in range(nf)]

    def call(self, packed):
        """the C call alone (results stay in pack()'s arrays)"""
        check(self._L.orbx_pose_optimize_host(self._h, packed[0], packed[2], packed[1]))

    def PoseOptimization(self, probs):
        """probs: one dict or a list of dicts(Xw[n,3], obs[n,3] (third < 0 = monocular), inv_sigma2[n], pose[7], K=(fx,fy,cx,cy,bf))
        -> dict(pose, outlier, n_inliers, n_bad, trials) per problem"""
        single = isinstance(probs, dict)
        res = self.run(self.pack([probs] if single else probs))
        return res[0] if single else res

    def from_matches_device(self, d_jobs, n_frames, d_inv_sigma2, nlevels, K, d_pose_out, d_n_inliers, d_outlier_kp, kp_pitch, stream=0):
        """PoseOptimization of a batch of frames from the device job array of ORBmatcher.search_frames_device (raw device pointers)"""
        fx, fy, cx, cy, bf = (float(v) for v in K[:5])
        check(self._L.orbx_pose_from_matches_device(self._h, d_jobs, n_frames, d_inv_sigma2, nlevels, fx, fy, cx, cy, bf, d_pose_out,
                                                    d_n_inliers, d_outlier_kp, kp_pitch, stream))

    def last_launches(self):
        return self._L.orbx_pose_last_launches(self._h)

"""Oracle of Optimizer::PoseOptimization (oracle/pose_oracle.c, reference src/Optimizer.cc:239-452) against an independent
numpy statement of the same algorithm: numeric (central-difference) Jacobians, numpy.linalg.solve for the 6x6 system, the same
Levenberg schedule (g2o core/optimization_algorithm_levenberg.cpp:61-189) and the same four-round outlier policy.  CPU only."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import oracle_py as O
from orbx import synth
from test_lba_oracle import se3_exp_apply


def project(R, t, X, K, stereo, smooth=False):
    fx, fy, cx, cy, bf = (float(v) for v in K)
    Xc = R @ X + t
    if not stereo:
        return np.array([Xc[0] / Xc[2] * fx + cx, Xc[1] / Xc[2] * fy + cy])
    iz = 1.0 / Xc[2] if smooth else float(np.float32(1.0 / Xc[2]))
    u = Xc[0] * iz * fx + cx
    return np.array([u, Xc[1] * iz * fy + cy, u - bf * iz])


class NumpyPose:
    def __init__(self, p):
        self.p = p
        self.n = len(p["Xw"])
        self.stereo = ~(p["obs"][:, 2] < 0)
        self.level1 = np.zeros(self.n, bool)
        self.stored = np.zeros(self.n)
        self.robust = True
        self.trials = 0

    def reset(self):
        self.R = Rotation.from_quat(self.p["pose"][:4]).as_matrix()
        self.t = self.p["pose"][4:].copy()

    def err(self, e, R=None, t=None, smooth=False):
        R = self.R if R is None else R
        t = self.t if t is None else t
        z = project(R, t, self.p["Xw"][e], self.p["K"], self.stereo[e], smooth)
        return self.p["obs"][e][:len(z)] - z

    def delta(self, e):
        return float(np.float32(np.sqrt(7.815 if self.stereo[e] else 5.991)))

    def chi(self):
        tot = 0.0
        for e in np.nonzero(~self.level1)[0]:
            r = self.err(e)
            c = float(self.p["inv_sigma2"][e]) * float(r @ r)
            self.stored[e] = c
            if self.robust:
                d = self.delta(e)
                if c > d * d:
                    c = 2 * np.sqrt(c) * d - d * d
            tot += c
        return tot

    def optimize(self, its):
        act = np.nonzero(~self.level1)[0]
        if len(act) == 0:
            return
        lam, ni, nbad, h = 0.0, 2.0, 0, 1e-6
        for it in range(its):
            cur = self.chi()
            ini = cur
            H, b = np.zeros((6, 6)), np.zeros(6)
            q0 = Rotation.from_matrix(self.R).as_quat()
            for e in act:
                r = self.err(e)
                J = np.zeros((len(r), 6))
                for a in range(6):
                    d = np.zeros(6); d[a] = h
                    Rp, tp = se3_exp_apply(d, q0, self.t); Rm, tm = se3_exp_apply(-d, q0, self.t)
                    J[:, a] = (self.err(e, Rp, tp, True) - self.err(e, Rm, tm, True)) / (2 * h)
                info = float(self.p["inv_sigma2"][e])
                rho1 = 1.0
                if self.robust:
                    dl, c = self.delta(e), info * float(r @ r)
                    if c > dl * dl:
                        rho1 = dl / np.sqrt(c)
                H += J.T @ J * (rho1 * info)
                b -= rho1 * (J.T @ (info * r))
            if it == 0:
                lam, ni, nbad = 1e-5 * np.abs(np.diag(H)).max(), 2.0, 0
            rho, q = 0.0, 0
            while True:
                bakR, bakt = self.R.copy(), self.t.copy()
                x = np.linalg.solve(H + lam * np.eye(6), b)
                self.R, self.t = se3_exp_apply(x, Rotation.from_matrix(self.R).as_quat(), self.t)
                tmp = self.chi()
                self.trials += 1
                rho = (cur - tmp) / (float(x @ (lam * x + b)) + 1e-3)
                if rho > 0 and np.isfinite(tmp):
                    lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3)); ni = 2.0; cur = tmp
                else:
                    lam *= ni; ni *= 2; self.R, self.t = bakR, bakt
                q += 1
                if not (rho < 0 and q < 10):
                    break
            if q == 10 or rho == 0:
                break
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            if nbad >= 3:
                break

    def run(self):
        outlier = np.zeros(self.n, bool)
        nbad = 0
        if self.n < 3:
            return self.p["pose"].copy(), outlier, 0
        for it in range(4):
            self.reset()
            self.optimize(10)
            nbad = 0
            for e in range(self.n):
                if outlier[e]:
                    r = self.err(e)
                    self.stored[e] = float(self.p["inv_sigma2"][e]) * float(r @ r)
                bad = np.float32(self.stored[e]) > np.float32(7.815 if self.stereo[e] else 5.991)
                outlier[e] = self.level1[e] = bad
                nbad += int(bad)
            if it == 2:
                self.robust = False
            if self.n < 10:
                break
        pose = np.concatenate([Rotation.from_matrix(self.R).as_quat(), self.t])
        if pose[3] < 0:
            pose[:4] = -pose[:4]
        return pose, outlier, nbad


def pose_delta(a, b):
    """rotation angle (rad) and translation distance between two (quat, t) poses"""
    ra, rb = Rotation.from_quat(a[:4]), Rotation.from_quat(b[:4])
    return (ra.inv() * rb).magnitude(), np.linalg.norm(a[4:] - b[4:])


@pytest.mark.parametrize("seed,n,stereo_frac", [(0, 120, 0.6), (1, 200, 0.0), (2, 150, 1.0), (3, 60, 0.5)])
def test_oracle_equals_numpy(seed, n, stereo_frac):
    p = synth.pose_problem(seed, n=n, stereo_frac=stereo_frac)
    r = O.pose_optimize(p)
    ref_pose, ref_out, ref_bad = NumpyPose(p).run()
    assert np.array_equal(r["outlier"].astype(bool), ref_out)
    assert r["n_bad"] == ref_bad and r["n_inliers"] == n - ref_bad
    moved_r, moved_t = pose_delta(p["pose"], ref_pose)
    dr, dt = pose_delta(r["pose"], ref_pose)
    assert moved_t > 1e-3 and dr < 1e-6 * max(moved_r, 1e-3) + 1e-9 and dt < 1e-6 * moved_t + 1e-9, (dr, dt, moved_r, moved_t)
    assert 0.05 * n < ref_bad < 0.4 * n          # the planted outliers are found


def test_recovers_the_true_pose():
    p = synth.pose_problem(7, n=400, outlier_frac=0.1)
    r = O.pose_optimize(p)
    # reproject the inliers with the optimised pose: residuals at the noise level
    R, t = Rotation.from_quat(r["pose"][:4]).as_matrix(), r["pose"][4:]
    inl = ~r["outlier"].astype(bool)
    res = [np.linalg.norm((p["obs"][e][:2] - project(R, t, p["Xw"][e], p["K"], False)) * np.sqrt(p["inv_sigma2"][e])) for e in np.nonzero(inl)[0]]
    assert np.median(res) < 1.5


def test_small_graphs():
    p = synth.pose_problem(5, n=2)
    r = O.pose_optimize(p)                       # fewer than 3 correspondences: returns 0, pose untouched (:355)
    assert r["n_inliers"] == 0 and np.array_equal(r["pose"], p["pose"]) and r["trials"] == 0
    p = synth.pose_problem(6, n=8, outlier_frac=0.0)
    r = O.pose_optimize(p)                       # fewer than 10 edges: one round only (:419)
    ref_pose, ref_out, ref_bad = NumpyPose(p).run()
    assert np.array_equal(r["outlier"].astype(bool), ref_out) and 0 < r["trials"] <= 100
    dr, dt = pose_delta(r["pose"], ref_pose)
    assert dr < 1e-8 and dt < 1e-8

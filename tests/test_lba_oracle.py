"""Oracle local BA (oracle/lba_oracle.c) against an independent numpy statement of the same algorithm: numeric
(central-difference) Jacobians, the full un-reduced normal equations solved densely, the same Levenberg schedule
(g2o core/optimization_algorithm_levenberg.cpp:61-189) and the same two-round outlier policy (Optimizer.cc:659-735).
CPU only."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import oracle_py as O
from orbx import synth


def se3_exp_apply(delta, q, t):
    """exp(delta) * (q, t), delta = (omega, upsilon) as in g2o SE3Quat::exp"""
    w, u = delta[:3], delta[3:]
    th = np.linalg.norm(w)
    Om = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-5:
        R = np.eye(3) + Om + Om @ Om
        V = R
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th ** 2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    R0 = Rotation.from_quat(q).as_matrix()
    return R @ R0, V @ u + R @ t


def project(R, t, X, K, stereo, smooth=False):
    """smooth=True drops g2o's float rounding of 1/z (used only to differentiate numerically)"""
    fx, fy, cx, cy, bf = K
    Xc = R @ X + t
    if not stereo:
        return np.array([Xc[0] / Xc[2] * fx + cx, Xc[1] / Xc[2] * fy + cy]), Xc[2]
    if smooth:
        u = Xc[0] / Xc[2] * fx + cx
        return np.array([u, Xc[1] / Xc[2] * fy + cy, u - bf / Xc[2]]), Xc[2]
    iz = np.float32(1.0 / Xc[2])      # `1.0f/trans_xyz[2]`: double division narrowed once
    u = Xc[0] * iz * fx + cx
    return np.array([u, Xc[1] * iz * fy + cy, u - float(np.float32(bf) * iz)]), Xc[2]


class NumpyLBA:
    def __init__(self, p):
        self.p = p
        self.R = [Rotation.from_quat(q[:4]).as_matrix() for q in p["kf_pose"]]
        self.t = [q[4:].copy() for q in p["kf_pose"]]
        self.X = p["pts"].copy()
        self.level1 = np.zeros(len(p["e_kf"]), bool)
        self.robust = True
        self.stored = np.zeros(len(p["e_kf"]))

    def err(self, e, R=None, t=None, X=None, smooth=False):
        p = self.p
        k, l, st = p["e_kf"][e], p["e_pt"][e], bool(p["e_stereo"][e])
        R = self.R[k] if R is None else R
        t = self.t[k] if t is None else t
        X = self.X[l] if X is None else X
        z, depth = project(R, t, X, p["K"], st, smooth)
        return p["e_obs"][e][:len(z)] - z, depth

    def chi(self, store=True):
        tot = 0.0
        for e in np.nonzero(~self.level1)[0]:
            r, _ = self.err(e)
            c = float(self.p["e_inv_sigma2"][e]) * float(r @ r)
            if store:
                self.stored[e] = c
            if self.robust:
                d = float(np.float32(np.sqrt(7.815 if self.p["e_stereo"][e] else 5.991)))
                if c > d * d:
                    c = 2 * np.sqrt(c) * d - d * d
            tot += c
        return tot

    def optimize(self, its):
        p = self.p
        act = np.nonzero(~self.level1)[0]
        kfs = [k for k in sorted(set(p["e_kf"][act])) if not p["kf_fixed"][k]]
        pts = sorted(set(p["e_pt"][act]))
        ki = {k: i for i, k in enumerate(kfs)}
        li = {l: i for i, l in enumerate(pts)}
        n = 6 * len(kfs) + 3 * len(pts)
        lam, ni, nbad = 0.0, 2.0, 0
        h = 1e-6
        for it in range(its):
            cur = self.chi()
            ini = cur
            H, b = np.zeros((n, n)), np.zeros(n)
            for e in act:
                k, l = p["e_kf"][e], p["e_pt"][e]
                r, _ = self.err(e)
                D = len(r)
                B = np.zeros((D, 6)); A = np.zeros((D, 3))
                q0 = Rotation.from_matrix(self.R[k]).as_quat()
                for a in range(6):
                    d = np.zeros(6); d[a] = h
                    Rp, tp = se3_exp_apply(d, q0, self.t[k]); Rm, tm = se3_exp_apply(-d, q0, self.t[k])
                    B[:, a] = (self.err(e, Rp, tp, smooth=True)[0] - self.err(e, Rm, tm, smooth=True)[0]) / (2 * h)   # dE/dxi
                for a in range(3):
                    d = np.zeros(3); d[a] = h
                    A[:, a] = (self.err(e, X=self.X[l] + d, smooth=True)[0] - self.err(e, X=self.X[l] - d, smooth=True)[0]) / (2 * h)
                info = float(p["e_inv_sigma2"][e])
                rho1 = 1.0
                if self.robust:
                    dl = float(np.float32(np.sqrt(7.815 if p["e_stereo"][e] else 5.991)))
                    c = info * float(r @ r)
                    if c > dl * dl:
                        rho1 = dl / np.sqrt(c)
                w = rho1 * info
                sl = slice(6 * len(kfs) + 3 * li[l], 6 * len(kfs) + 3 * li[l] + 3)
                H[sl, sl] += A.T @ A * w
                b[sl] += A.T @ (-info * r) * rho1
                if k in ki:
                    sp = slice(6 * ki[k], 6 * ki[k] + 6)
                    H[sp, sp] += B.T @ B * w
                    b[sp] += B.T @ (-info * r) * rho1
                    H[sp, sl] += B.T @ A * w
                    H[sl, sp] += A.T @ B * w
            if it == 0:
                lam, ni, nbad = 1e-5 * np.abs(np.diag(H)).max(), 2.0, 0
            rho, q = 0.0, 0
            while True:
                bak = ([r.copy() for r in self.R], [t.copy() for t in self.t], self.X.copy())
                try:
                    x = np.linalg.solve(H + lam * np.eye(n), b)
                    np.linalg.cholesky(H + lam * np.eye(n))
                    ok = True
                except np.linalg.LinAlgError:
                    x, ok = np.zeros(n), False
                if ok:
                    for k, i in ki.items():
                        self.R[k], self.t[k] = se3_exp_apply(x[6 * i:6 * i + 6], Rotation.from_matrix(self.R[k]).as_quat(), self.t[k])
                    for l, i in li.items():
                        self.X[l] = self.X[l] + x[6 * len(kfs) + 3 * i:6 * len(kfs) + 3 * i + 3]
                tmp = self.chi() if ok else np.inf
                rho = (cur - tmp) / (float(x @ (lam * x + b)) + 1e-3)
                if rho > 0 and np.isfinite(tmp):
                    lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3)); ni = 2.0; cur = tmp
                else:
                    lam *= ni; ni *= 2
                    self.R, self.t, self.X = bak
                q += 1
                if not (rho < 0 and q < 10):
                    break
            if q == 10 or rho == 0:
                break
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            if nbad >= 3:
                break

    def run(self, its1, its2):
        self.optimize(its1)
        th = np.where(self.p["e_stereo"] > 0, 7.815, 5.991)
        depth = np.array([self.err(e)[1] for e in range(len(th))])
        self.level1 = (self.stored > th) | ~(depth > 0)
        self.robust = False
        self.optimize(its2)
        depth = np.array([self.err(e)[1] for e in range(len(th))])
        return (self.stored > th) | ~(depth > 0)


@pytest.mark.parametrize("stereo,n_fixed,seed", [(False, 1, 0), (True, 2, 1), (False, 0, 2)])
def test_oracle_matches_numpy_lm(stereo, n_fixed, seed):
    p = synth.lba_problem(seed, n_kf=5, n_pts=60, obs_per_pt=3, stereo=stereo, n_fixed=n_fixed)
    ref = O.lba_solve(p, 3, 4)
    m = NumpyLBA(p)
    erase = m.run(3, 4)
    q = np.array([Rotation.from_matrix(R).as_quat() for R in m.R])
    q *= np.sign(q[:, 3:4])
    assert np.allclose(ref["kf"][:, :4], q, atol=2e-6)
    assert np.allclose(ref["kf"][:, 4:], np.array(m.t), atol=2e-6)
    assert np.allclose(ref["pts"], m.X, rtol=1e-5, atol=1e-5)
    assert np.array_equal(ref["erase"].astype(bool), erase)


def test_fixed_keyframes_do_not_move_and_chi2_drops():
    p = synth.lba_problem(3, n_kf=8, n_pts=300, n_fixed=3)
    r = O.lba_solve(p)
    assert np.array_equal(r["kf"][:3], p["kf_pose"][:3])
    assert r["chi2_trace"][0] < 60000 and r["trials"] >= 5
    inl = r["erase"] == 0
    assert r["chi2"][inl].max() <= 5.991 + 1e-9


def test_stop_flag_before_start_returns_input():
    p = synth.lba_problem(4, n_kf=4, n_pts=50)
    p["stop_flag"] = np.ones(1, np.uint8)
    r = O.lba_solve(p)
    assert r["stopped"] == 1 and np.array_equal(r["kf"], p["kf_pose"]) and np.array_equal(r["pts"], p["pts"])

"""Oracle of Frame::ComputeStereoMatches (oracle/stereo_oracle.c, reference src/Frame.cc:495-669) against an independent
numpy / cv2 statement of the same rules: row table as Python lists, cv2.norm(IL, IR, NORM_L1) on CV_32F patches exactly as the
reference composes them, float32 scalar arithmetic for the parabola and the disparity gates."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth

cv2 = pytest.importorskip("cv2")
F = np.float32


def hamming(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def stereo_reference_py(kl, dl, kr, dr, pyr_l, pyr_r, scale, inv_scale, bf, b):
    """line-by-line Python transcription of the reference's control flow with cv2 doing what cv:: does there"""
    N = len(kl)
    u_right, depth = np.full(N, -1, F), np.full(N, -1, F)
    n_rows = pyr_l[0].shape[0]
    rows = [[] for _ in range(n_rows)]
    for iR in range(len(kr)):
        y = F(kr["y"][iR])
        r = F(2.0) * scale[kr["octave"][iR]]
        for yi in range(int(np.floor(y - r)), int(np.ceil(y + r)) + 1):
            rows[yi].append(iR)
    min_d, max_d = F(0), F(bf) / F(b)
    dist_idx = []
    for iL in range(N):
        lvl = int(kl["octave"][iL])
        uL, vL = F(kl["x"][iL]), F(kl["y"][iL])
        cand = rows[int(vL)]
        if not cand:
            continue
        min_u, max_u = uL - max_d, uL - min_d
        if max_u < 0:
            continue
        best, best_r = 100, 0
        for iR in cand:
            if kr["octave"][iR] < lvl - 1 or kr["octave"][iR] > lvl + 1:
                continue
            uR = F(kr["x"][iR])
            if min_u <= uR <= max_u:
                d = hamming(dl[iL], dr[iR])
                if d < best:
                    best, best_r = d, iR
        if best >= 75:
            continue
        sf = inv_scale[lvl]
        # round(): half away from zero (np.round is half-to-even, and products like 249 * (1/1.2) do land on .5)
        su, sv, sr0 = (int(np.floor(float(F(v) * sf) + 0.5)) for v in (kl["x"][iL], kl["y"][iL], kr["x"][best_r]))
        w = L = 5
        IL = pyr_l[lvl][sv - w:sv + w + 1, su - w:su + w + 1].astype(F)
        IL = IL - IL[w, w] * np.ones_like(IL)
        if sr0 + L - w < 0 or sr0 + L + w + 1 >= pyr_r[lvl].shape[1]:
            continue
        best_sad, best_inc, dists = 2 ** 31 - 1, 0, np.zeros(2 * L + 1, F)
        for inc in range(-L, L + 1):
            IR = pyr_r[lvl][sv - w:sv + w + 1, sr0 + inc - w:sr0 + inc + w + 1].astype(F)
            IR = IR - IR[w, w] * np.ones_like(IR)
            d = F(cv2.norm(IL, IR, cv2.NORM_L1))
            if d < best_sad:
                best_sad, best_inc = int(d), inc
            dists[L + inc] = d
        if best_inc in (-L, L):
            continue
        d1, d2, d3 = dists[L + best_inc - 1], dists[L + best_inc], dists[L + best_inc + 1]
        with np.errstate(all="ignore"):
            delta = (d1 - d3) / (F(2.0) * (d1 + d3 - F(2.0) * d2))
        if delta < -1 or delta > 1:
            continue
        best_ur = scale[lvl] * (F(sr0) + F(best_inc) + delta)
        disp = uL - best_ur
        if disp >= min_d and disp < max_d:
            if disp <= 0:
                disp = F(0.01)
                best_ur = F(np.float64(uL) - 0.01)
            depth[iL] = F(bf) / disp
            u_right[iL] = best_ur
            dist_idx.append((best_sad, iL))
    if dist_idx:
        dist_idx.sort()
        th = F(1.5) * F(1.4) * F(dist_idx[len(dist_idx) // 2][0])
        for d, i in reversed(dist_idx):
            if F(d) < th:
                break
            u_right[i] = depth[i] = -1
    return u_right, depth


def stereo_scene(seed, w=640, h=480, nfeat=1000, yaw=0.0):
    world = synth.stereo_world(seed, w, h)
    left, right = world.render(0.02 * seed, 0.01, yaw), world.render(0.02 * seed, 0.01, yaw, right=True)
    exl, exr = O.Extractor(nfeat, 1.2, 8, 20, 7), O.Extractor(nfeat, 1.2, 8, 20, 7)
    kl, dl = exl(left)
    kr, dr = exr(right)
    t = exl.tables()
    pl = [exl.level(l) for l in range(8)]
    pr = [exr.level(l) for l in range(8)]
    return dict(kl=kl, dl=dl, kr=kr, dr=dr, pl=pl, pr=pr, scale=t["scale"], inv_scale=t["inv_scale"], bf=world.bf, b=world.bf / world.fx)


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_equals_python_statement(seed):
    s = stereo_scene(seed)
    r = O.stereo_matches(s["kl"], s["dl"], s["kr"], s["dr"], s["pl"], s["pr"], s["scale"], s["inv_scale"], s["bf"], s["b"])
    ur, dp = stereo_reference_py(s["kl"], s["dl"], s["kr"], s["dr"], s["pl"], s["pr"], s["scale"], s["inv_scale"], s["bf"], s["b"])
    assert (r["depth"] > 0).sum() > 200, "scene too poor to mean anything"
    assert r["u_right"].tobytes() == ur.tobytes()
    assert r["depth"].tobytes() == dp.tobytes()
    assert r["kept"] == int((dp > 0).sum())


def test_depth_is_plausible():
    """planes at 2, 4 and 8 m: the recovered depths cluster there"""
    s = stereo_scene(3)
    r = O.stereo_matches(s["kl"], s["dl"], s["kr"], s["dr"], s["pl"], s["pr"], s["scale"], s["inv_scale"], s["bf"], s["b"])
    d = r["depth"][r["depth"] > 0]
    near = np.minimum.reduce([np.abs(d - z) / z for z in (2.0, 4.0, 8.0)])
    assert len(d) > 200 and np.median(near) < 0.05


def test_edge_cases():
    s = stereo_scene(2)
    # no right keypoints / no left keypoints
    e = O.stereo_matches(s["kl"], s["dl"], s["kr"][:0], s["dr"][:0], s["pl"], s["pr"], s["scale"], s["inv_scale"], s["bf"], s["b"])
    assert e["kept"] == 0 and (e["depth"] == -1).all() and (e["u_right"] == -1).all()
    e = O.stereo_matches(s["kl"][:0], s["dl"][:0], s["kr"], s["dr"], s["pl"], s["pr"], s["scale"], s["inv_scale"], s["bf"], s["b"])
    assert e["kept"] == 0 and len(e["depth"]) == 0
    # identical images: every SAD is 0, so the median is 0, thDist is 0 and the cut (dist >= thDist) removes everything
    z = O.stereo_matches(s["kl"], s["dl"], s["kl"], s["dl"], s["pl"], s["pl"], s["scale"], s["inv_scale"], s["bf"], s["b"])
    assert (z["sad"] == 0).sum() > 100 and z["kept"] == 0 and (z["depth"] == -1).all()

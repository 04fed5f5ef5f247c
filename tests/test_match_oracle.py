"""Oracle matchers (oracle/match_oracle.c) against an independent brute-force statement of the reference rules
(src/ORBmatcher.cc:45-129, :1328-1470; src/Frame.cc:356-421).  CPU only."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth

f32 = np.float32


def popcount_dist(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_hamming_matches_popcount():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert O.hamming256(a, b) == popcount_dist(a, b)
    z = np.zeros(32, np.uint8)
    assert O.hamming256(z, z) == 0 and O.hamming256(z, ~z) == 256


def brute_window(fr, x, y, r, min_level, max_level):
    """GetFeaturesInArea without the grid: same members, same order (cell column, cell row, index)"""
    k = fr["keys_un"]
    mnx, mny, mxx, mxy = (f32(v) for v in fr["bounds"])
    wi, hi = f32(64) / (mxx - mnx), f32(48) / (mxy - mny)
    x, y, r = f32(x), f32(y), f32(r)
    x0 = max(0, int(np.floor((x - mnx - r) * wi))); x1 = min(63, int(np.ceil((x - mnx + r) * wi)))
    y0 = max(0, int(np.floor((y - mny - r) * hi))); y1 = min(47, int(np.ceil((y - mny + r) * hi)))
    if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
        return []
    out = []
    for i in range(len(k)):
        px = int(np.floor(abs((k["x"][i] - mnx) * wi) + f32(0.5)) * np.sign((k["x"][i] - mnx) * wi))
        py = int(np.floor(abs((k["y"][i] - mny) * hi) + f32(0.5)) * np.sign((k["y"][i] - mny) * hi))
        if not (0 <= px < 64 and 0 <= py < 48) or not (x0 <= px <= x1 and y0 <= py <= y1):
            continue
        if (min_level > 0 or max_level >= 0):
            if k["octave"][i] < min_level or (max_level >= 0 and k["octave"][i] > max_level):
                continue
        if abs(k["x"][i] - x) < r and abs(k["y"][i] - y) < r:
            out.append((px, py, i))
    return [i for _, _, i in sorted(out)]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_features_in_area(seed):
    rng = np.random.default_rng(seed)
    fr = synth.random_frame(rng, 600)
    for _ in range(60):
        x, y, r = rng.uniform(-20, 660), rng.uniform(-20, 500), rng.uniform(1, 60)
        lv = [(-1, -1), (0, 3), (2, -1), (1, 2), (0, -1)][rng.integers(0, 5)]
        assert O.features_in_area(fr, x, y, r, *lv).tolist() == brute_window(fr, x, y, r, *lv)


def test_three_maxima():
    assert O.three_maxima([0] * 30) == (-1, -1, -1)
    h = [0] * 30; h[3], h[7], h[9] = 50, 40, 30
    assert O.three_maxima(h) == (3, 7, 9)
    h[7], h[9] = 4, 3
    assert O.three_maxima(h) == (3, -1, -1)
    h[7] = 6
    assert O.three_maxima(h) == (3, 7, -1)


def brute_frame(cur, pts, desc, R, t, forward, backward, th, check_ori, kf=False, orb_dist=100):
    k = cur["keys_un"]
    fx, fy, cx, cy, bf, b = (f32(v) for v in cur["K"])
    mnx, mny, mxx, mxy = (f32(v) for v in cur["bounds"])
    blocked = cur["claimed"].astype(bool).copy()
    match = np.full(len(k), -1, np.int32)
    hist, nm = [], 0
    R = R.astype(f32)
    for i in range(len(pts)):
        p = pts[i]
        if not p["valid"]:
            continue
        X = np.array([p["x"], p["y"], p["z"]], f32)
        c = [f32(f32(f32(R[r, 0] * X[0]) + f32(R[r, 1] * X[1])) + f32(R[r, 2] * X[2])) + t[r] for r in range(3)]
        invz = f32(1.0 / np.float64(c[2]))
        if invz < 0 and not kf:
            continue
        u = f32(f32(f32(fx * c[0]) * invz) + cx); v = f32(f32(f32(fy * c[1]) * invz) + cy)
        if u < mnx or u > mxx or v < mny or v > mxy:
            continue
        o = int(p["octave"])
        rad = f32(f32(th) * cur["scale_factors"][o])
        lv = (o, -1) if forward and not kf else ((0, o) if backward and not kf else (o - 1, o + 1))
        best, bi = 256, -1
        for j in brute_window(cur, u, v, rad, *lv):
            if blocked[j]:
                continue
            if not kf and cur["u_right"][j] > 0 and abs(f32(f32(u - f32(bf * invz)) - cur["u_right"][j])) > rad:
                continue
            d = popcount_dist(desc[i], cur["desc"][j])
            if d < best:
                best, bi = d, j
        if best <= orb_dist:
            match[bi] = i; blocked[bi] = True if kf else bool(p["blocks"]); nm += 1
            if check_ori:
                rot = f32(p["angle"] - k["angle"][bi])
                if rot < 0:
                    rot = f32(rot + f32(360))
                bn = int(np.floor(f32(rot * f32(1.0 / 30)) + f32(0.5)))
                hist.append((0 if bn == 30 else bn, bi))
    if check_ori:
        cnt = np.bincount([h[0] for h in hist], minlength=30)
        keep = O.three_maxima(cnt)
        for bn, bi in hist:
            if bn not in keep:
                match[bi] = -1; nm -= 1
    return nm, match


@pytest.mark.parametrize("seed,mode", [(0, "win"), (1, "fwd"), (2, "bwd"), (3, "win")])
def test_search_by_projection_frame(seed, mode):
    rng = np.random.default_rng(seed)
    cur = synth.random_frame(rng, 500)
    pts, desc, R, t = synth.last_frame_points(rng, cur, 400)
    fw, bw = mode == "fwd", mode == "bwd"
    for th in (7.0, 15.0):
        n, m = O.search_by_projection_frame(cur, pts, desc, R, t, fw, bw, th, True)
        n2, m2 = brute_frame(cur, pts, desc, R, t, fw, bw, th, True)
        assert n == n2 and np.array_equal(m, m2)
        assert n > 50


def brute_points(F, pts, desc, th, nnratio):
    k = F["keys_un"]
    blocked = F["claimed"].astype(bool).copy()
    match = np.full(len(k), -1, np.int32)
    nm = 0
    for i in range(len(pts)):
        p = pts[i]
        if not p["in_view"]:
            continue
        r = f32(2.5) if p["view_cos"] > f32(0.998) else f32(4.0)
        if th != 1.0:
            r = f32(r * f32(th))
        rs = f32(r * F["scale_factors"][p["level"]])
        bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
        for j in brute_window(F, p["proj_x"], p["proj_y"], rs, int(p["level"]) - 1, int(p["level"])):
            if blocked[j]:
                continue
            if F["u_right"][j] > 0 and abs(f32(p["proj_xr"] - F["u_right"][j])) > rs:
                continue
            d = popcount_dist(desc[i], F["desc"][j])
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, int(k["octave"][j]), j
            elif d < bd2:
                bl2, bd2 = int(k["octave"][j]), d
        if bd <= 100:
            if bl == bl2 and f32(bd) > f32(f32(nnratio) * f32(bd2)):
                continue
            match[bi] = i; blocked[bi] = bool(p["blocks"]); nm += 1
    return nm, match


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_search_by_projection_points(seed):
    rng = np.random.default_rng(100 + seed)
    F = synth.random_frame(rng, 700)
    pts, desc = synth.track_points(rng, F, 500)
    for th in (1.0, 3.0, 5.0):
        n, m = O.search_by_projection_points(F, pts, desc, th, 0.8)
        n2, m2 = brute_points(F, pts, desc, th, 0.8)
        assert n == n2 and np.array_equal(m, m2)
    assert n > 50


# ---- vocabulary-node matchers (ORBmatcher.cc:159-288, :522-655, :657-823) ---------------------------------------
def brute_buckets(mode, A, B, nnratio, check_ori, only_stereo, F12, epi, sigma2, scale):
    va, vb = O.bucket_valid(mode, A, B)
    nodes_a = {int(n): A["node_feat"][A["node_start"][j]:A["node_start"][j + 1]] for j, n in enumerate(A["node_id"])}
    nodes_b = {int(n): B["node_feat"][B["node_start"][j]:B["node_start"][j + 1]] for j, n in enumerate(B["node_id"])}
    match = np.full(len(va), -1, np.int32)
    taken = np.zeros(len(vb), bool)
    hist, nm = [], 0
    ka, kb = A["keys_un"], B["keys_un"]
    F = np.asarray(F12, f32).reshape(3, 3) if F12 is not None else None
    for nid in sorted(set(nodes_a) & set(nodes_b)):
        for i1 in nodes_a[nid]:
            if not va[i1]:
                continue
            if mode == 2:
                st1 = A["u_right"][i1] >= 0
                if only_stereo and not st1:
                    continue
                best, bi = 50, -1
                a = f32(f32(f32(ka["x"][i1] * F[0, 0]) + f32(ka["y"][i1] * F[1, 0])) + F[2, 0])
                b = f32(f32(f32(ka["x"][i1] * F[0, 1]) + f32(ka["y"][i1] * F[1, 1])) + F[2, 1])
                c = f32(f32(f32(ka["x"][i1] * F[0, 2]) + f32(ka["y"][i1] * F[1, 2])) + F[2, 2])
                for i2 in nodes_b[nid]:
                    if not vb[i2]:
                        continue
                    st2 = B["u_right"][i2] >= 0
                    if only_stereo and not st2:
                        continue
                    d = popcount_dist(A["desc"][i1], B["desc"][i2])
                    if d > 50 or d > best:
                        continue
                    if not st1 and not st2:
                        dx, dy = f32(epi[0] - kb["x"][i2]), f32(epi[1] - kb["y"][i2])
                        if f32(f32(dx * dx) + f32(dy * dy)) < f32(f32(100) * scale[kb["octave"][i2]]):
                            continue
                    num = f32(f32(f32(a * kb["x"][i2]) + f32(b * kb["y"][i2])) + c)
                    den = f32(f32(a * a) + f32(b * b))
                    if den == 0 or not (np.float64(f32(f32(num * num) / den)) < 3.84 * np.float64(sigma2[kb["octave"][i2]])):
                        continue
                    best, bi = d, i2
                if bi >= 0:
                    match[i1] = bi; nm += 1
                    hist.append(i1)
            else:
                b1, b2, bi = 256, 256, -1
                for i2 in nodes_b[nid]:
                    if taken[i2] or (mode == 1 and not vb[i2]):
                        continue
                    d = popcount_dist(A["desc"][i1], B["desc"][i2])
                    if d < b1:
                        b2, b1, bi = b1, d, i2
                    elif d < b2:
                        b2 = d
                if (b1 <= 50 if mode == 0 else b1 < 50) and f32(b1) < f32(f32(nnratio) * f32(b2)):
                    match[i1] = bi; taken[bi] = True; nm += 1
                    hist.append(i1)
    if check_ori:
        bins = []
        for i1 in hist:
            rot = f32(ka["angle"][i1] - kb["angle"][match[i1]])
            if rot < 0:
                rot = f32(rot + f32(360))
            bn = int(np.floor(f32(rot * f32(1.0 / 30)) + f32(0.5)))
            bins.append(0 if bn == 30 else bn)
        keep = O.three_maxima(np.bincount(bins, minlength=30)) if bins else (-1, -1, -1)
        for i1, bn in zip(hist, bins):
            if bn not in keep:
                match[i1] = -1; nm -= 1
    return nm, match


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("seed", [0, 1])
def test_bucket_matchers(mode, seed):
    A, B, F12, epi, s2, sc = synth.bow_pair(seed, 400, 450, 250, n_nodes=25)
    for only_stereo in ([False, True] if mode == 2 else [False]):
        n, m = O.match_buckets(mode, A, B, 0.75, True, only_stereo, F12, epi, s2, sc)
        n2, m2 = brute_buckets(mode, A, B, 0.75, True, only_stereo, F12, epi, s2, sc)
        assert n == n2 and np.array_equal(m, m2)
        assert n > (0 if only_stereo else 30)


@pytest.mark.parametrize("seed", [0, 1])
def test_search_by_projection_keyframe(seed):
    """relocalisation overload, ORBmatcher.cc:1472-1599"""
    rng = np.random.default_rng(50 + seed)
    cur = synth.random_frame(rng, 500, claimed_frac=0.1)
    pts, desc, R, t = synth.last_frame_points(rng, cur, 400)
    for th, od in ((10.0, 100), (3.0, 64)):
        n, m = O.search_by_projection_kf(cur, pts, desc, R, t, th, od, True)
        n2, m2 = brute_frame(cur, pts, desc, R, t, False, False, th, True, kf=True, orb_dist=od)
        assert n == n2 and np.array_equal(m, m2)
    assert n > 20


def brute_window_match(F, pts, desc, flags, inv_s2, max_dist):
    k = F["keys_un"]
    blocked = F["claimed"].astype(bool).copy() if flags & 2 else np.zeros(len(k), bool)
    bi, bd = np.full(len(pts), -1, np.int32), np.full(len(pts), 256, np.int32)
    for i, p in enumerate(pts):
        if not p["valid"]:
            continue
        best, b = 256, -1
        for j in brute_window(F, p["u"], p["v"], p["radius"], -1, -1):
            if blocked[j] or k["octave"][j] < p["min_level"] or k["octave"][j] > p["max_level"]:
                continue
            if flags & 1:
                ex, ey = f32(p["u"] - k["x"][j]), f32(p["v"] - k["y"][j])
                e2, lim = f32(f32(ex * ex) + f32(ey * ey)), 5.99
                if F["u_right"][j] >= 0:
                    er = f32(p["ur"] - F["u_right"][j]); e2, lim = f32(e2 + f32(er * er)), 7.8
                if np.float64(f32(e2 * inv_s2[k["octave"][j]])) > lim:
                    continue
            d = popcount_dist(desc[i], F["desc"][j])
            if d < best:
                best, b = d, j
        if b >= 0 and best <= max_dist:
            bi[i], bd[i] = b, best
            if flags & 2:
                blocked[b] = True
    return int((bi >= 0).sum()), bi, bd


@pytest.mark.parametrize("flags,max_dist", [(0, 100), (1, 50), (2, 50), (3, 50)])
def test_match_window(flags, max_dist):
    rng = np.random.default_rng(70 + flags)
    F = synth.random_frame(rng, 500, claimed_frac=0.1)
    pts, desc = synth.window_points(rng, F, 400)
    inv_s2 = (1.0 / (F["scale_factors"] ** 2)).astype(f32)
    n, bi, bd = O.match_window(F, pts, desc, flags, inv_s2, max_dist)
    n2, bi2, bd2 = brute_window_match(F, pts, desc, flags, inv_s2, max_dist)
    assert n == n2 and np.array_equal(bi, bi2) and np.array_equal(bd, bd2)
    assert n > 30


def brute_init(F1, F2, prev, win, nnratio, check_ori):
    k1, k2 = F1["keys_un"], F2["keys_un"]
    md = np.full(len(k2), 2 ** 31 - 1, np.int64)
    m21 = np.full(len(k2), -1, np.int64)
    m12 = np.full(len(k1), -1, np.int32)
    hist, nm = [], 0
    for i1 in range(len(k1)):
        if k1["octave"][i1] > 0:
            continue
        b1, b2, bi = 2 ** 31 - 1, 2 ** 31 - 1, -1
        for i2 in brute_window(F2, prev[i1, 0], prev[i1, 1], f32(win), 0, 0):
            d = popcount_dist(F1["desc"][i1], F2["desc"][i2])
            if md[i2] <= d:
                continue
            if d < b1:
                b2, b1, bi = b1, d, i2
            elif d < b2:
                b2 = d
        if b1 <= 50 and f32(b1) < f32(f32(b2) * f32(nnratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1; nm -= 1
            m12[i1] = bi; m21[bi] = i1; md[bi] = b1; nm += 1
            rot = f32(k1["angle"][i1] - k2["angle"][bi])
            if rot < 0:
                rot = f32(rot + f32(360))
            bn = int(np.floor(f32(rot * f32(1.0 / 30)) + f32(0.5)))
            hist.append((0 if bn == 30 else bn, i1))
    if check_ori:
        keep = O.three_maxima(np.bincount([h[0] for h in hist], minlength=30)) if hist else (-1, -1, -1)
        for bn, i1 in hist:
            if bn not in keep and m12[i1] >= 0:
                m12[i1] = -1; nm -= 1
    return nm, m12


def init_pair(seed, n=900):
    rng = np.random.default_rng(seed)
    F1 = synth.random_frame(rng, n)
    F2 = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in F1.items()}
    F2["keys_un"]["x"] = F1["keys_un"]["x"] + rng.normal(0, 6, n).astype(f32)
    F2["keys_un"]["y"] = F1["keys_un"]["y"] + rng.normal(0, 6, n).astype(f32)
    F2["keys_un"]["angle"] = np.mod(F1["keys_un"]["angle"] + rng.normal(0, 5, n), 360).astype(f32)
    F2["desc"] = synth.flip_bits(rng, F1["desc"], rng.integers(0, 60, n))
    dup = rng.integers(0, n, n // 5)                 # several F1 keypoints look alike: later ones steal matches
    F1["desc"][dup[: len(dup) // 2]] = synth.flip_bits(rng, F1["desc"][dup[len(dup) // 2: 2 * (len(dup) // 2)]], rng.integers(0, 10, len(dup) // 2))
    F1["keys_un"]["x"][dup[: len(dup) // 2]] = F1["keys_un"]["x"][dup[len(dup) // 2: 2 * (len(dup) // 2)]] + 3
    F1["keys_un"]["y"][dup[: len(dup) // 2]] = F1["keys_un"]["y"][dup[len(dup) // 2: 2 * (len(dup) // 2)]] + 3
    F1["keys_un"]["octave"][dup[: len(dup) // 2]] = 0
    F1["keys_un"]["octave"][dup[len(dup) // 2: 2 * (len(dup) // 2)]] = 0
    F2["keys_un"]["octave"] = F1["keys_un"]["octave"].copy()
    prev = np.stack([F1["keys_un"]["x"], F1["keys_un"]["y"]], 1).astype(f32)
    return F1, F2, prev


@pytest.mark.parametrize("seed", [0, 1])
def test_search_for_initialization(seed):
    F1, F2, prev = init_pair(80 + seed, 500)
    for win in (30, 100):
        n, m = O.search_for_initialization(F1, F2, prev, win, 0.9, True)
        n2, m2 = brute_init(F1, F2, prev, win, 0.9, True)
        assert n == n2 and np.array_equal(m, m2)
    assert n > 20

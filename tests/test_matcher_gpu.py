"""GPU parity of the matchers (liborbx.so through the C ABI) against the CPU oracle: identical match arrays and
counts (bit-exact bar: indices)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth
from orbx.matcher import FrameMatchJob, ORBmatcher, fill_view

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher():
    m = ORBmatcher(0.8, True, max_keypoints=4096, max_points=4096, max_jobs=16)
    yield m
    m.close()


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    for _ in range(100):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == O.hamming256(a, b)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("n_kp,n_pts", [(1000, 900), (300, 1500), (2000, 2000)])
def test_search_by_projection_frame(matcher, seed, n_kp, n_pts):
    rng = np.random.default_rng(seed * 31 + n_kp)
    cur = synth.random_frame(rng, n_kp)
    pts, desc, R, t = synth.last_frame_points(rng, cur, n_pts, dup_frac=0.3 if seed % 2 else 0.1)
    fw, bw = seed % 3 == 1, seed % 3 == 2
    for th in (7.0, 15.0):
        n_ref, m_ref = O.search_by_projection_frame(cur, pts, desc, R, t, fw, bw, th, True)
        n, m = matcher.SearchByProjectionLast(cur, pts, desc, R, t, fw, bw, th)
        assert n == n_ref and np.array_equal(m, m_ref), (seed, th, n, n_ref, int((m != m_ref).sum()))
    assert n_ref > 0


def test_search_by_projection_frame_no_orientation_check():
    m = ORBmatcher(0.9, False, max_keypoints=2048, max_points=2048)
    rng = np.random.default_rng(5)
    cur = synth.random_frame(rng, 800)
    pts, desc, R, t = synth.last_frame_points(rng, cur, 800)
    n_ref, m_ref = O.search_by_projection_frame(cur, pts, desc, R, t, False, False, 7.0, False)
    n, mm = m.SearchByProjectionLast(cur, pts, desc, R, t, False, False, 7.0)
    assert n == n_ref and np.array_equal(mm, m_ref)
    m.close()


@pytest.mark.parametrize("seed", range(5))
def test_search_by_projection_points(matcher, seed):
    rng = np.random.default_rng(200 + seed)
    F = synth.random_frame(rng, 1200)
    pts, desc = synth.track_points(rng, F, 1500, dup_frac=0.25)
    for th in (1.0, 3.0, 5.0):
        n_ref, m_ref = O.search_by_projection_points(F, pts, desc, th, 0.8)
        n, m = matcher.SearchByProjection(F, pts, desc, th)
        assert n == n_ref and np.array_equal(m, m_ref), (seed, th, n, n_ref)


def test_prefilled_match_is_preserved(matcher):
    rng = np.random.default_rng(9)
    F = synth.random_frame(rng, 500)
    pts, desc = synth.track_points(rng, F, 300)
    pre = np.where(F["claimed"] > 0, 10000 + np.arange(500), -1).astype(np.int32)
    n_ref, m_ref = O.search_by_projection_points(F, pts, desc, 3.0, 0.8, match=pre)
    n, m = matcher.SearchByProjection(F, pts, desc, 3.0, match=pre)
    assert n == n_ref and np.array_equal(m, m_ref)
    assert np.array_equal(m[F["claimed"] > 0], pre[F["claimed"] > 0])


def test_empty_inputs(matcher):
    rng = np.random.default_rng(1)
    F = synth.random_frame(rng, 100)
    pts, desc = synth.track_points(rng, F, 10)
    n, m = matcher.SearchByProjection(F, pts[:0], desc[:0], 3.0)
    assert n == 0 and (m == -1).all()
    E = dict(F, keys_un=F["keys_un"][:0], desc=F["desc"][:0], u_right=F["u_right"][:0], claimed=F["claimed"][:0])
    n, m = matcher.SearchByProjection(E, pts, desc, 3.0)
    assert n == 0 and len(m) == 0


def test_batched_device_jobs(matcher):
    """orbx_match_projection_frame_device: 8 independent (current, last) pairs in one launch, device-resident"""
    import torch
    rng = np.random.default_rng(77)
    keep, jobs, refs = [], (FrameMatchJob * 8)(), []
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
    for j in range(8):
        cur = synth.random_frame(rng, 600 + 50 * j)
        pts, desc, R, t = synth.last_frame_points(rng, cur, 500 + 30 * j)
        refs.append(O.search_by_projection_frame(cur, pts, desc, R, t, False, False, 7.0, True))
        n = len(cur["keys_un"])
        t_keys, t_desc, t_ur, t_cl = dev(cur["keys_un"]), dev(cur["desc"]), dev(cur["u_right"]), dev(cur["claimed"])
        t_sf, t_pts, t_pd = dev(cur["scale_factors"]), dev(pts), dev(desc)
        t_n = torch.tensor([n], dtype=torch.int32, device="cuda")
        t_match = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        t_nm = torch.zeros(1, dtype=torch.int32, device="cuda")
        keep += [t_keys, t_desc, t_ur, t_cl, t_sf, t_pts, t_pd, t_n, t_match, t_nm]
        J = jobs[j]
        J.cur.n = 0 if j % 2 else n
        J.cur.n_dev = t_n.data_ptr() if j % 2 else None       # odd jobs read N on the device
        J.cur.keys_un, J.cur.desc, J.cur.u_right, J.cur.claimed = t_keys.data_ptr(), t_desc.data_ptr(), t_ur.data_ptr(), t_cl.data_ptr()
        J.cur.scale_factors = t_sf.data_ptr()
        fill_view(J.cur, cur["bounds"], cur["K"], 8)
        J.n_last, J.pts, J.last_desc = len(pts), t_pts.data_ptr(), t_pd.data_ptr()
        J.Rcw[:] = R.reshape(9).tolist(); J.tcw[:] = t.tolist()
        J.forward = J.backward = 0
        J.th, J.check_ori = 7.0, 1
        J.match, J.nmatches = t_match.data_ptr(), t_nm.data_ptr()
    d_jobs = torch.from_numpy(np.frombuffer(bytes(jobs), np.uint8).copy()).cuda()
    matcher.search_frames_device(d_jobs.data_ptr(), 8, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for j in range(8):
        n_ref, m_ref = refs[j]
        assert keep[10 * j + 9].item() == n_ref
        assert np.array_equal(keep[10 * j + 8].cpu().numpy(), m_ref)


@pytest.mark.parametrize("seed", range(3))
def test_dense_windows_overflow_path(matcher, seed):
    """small image, many keypoints, large radius: windows hold hundreds of candidates (> the per-point list) """
    rng = np.random.default_rng(300 + seed)
    cur = synth.random_frame(rng, 3000, w=320, h=240)
    pts, desc, R, t = synth.last_frame_points(rng, cur, 1500, dup_frac=0.3)
    n_ref, m_ref = O.search_by_projection_frame(cur, pts, desc, R, t, False, False, 15.0, True)
    n, m = matcher.SearchByProjectionLast(cur, pts, desc, R, t, False, False, 15.0)
    assert n == n_ref and np.array_equal(m, m_ref)
    tp, tdesc = synth.track_points(rng, cur, 1500, dup_frac=0.3)
    n_ref, m_ref = O.search_by_projection_points(cur, tp, tdesc, 5.0, 0.8)
    n, m = matcher.SearchByProjection(cur, tp, tdesc, 5.0)
    assert n == n_ref and np.array_equal(m, m_ref)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("seed", range(4))
def test_bucket_matchers(matcher, mode, seed):
    """SearchByBoW(KF, F), SearchByBoW(KF, KF), SearchForTriangulation vs the oracle"""
    A, B, F12, epi, s2, sc = synth.bow_pair(40 + seed, 1000 + 100 * seed, 1100, 600, n_nodes=90 if seed % 2 else 12)
    matcher.mfNNratio = 0.75
    try:
        for only_stereo in ([False, True] if mode == 2 else [False]):
            n_ref, m_ref = O.match_buckets(mode, A, B, 0.75, True, only_stereo, F12, epi, s2, sc)
            if mode == 0:
                n, mf = matcher.SearchByBoW(A, B)
                inv = np.full(len(B["keys_un"]), -1, np.int32)
                inv[m_ref[m_ref >= 0]] = np.nonzero(m_ref >= 0)[0]
                assert n == n_ref and np.array_equal(mf, inv)
            elif mode == 1:
                n, m = matcher.SearchByBoWKF(A, B)
                assert n == n_ref and np.array_equal(m, m_ref)
            else:
                n, pairs = matcher.SearchForTriangulation(A, B, F12, epi, s2, sc, only_stereo)
                idx1 = np.nonzero(m_ref >= 0)[0]
                assert n == n_ref and np.array_equal(pairs, np.stack([idx1, m_ref[idx1]], 1))
            assert n_ref > 0
    finally:
        matcher.mfNNratio = 0.8


def test_bucket_matchers_disjoint_vocabularies(matcher):
    A, B, F12, epi, s2, sc = synth.bow_pair(9, 300, 300, 100, n_nodes=10)
    B = dict(B, node_id=B["node_id"] + 1000)          # no shared node at all
    n, m = matcher.SearchByBoWKF(A, B)
    assert n == 0 and (m == -1).all()


@pytest.mark.parametrize("seed", range(4))
def test_search_by_projection_keyframe(matcher, seed):
    """relocalisation overload SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist), ORBmatcher.cc:1472-1599"""
    rng = np.random.default_rng(400 + seed)
    cur = synth.random_frame(rng, 1200, claimed_frac=0.15)
    pts, desc, R, t = synth.last_frame_points(rng, cur, 1000, dup_frac=0.3)
    for th, od in ((10.0, 100), (3.0, 64)):
        n_ref, m_ref = O.search_by_projection_kf(cur, pts, desc, R, t, th, od, True)
        n, m = matcher.SearchByProjectionKF(cur, pts, desc, R, t, th, od)
        assert n == n_ref and np.array_equal(m, m_ref)
    assert n_ref > 0


@pytest.mark.parametrize("flags,max_dist", [(0, 100), (1, 50), (2, 50), (3, 50)])
@pytest.mark.parametrize("seed", range(3))
def test_match_window(matcher, flags, max_dist, seed):
    """window + Hamming core of SearchByProjection(KF, Scw, ...), Fuse x2, SearchBySim3 (ORBmatcher.cc:290-403, :825-1326)"""
    rng = np.random.default_rng(500 + 10 * flags + seed)
    F = synth.random_frame(rng, 1500, claimed_frac=0.1)
    pts, desc = synth.window_points(rng, F, 1200, th=3.0 if seed else 8.0, dup_frac=0.3)
    inv_s2 = (1.0 / (F["scale_factors"] ** 2)).astype(np.float32)
    n_ref, bi_ref, bd_ref = O.match_window(F, pts, desc, flags, inv_s2, max_dist)
    n, bi, bd = matcher.MatchWindow(F, pts, desc, flags, inv_s2, max_dist)
    assert n == n_ref and np.array_equal(bi, bi_ref) and np.array_equal(bd, bd_ref)
    assert n_ref > 100


@pytest.mark.parametrize("seed", range(4))
def test_search_for_initialization(seed):
    """monocular initialisation matcher, ORBmatcher.cc:405-520 (sequential stealing semantics)"""
    from test_match_oracle import init_pair
    F1, F2, prev = init_pair(600 + seed, 1500)
    m = ORBmatcher(0.9, True, max_keypoints=2048, max_points=2048)
    try:
        for win in (30, 100):
            n_ref, m_ref = O.search_for_initialization(F1, F2, prev, win, 0.9, True)
            n, mm, newprev = m.SearchForInitialization(F1, F2, prev, win)
            assert n == n_ref and np.array_equal(mm, m_ref)
            sel = m_ref >= 0
            assert np.array_equal(newprev[sel, 0], F2["keys_un"]["x"][m_ref[sel]]) and np.array_equal(newprev[~sel], prev[~sel])
        assert n_ref > 50
    finally:
        m.close()

"""GPU parity of Frame::ComputeStereoMatches (liborbx.so, orbx_stereo_* through the C ABI) against the CPU oracle.
Bar: mvuRight and mvDepth bit-equal (float32 bit patterns), same survivors of the median cut."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth
from orbx.extractor import ORBextractor
from orbx.stereo import StereoMatcher, StereoSide

pytestmark = pytest.mark.gpu


def world_pair(seed, w=640, h=480, **kw):
    world = synth.stereo_world(seed, w, h, **kw)
    return world, world.render(0.03 * seed, 0.01, 0.002 * seed), world.render(0.03 * seed, 0.01, 0.002 * seed, right=True)


def oracle_pair(left, right, nfeat):
    exl, exr = O.Extractor(nfeat, 1.2, 8, 20, 7), O.Extractor(nfeat, 1.2, 8, 20, 7)
    kl, dl = exl(left)
    kr, dr = exr(right)
    t = exl.tables()
    return kl, dl, kr, dr, [exl.level(l) for l in range(8)], [exr.level(l) for l in range(8)], t["scale"], t["inv_scale"]


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("seed,w,h,nfeat,K", [(0, 640, 480, 1000, None), (1, 640, 480, 1000, None), (4, 640, 480, 2500, None),
                                              (2, 1241, 376, 2000, dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=386.1448))])
def test_host_entry_point(seed, w, h, nfeat, K):
    world, left, right = world_pair(seed, w, h, **(K or {}))
    bf, b = world.bf, world.bf / world.fx
    el = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    er = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    kl, dl = el(left)
    kr, dr = er(right)
    okl, odl, okr, odr, pl, pr, sc, isc = oracle_pair(left, right, nfeat)
    assert kl.tobytes() == okl.tobytes() and kr.tobytes() == okr.tobytes()
    ref = O.stereo_matches(okl, odl, okr, odr, pl, pr, sc, isc, bf, b)
    sm = StereoMatcher(max_keypoints=4096)
    ur, dp, kept = sm.ComputeStereoMatches(el, er, kl, dl, kr, dr, bf, b)
    assert (ref["depth"] > 0).sum() > 150
    assert same_bits(ur, ref["u_right"]), np.nonzero(ur != ref["u_right"])[0][:8]
    assert same_bits(dp, ref["depth"])
    assert kept == ref["kept"] and sm.last_launches() == 2
    # the variant that reads keypoints and descriptors from the extractors' own device buffers
    ur2, dp2, kept2 = sm.ComputeStereoMatchesFromExtractors(el, er, len(kl), bf, b)
    assert same_bits(ur2, ref["u_right"]) and same_bits(dp2, ref["depth"]) and kept2 == ref["kept"]
    sm.close(); el.close(); er.close()


def test_edge_cases():
    world, left, right = world_pair(3)
    bf, b = world.bf, world.bf / world.fx
    el = ORBextractor(1000, 1.2, 8, 20, 7)
    er = ORBextractor(1000, 1.2, 8, 20, 7)
    kl, dl = el(left)
    kr, dr = er(right)
    sm = StereoMatcher(max_keypoints=2048)
    ur, dp, kept = sm.ComputeStereoMatches(el, er, kl, dl, kr[:0], dr[:0], bf, b)
    assert kept == 0 and (ur == -1).all() and (dp == -1).all()
    ur, dp, kept = sm.ComputeStereoMatches(el, er, kl[:0], dl[:0], kr, dr, bf, b)
    assert kept == 0 and len(ur) == 0
    # the same image on both sides: every SAD is 0, so the median cut removes everything (reference behaviour)
    ur, dp, kept = sm.ComputeStereoMatches(el, el, kl, dl, kl, dl, bf, b)
    assert kept == 0 and (dp == -1).all()
    # a tiny baseline limits the disparity band (maxD = bf / b): compare with the oracle again
    _, _, _, _, pl, pr, sc, isc = oracle_pair(left, right, 1000)
    ref = O.stereo_matches(kl, dl, kr, dr, pl, pr, sc, isc, 12.0, 1.0)
    ur, dp, kept = sm.ComputeStereoMatches(el, er, kl, dl, kr, dr, 12.0, 1.0)
    assert same_bits(ur, ref["u_right"]) and same_bits(dp, ref["depth"]) and kept == ref["kept"]
    # too many keypoints for the handle is an error, not a truncation
    small = StereoMatcher(max_keypoints=100)
    from orbx._lib import OrbxError
    with pytest.raises(OrbxError):
        small.ComputeStereoMatches(el, er, kl, dl, kr, dr, bf, b)
    # more keypoints per image than the search kernel's shared memory holds (8 bytes per right keypoint) is refused at create
    with pytest.raises(OrbxError) as e:
        StereoMatcher(max_keypoints=40000)
    assert e.value.status == -4 or "CAPACITY" in str(e.value)
    StereoMatcher(max_keypoints=20000).close()
    small.close(); sm.close(); el.close(); er.close()


def test_batched_device_pairs():
    """one extractor run over L0 R0 L1 R1 ...; the stereo kernels read keypoints, descriptors, counts and pyramids in place"""
    import torch
    n_pairs, w, h = 4, 640, 480
    frames, refs = [], []
    for p in range(n_pairs):
        world, left, right = world_pair(10 + p)
        frames += [left, right]
        okl, odl, okr, odr, pl, pr, sc, isc = oracle_pair(left, right, 1000)
        refs.append(O.stereo_matches(okl, odl, okr, odr, pl, pr, sc, isc, world.bf, world.bf / world.fx))
    bf, b = world.bf, world.bf / world.fx
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2 * n_pairs)
    cap = ex.capacity
    d_img = torch.from_numpy(np.stack(frames)).cuda()
    d_kps = torch.zeros((2 * n_pairs, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((2 * n_pairs, cap, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2 * n_pairs, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ex.run_device(d_img.data_ptr(), w * h, 2 * n_pairs, w, h, w, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), s)
    d_ur = torch.zeros((n_pairs, cap), dtype=torch.float32, device="cuda")
    d_dp = torch.zeros((n_pairs, cap), dtype=torch.float32, device="cuda")
    d_kept = torch.zeros(n_pairs, dtype=torch.int32, device="cuda")
    left = StereoSide(d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), 2 * cap, 2, ex._h, 0, 2, cap)
    right = StereoSide(d_kps.data_ptr() + 28 * cap, d_desc.data_ptr() + 32 * cap, d_cnt.data_ptr() + 4, 2 * cap, 2, ex._h, 1, 2, cap)
    sm = StereoMatcher(max_keypoints=cap, max_pairs=n_pairs)
    sm.matches_device(left, right, n_pairs, bf, b, d_ur.data_ptr(), d_dp.data_ptr(), cap, d_kept.data_ptr(), s)
    torch.cuda.synchronize()
    cnt, ur, dp, kept = d_cnt.cpu().numpy(), d_ur.cpu().numpy(), d_dp.cpu().numpy(), d_kept.cpu().numpy()
    for p in range(n_pairs):
        n = cnt[2 * p]
        assert n == len(refs[p]["depth"])
        assert same_bits(ur[p, :n], refs[p]["u_right"]) and same_bits(dp[p, :n], refs[p]["depth"]), p
        assert kept[p] == refs[p]["kept"]
    sm.close(); ex.close()

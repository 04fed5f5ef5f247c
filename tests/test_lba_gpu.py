"""GPU parity of LocalBundleAdjustment (liborbx.so through the C ABI) against the CPU oracle.
Tolerance (north_star): 1e-4 relative on pose / point updates, identical outlier sets; the first trial's reduced camera
system is held to 1e-9 relative (same arithmetic, different summation order)."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth
from orbx.optimizer import Optimizer

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def opt():
    o = Optimizer(max_keyframes=40, max_points=4000, max_edges=20000)
    yield o
    o.close()


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def check_against_oracle(opt, p, its1=5, its2=10):
    ref = O.lba_solve(p, its1, its2, want_system=True)
    got = opt.LocalBundleAdjustment(p, its1, its2, want_system=True)
    # first Levenberg trial: lambda, reduced system, pose update
    assert abs(got["lambda0"] - ref["lambda0"]) <= 1e-12 * abs(ref["lambda0"])
    assert rel(got["Hschur"], ref["Hschur"]) < 1e-9 and rel(got["bschur"], ref["bschur"]) < 1e-9
    assert rel(got["xp"], ref["xp"]) < 1e-6
    # result of the whole 5 + 10 schedule: updates within 1e-4 relative
    assert got["trials"] == ref["trials"]
    d_ref_kf, d_got_kf = ref["kf"] - p["kf_pose"], got["kf"] - p["kf_pose"]
    d_ref_pt, d_got_pt = ref["pts"] - p["pts"], got["pts"] - p["pts"]
    assert rel(d_got_kf, d_ref_kf) < TOL, rel(d_got_kf, d_ref_kf)
    assert rel(d_got_pt, d_ref_pt) < TOL, rel(d_got_pt, d_ref_pt)
    assert np.array_equal(got["erase"], ref["erase"])
    assert np.allclose(got["chi2"], ref["chi2"], rtol=1e-4, atol=1e-6)
    return ref, got


@pytest.mark.parametrize("stereo", [False, True])
@pytest.mark.parametrize("n_fixed", [0, 1, 5])
def test_c3_config(opt, stereo, n_fixed):
    """20 keyframes x 3000 points x ~12k edges, 5 + 10 iterations"""
    p = synth.lba_problem(1 + n_fixed, n_kf=20 + n_fixed if n_fixed == 5 else 20, n_pts=3000, stereo=stereo, n_fixed=n_fixed)
    ref, got = check_against_oracle(opt, p)
    assert ref["erase"].sum() > 100 and got["trials"] >= 10


@pytest.mark.parametrize("stereo", [False, True])
def test_single_window_on_the_cluster_kernel(stereo, monkeypatch):
    """orbx_lba_solve_host normally gives a single window the whole GPU (the cooperative kernel); ORBX_LBA_GRID=0 at handle creation
    keeps it on the 16-CTA cluster kernel, the path a deployment picks when other threads' kernels must not wait for the window"""
    monkeypatch.setenv("ORBX_LBA_GRID", "0")
    o = Optimizer(max_keyframes=40, max_points=4000, max_edges=20000)
    monkeypatch.delenv("ORBX_LBA_GRID")
    p = synth.lba_problem(3, n_kf=20, n_pts=3000, stereo=stereo, n_fixed=1)
    ref, got = check_against_oracle(o, p)
    assert got["trials"] >= 10
    both = opt_result_pair(o, p)
    assert rel(both[0]["pts"] - p["pts"], both[1]["pts"] - p["pts"]) < 1e-6      # cluster kernel against whole-GPU kernel: summation order only
    o.close()


def opt_result_pair(o_cluster, p):
    o_grid = Optimizer(max_keyframes=40, max_points=4000, max_edges=20000)
    r = (o_cluster.LocalBundleAdjustment(p), o_grid.LocalBundleAdjustment(p))
    o_grid.close()
    return r


@pytest.mark.parametrize("seed", range(4))
def test_small_and_mixed(opt, seed):
    p = synth.lba_problem(10 + seed, n_kf=4 + seed, n_pts=80 + 40 * seed, obs_per_pt=2 + seed % 3, n_fixed=1)
    rng = np.random.default_rng(seed)
    p["e_stereo"] = (rng.random(len(p["e_kf"])) < 0.5).astype(np.uint8)      # mono and stereo edges in one window
    check_against_oracle(opt, p, 3, 4)


def test_free_keyframe_without_edges_and_disjoint_pairs(opt):
    """a free keyframe nobody observes from (its diagonal block is lambda I alone) and keyframe pairs that share no landmark (empty
    blocks of the reduced system): the empty chunks the host lists for them are what writes those blocks on both one-launch paths"""
    p = synth.lba_problem(21, n_kf=8, n_pts=300, obs_per_pt=3, n_fixed=1)
    free = np.nonzero(p["kf_fixed"] == 0)[0]
    lonely = int(free[-1])
    keep = p["e_kf"] != lonely
    for k in ("e_kf", "e_pt", "e_inv_sigma2", "e_stereo"):
        p[k] = p[k][keep]
    p["e_obs"] = p["e_obs"][keep]
    ref = O.lba_solve(p)
    got = opt.LocalBundleAdjustment(p)
    assert got["trials"] == ref["trials"] and np.array_equal(got["erase"], ref["erase"])
    assert np.allclose(got["kf"][lonely], p["kf_pose"][lonely], rtol=0, atol=1e-14)   # nothing pulls on it (15 zero updates, each renormalising the quaternion)
    assert rel(got["pts"] - p["pts"], ref["pts"] - p["pts"]) < TOL
    opt.begin(p)
    got2 = opt.end()
    assert got2["trials"] == ref["trials"] and rel(got2["pts"] - p["pts"], ref["pts"] - p["pts"]) < TOL


def test_fixed_keyframes_do_not_move(opt):
    p = synth.lba_problem(3, n_kf=8, n_pts=300, n_fixed=3)
    got = opt.LocalBundleAdjustment(p)
    assert np.array_equal(got["kf"][:3], p["kf_pose"][:3])


def test_stop_flag(opt):
    p = synth.lba_problem(4, n_kf=4, n_pts=50, n_fixed=1)
    p["stop_flag"] = np.ones(1, np.uint8)
    got = opt.LocalBundleAdjustment(p)
    assert got["stopped"] == 1 and np.array_equal(got["kf"], p["kf_pose"]) and np.array_equal(got["pts"], p["pts"])
    assert got["trials"] == 0


def test_capacity_and_bad_edges(opt):
    from orbx._lib import OrbxError
    p = synth.lba_problem(5, n_kf=50, n_pts=100)
    with pytest.raises(OrbxError):
        opt.LocalBundleAdjustment(p)          # 50 keyframes > handle's 40
    p = synth.lba_problem(5, n_kf=4, n_pts=50)
    p["e_pt"][0] = 999
    with pytest.raises(OrbxError):
        opt.LocalBundleAdjustment(p)


def test_build_schur_timed_matches_first_system(opt):
    p = synth.lba_problem(7, n_kf=12, n_pts=1500, n_fixed=2)
    ref = O.lba_solve(p, 1, 0, want_system=True)
    ms, Hs, bs = opt.build_schur_timed(p, ref["lambda0"], reps=3, want_system=True)
    assert ms > 0 and rel(Hs, ref["Hschur"]) < 1e-9 and rel(bs, ref["bschur"]) < 1e-9


def test_multikernel_path_matches_too(monkeypatch):
    """windows too large for the cluster kernel take the host-driven multi-kernel path; force it on a small problem"""
    monkeypatch.setenv("ORBX_LBA_MULTIKERNEL", "1")
    o = Optimizer(max_keyframes=40, max_points=4000, max_edges=20000)
    try:
        p = synth.lba_problem(21, n_kf=10, n_pts=800, n_fixed=2, stereo=True)
        check_against_oracle(o, p)
    finally:
        o.close()


def test_large_window_uses_multikernel_path():
    """45 free keyframes: the upper block triangle of H_schur no longer fits shared memory"""
    o = Optimizer(max_keyframes=48, max_points=3000, max_edges=20000)
    try:
        p = synth.lba_problem(22, n_kf=46, n_pts=1500, n_fixed=1)
        check_against_oracle(o, p, 3, 3)
    finally:
        o.close()


def test_many_windows_in_flight():
    """orbx_lba_solve_begin / _end: 6 independent windows on 6 handles, results identical to the synchronous call"""
    probs = [synth.lba_problem(30 + i, n_kf=8 + i, n_pts=400 + 100 * i, n_fixed=1, stereo=bool(i % 2)) for i in range(6)]
    hs = [Optimizer(max_keyframes=16, max_points=1000, max_edges=5000) for _ in probs]
    try:
        for h, p in zip(hs, probs):
            h.begin(p)
        outs = [h.end() for h in hs]
        for h, p, o in zip(hs, probs, outs):
            ref = h.LocalBundleAdjustment(p)
            assert o["trials"] == ref["trials"]
            assert rel(o["kf"], ref["kf"]) < 1e-9 and rel(o["pts"], ref["pts"]) < 1e-9
            assert np.array_equal(o["erase"], ref["erase"])
            orc = O.lba_solve(p)
            assert np.array_equal(o["erase"], orc["erase"]) and rel(o["pts"] - p["pts"], orc["pts"] - p["pts"]) < TOL
    finally:
        for h in hs:
            h.close()

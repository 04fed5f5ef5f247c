#!/usr/bin/env python
"""Writes tests/golden/*.npz: outputs of the CPU oracle on seeded synthetic inputs (the reference has no golden vectors
of its own and cannot be built here).  They freeze the oracle (a change in oracle/ that moves a result
fails tests/test_golden.py) and give the -m gpu tests fixtures that do not need the oracle at run time.
Run from the repo root:  python tools/make_golden.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from oracle import oracle_py as O  # noqa: E402
from orbx import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    ex = O.Extractor(1000, 1.2, 8, 20, 7)
    digests = {}
    for kind in ("rect", "noise", "sparse"):
        for seed in range(4):
            kp, de = ex(synth.frame(kind, seed))
            digests["extract/%s/%d" % (kind, seed)] = "%d:%s" % (len(kp), sha(kp, de))
            if seed == 0:
                np.savez_compressed(os.path.join(OUT, "extract_vga_%s_seed0.npz" % kind), kps=kp, desc=de)
    rng = np.random.default_rng(1234)
    cur = synth.random_frame(rng, 800)
    pts, desc, R, t = synth.last_frame_points(rng, cur, 700)
    n, m = O.search_by_projection_frame(cur, pts, desc, R, t, False, False, 7.0, True)
    tp, tdesc = synth.track_points(rng, cur, 700)
    n2, m2 = O.search_by_projection_points(cur, tp, tdesc, 3.0, 0.8)
    np.savez_compressed(os.path.join(OUT, "match_projection_seed1234.npz"), n_frame=n, match_frame=m, n_points=n2, match_points=m2)
    A, B, F12, epi, s2, sc = synth.bow_pair(77, 600, 650, 350, n_nodes=40)
    bow = {}
    for mode in range(3):
        nm, ma = O.match_buckets(mode, A, B, 0.75, True, False, F12, epi, s2, sc)
        bow["n%d" % mode], bow["m%d" % mode] = nm, ma
    np.savez_compressed(os.path.join(OUT, "match_buckets_seed77.npz"), **bow)
    p = synth.lba_problem(5, n_kf=8, n_pts=400, n_fixed=1)
    r = O.lba_solve(p, 5, 10, want_system=True)
    np.savez_compressed(os.path.join(OUT, "lba_seed5.npz"), kf=r["kf"], pts=r["pts"], erase=r["erase"], chi2=r["chi2"], trials=r["trials"],
                        Hschur=r["Hschur"], bschur=r["bschur"], lambda0=r["lambda0"])
    p = synth.lba_problem(6, n_kf=6, n_pts=300, stereo=True, n_fixed=1)
    r = O.lba_solve(p, 5, 10)
    np.savez_compressed(os.path.join(OUT, "lba_stereo_seed6.npz"), kf=r["kf"], pts=r["pts"], erase=r["erase"], trials=r["trials"])
    # Frame::ComputeStereoMatches on one rendered pair of the 3-plane world
    world = synth.stereo_world(0)
    exl, exr = O.Extractor(1000, 1.2, 8, 20, 7), O.Extractor(1000, 1.2, 8, 20, 7)
    kl, dl = exl(world.render(0.0, 0.01, 0.0))
    kr, dr = exr(world.render(0.0, 0.01, 0.0, right=True))
    t = exl.tables()
    st = O.stereo_matches(kl, dl, kr, dr, [exl.level(l) for l in range(8)], [exr.level(l) for l in range(8)], t["scale"], t["inv_scale"],
                          world.bf, world.bf / world.fx)
    np.savez_compressed(os.path.join(OUT, "stereo_seed0.npz"), u_right=st["u_right"], depth=st["depth"], kept=st["kept"])
    # Optimizer::PoseOptimization
    pp = synth.pose_problem(3, n=300)
    pr_ = O.pose_optimize(pp)
    np.savez_compressed(os.path.join(OUT, "pose_seed3.npz"), pose=pr_["pose"], outlier=pr_["outlier"], n_inliers=pr_["n_inliers"])
    with open(os.path.join(OUT, "digests.txt"), "w") as f:
        for k in sorted(digests):
            f.write("%s %s\n" % (k, digests[k]))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()

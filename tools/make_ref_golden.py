#!/usr/bin/env python
"""Writes tests/golden/ref_extract_digests.txt, ref_extract_kitti_rect_seed1.npz and ref_match.npz: outputs of the REFERENCE's own
ORBextractor and ORBmatcher (src/ORBextractor.cc, ORBmatcher.cc, Frame.cc, MapPoint.cc compiled unmodified into oracle/_ref/*.so, see
oracle/cvmini/cvmini.hpp for what is real and what is stood in) on seeded synthetic inputs.  Needs /root/reference (build container only); the
fixtures it writes travel, so the oracle and the CUDA path are checked against the reference's results on any box.
Run from the repo root:  make -C oracle ref && python tools/make_ref_golden.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from oracle import oracle_py as O  # noqa: E402
from orbx import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, kind, seed, width, height, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
CASES = (
    [("vga/%s/%d" % (k, s), k, s, 640, 480, 1000, 1.2, 8, 20, 7) for k in ("rect", "noise", "sparse") for s in range(4)]
    + [("kitti/rect/%d" % s, "rect", s, 1241, 376, 2000, 1.2, 8, 20, 7) for s in range(2)]
    + [("kitti/sparse/0", "sparse", 0, 1241, 376, 2000, 1.2, 8, 20, 7),
       ("euroc/rect/5", "rect", 5, 752, 480, 1200, 1.2, 8, 20, 7),
       ("qvga/noise/6", "noise", 6, 320, 240, 500, 1.2, 8, 20, 7),
       ("vga/flat/0", "flat", 0, 640, 480, 1000, 1.2, 8, 20, 7),
       ("vga/rect/levels4", "rect", 7, 640, 480, 800, 1.5, 4, 20, 7),
       ("vga/rect/levels6_sf1.1", "rect", 8, 640, 480, 1500, 1.1, 6, 15, 5),
       ("vga/noise/th40", "noise", 9, 640, 480, 600, 1.2, 8, 40, 12),
       ("odd/rect/10", "rect", 10, 601, 397, 700, 1.2, 8, 20, 7)]
)


# matcher cases: (name, seed, keypoints, points, dz of the last pose (decides bForward / bBackward), mono, th)
MATCH_CASES = [("frame/none", 11, 1000, 900, 0.0, 0, 7.0), ("frame/forward", 12, 1000, 900, 1.0, 0, 7.0),
               ("frame/backward", 13, 800, 1000, -1.0, 0, 15.0), ("frame/mono", 14, 1200, 700, 1.0, 1, 7.0)]
POINT_CASES = [("points/th1", 21, 1000, 900, 1.0), ("points/th3", 22, 1000, 1500, 3.0), ("points/th5", 23, 600, 900, 5.0)]


def match_inputs(case):
    """seeded inputs of a MATCH_CASES entry and the reference's own bForward / bBackward for them (ORBmatcher.cc:1346-1351)"""
    name, seed, n, npts, dz, mono, th = case
    rng = np.random.default_rng(seed)
    cur = synth.random_frame(rng, n)
    pts, desc, R, t = synth.last_frame_points(rng, cur, npts)
    tlw = (t + np.array([0, 0, dz], np.float32)).astype(np.float32)
    return cur, pts, desc, R, t, R.copy(), tlw


def point_inputs(case):
    name, seed, n, npts, th = case
    rng = np.random.default_rng(seed)
    cur = synth.random_frame(rng, n)
    tp, tdesc = synth.track_points(rng, cur, npts)
    return cur, tp, tdesc


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    lines = []
    for name, kind, seed, w, h, nf, sf, nl, ith, mth in CASES:
        ex = O.RefExtractor(nf, sf, nl, ith, mth)
        kp, de = ex(synth.frame(kind, seed, w, h))
        ex.close()
        lines.append("%s %d:%s" % (name, len(kp), sha(kp, de)))
        if name == "kitti/rect/1":
            np.savez_compressed(os.path.join(OUT, "ref_extract_kitti_rect_seed1.npz"), kps=kp, desc=de)
        print(lines[-1])
    with open(os.path.join(OUT, "ref_extract_digests.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    # the reference's own ORBmatcher on seeded frames (oracle/_ref/liborbmatcher_ref.so)
    out = {}
    for case in MATCH_CASES:
        cur, pts, desc, R, t, Rlw, tlw = match_inputs(case)
        n, m = O.ref_search_by_projection_frame(cur, pts, desc, R, t, Rlw, tlw, case[5], case[6])
        out[case[0] + "/n"], out[case[0] + "/match"] = np.int32(n), m
        print(case[0], n)
    for case in POINT_CASES:
        cur, tp, tdesc = point_inputs(case)
        n, m = O.ref_search_by_projection_points(cur, tp, tdesc, case[4], 0.8)
        out[case[0] + "/n"], out[case[0] + "/match"] = np.int32(n), m
        print(case[0], n)
    np.savez_compressed(os.path.join(OUT, "ref_match.npz"), **out)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Writes tests/golden/ref_extract_digests.txt and ref_extract_kitti_rect_seed1.npz: outputs of the REFERENCE's own
ORBextractor (src/ORBextractor.cc compiled unmodified into oracle/_ref/liborbextractor_ref.so, see oracle/cvmini/cvmini.hpp
for what is real and what is stood in) on seeded synthetic frames.  Needs /root/reference (build container only); the
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


if __name__ == "__main__":
    main()

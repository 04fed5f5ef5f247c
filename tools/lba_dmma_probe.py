"""Schur build of one Levenberg trial of config C3 (20 keyframes, 3000 points, ~11 k edges), scalar pair accumulation (k2_pairs) or, with
ORBX_LBA_DMMA=1 in the environment, the mma.sync.m8n8k4.f64 form (k2_pairs_mma): device time per build and agreement of H_schur / b_schur
with the CPU oracle.  python tools/lba_dmma_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from orbx import synth  # noqa: E402
from orbx.optimizer import Optimizer  # noqa: E402

prob = synth.lba_problem(0, n_kf=20, n_pts=3000, stereo=False, n_fixed=1)
op = Optimizer(max_keyframes=32, max_points=4096, max_edges=20000)
op.LocalBundleAdjustment(prob)
ms, Hs, bs = op.build_schur_timed(prob, 100.0, reps=1, want_system=True)
ms, _, _ = op.build_schur_timed(prob, 100.0, reps=200)
ms2, _, _ = op.build_schur_timed(prob, 100.0, reps=200)
from oracle import oracle_py as O  # noqa: E402
ref = O.lba_solve(prob, want_system=True)
print("variant", "dmma" if os.environ.get("ORBX_LBA_DMMA") else "scalar", "schur_build_us %.2f %.2f" % (1e3 * ms / 200, 1e3 * ms2 / 200))
op.close()

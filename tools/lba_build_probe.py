"""runs orbx_lba_build_schur_timed on the C3 problem (under ncu: per-kernel durations of one Schur build)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from orbx import synth
from orbx.optimizer import Optimizer
p = synth.lba_problem(0, n_kf=20, n_pts=3000, stereo=False, n_fixed=1)
o = Optimizer(32, 4096, 20000)
for reps in (1, 50):
    ms, _, _ = o.build_schur_timed(p, 100.0, reps=reps)
    print("reps", reps, "us per build", 1e3 * ms / reps)

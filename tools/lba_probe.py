"""times orbx_lba_solve_host on the C3 problem (used under ncu to see the kernel durations)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200")]
from orbx import synth
from orbx.optimizer import Optimizer
p = synth.lba_problem(0, n_kf=20, n_pts=3000, stereo=False, n_fixed=1)
o = Optimizer(32, 4096, 20000)
for i in range(3):
    t = time.perf_counter(); r = o.LocalBundleAdjustment(p); print("ms", 1e3 * (time.perf_counter() - t), r["trials"], o.last_launches(), {k: round(v) for k, v in o.phase_us().items()})

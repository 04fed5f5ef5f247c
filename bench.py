#!/usr/bin/env python
"""bench.py — frames/s of the ORB hot path on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl orbx|reference]

A "step" is one pass of the hot path over one batch of B synthetic 640x480 frames (1000 features, 8 levels):
ORBextractor::operator() on every frame, then ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th=7) of every
frame against its predecessor (the predecessor is the same scene shifted by (13, 7) px, its
map points lie on a plane at z = 4 m and the current pose is the matching translation).
  value     frames/s with the batch already resident in HBM (CUDA events on the launching stream)
  e2e       frames/s with HOST buffers: pinned frames + last-frame points in, H2D copies, the C-ABI calls, D2H of
            keypoints + descriptors + counts + matches, stream-synchronised every step
  roofline  dominant kernel: algorithmic bytes per launch / its mean CUDA-event duration vs measured HBM peak
  cpu_baseline  the CPU oracle (a dependency-free port of the reference path) on this box's host cores
N > 1 (torchrun, one rank per GPU): frames are independent, so every rank runs its own batch (weak scaling, no
data-path collective); time = max over ranks.  --impl reference times the CPU path only (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "active-orb-slam2_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

# rank 0 prints exactly one line on stdout.  Native libraries write there too (NCCL prints its version banner at every
# NCCL_DEBUG level from VERSION up, and this image sets VERSION), so file descriptor 1 is pointed at stderr for the whole run
# and the JSON line goes to the saved original descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


import numpy as np  # noqa: E402

W, H, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH = 640, 480, 1000, 1.2, 8, 20, 7
METRIC = "frames/sec ORB extract+match (640x480, 1000 kpts)"
# SURVEY.md §8(d) / DESIGN.md: compulsory bytes per VGA frame of each stage (level sizes of the 8-level pyramid)
PYR_PADDED = 1158012
PYR_INTERIOR = 950532
ALGO_BYTES, FRAME_ALGO_BYTES = {}, 0


def set_workload(w, h, nfeat):
    """byte model of SURVEY §8(d) for a w x h frame with nfeat features (pyramid sizes by the reference's float formulas)"""
    global W, H, NFEAT, PYR_PADDED, PYR_INTERIOR, FRAME_ALGO_BYTES
    W, H, NFEAT = w, h, nfeat
    sf, PYR_PADDED, PYR_INTERIOR = np.float32(1.0), 0, 0
    for l in range(NLEVELS):
        if l:
            sf = np.float32(np.float64(sf) * np.float64(np.float32(SCALE)))          # mvScaleFactor[i] = mvScaleFactor[i-1] * scaleFactor
        inv = np.float32(1.0) / sf
        lw, lh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))   # cvRound(cols * inv), ORBextractor.cc:1112
        PYR_PADDED += (lw + 38) * (lh + 38)
        PYR_INTERIOR += lw * lh
    ALGO_BYTES.update({
        "match": 2 * NFEAT * (32 + 24) + NFEAT * 8,    # both descriptor sets + points/keypoints in, match array out (~104 KB)
        "pyramid": W * H + PYR_PADDED,                 # read the frame, write the padded pyramid
        "fast": PYR_INTERIOR,                          # read every level once (+ a few KB of candidates)
        "quadtree": 0,
        "blur": 2 * PYR_INTERIOR,                      # read + write every level
        "describe": NFEAT * (749 + 512 + 60),          # patch + 512 samples + outputs per keypoint
    })
    FRAME_ALGO_BYTES = W * H + PYR_PADDED + NFEAT * 60   # 1,525,212 B for VGA / 1000


set_workload(W, H, NFEAT)
assert (PYR_PADDED, PYR_INTERIOR, FRAME_ALGO_BYTES) == (1158012, 950532, 1525212)

def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def workload_config(batch):
    """the `config` of the JSON line: identical in both arms (orbx and --impl reference), nothing run-dependent in it"""
    return {"workload": "C2 extract+match: %dx%d G-rect frames, %d features, 8 levels, th 20/7; each frame matched against its predecessor "
                        "(same scene shifted by (13, 7) px, map points on a plane at z = 4 m) with SearchByProjection(Cur, Last, th=7)" % (W, H, NFEAT),
            "batch": batch,
            "l2": "inputs larger than L2: every step reads a fresh batch from a pool of 8 x batch frames (157 MB at batch 64 > 126 MB L2)"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch at batch 64 of every kernel of the step, from the committed summary of the
    round's `ncu --set full` capture (profiles/r2_traffic.json, written by tools/ncu_summary.py traffic <report>); None if absent"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:
        return None


def binary_identity():
    """which liborbx.so ran: its hash, and whether it was built from the sources in this tree (build() records their hash next to it)"""
    import hashlib
    pkg = os.path.join(ROOT, "active-orb-slam2_b200")
    out = {}
    try:
        out["liborbx_sha16"] = hashlib.sha256(open(os.path.join(pkg, "liborbx.so"), "rb").read()).hexdigest()[:16]
        sys.path.insert(0, ROOT)
        from __graft_entry__ import source_hash
        out["source_sha16"] = source_hash()
        out["built_from_sha16"] = json.load(open(os.path.join(pkg, "liborbx.build.json"))).get("source_sha16")
        out["binary_matches_source"] = out["source_sha16"] == out["built_from_sha16"]
    except Exception as ex_:                             # noqa: BLE001
        out["error"] = str(ex_)[:120]
    return out


def bind_rank_to_cores(local_rank, world):
    """one slice of the host cores per rank: the feeding thread of a rank and its pinned allocations stay on its own cores"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(len(cores) // max(world, 1), 1)
        mine = cores[(local_rank * per) % len(cores):][:per] or cores
        os.sched_setaffinity(0, mine)
    except Exception:                                    # noqa: BLE001
        pass


def make_frames(n, seed0=0):
    """n distinct synthetic frames: a few G-rect bases and cheap integer variants of them"""
    from orbx import synth
    nbase = min(n, 16)
    base = [synth.g_rect(seed0 + i, W, H) for i in range(nbase)]
    out = np.empty((n, H, W), np.uint8)
    for i in range(n):
        b = base[i % nbase]
        k = i // nbase
        out[i] = np.roll(b, (7 * k, 13 * k), (0, 1)) if k else b
    return out


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_path(frames, seconds_budget, threads, nbase=16):
    """the same extract+match workload on the CPU oracle, `threads` host threads (ctypes releases the GIL), one
    extractor per thread; frames[k * nbase + b] is variant k of base image b.  Returns (fps, frames_done, seconds)."""
    from oracle import oracle_py as O
    from orbx import synth
    O.lib()
    done = [0] * threads
    stop_at = [None]
    nvar = max(len(frames) // nbase, 1)
    fx, fy, cx, cy, bf, bb = synth.TUM1_K
    Z = 4.0
    sf = synth.scale_factors(NLEVELS, SCALE)
    R = np.eye(3, dtype=np.float32)
    tshift = np.array([13.0 * Z / fx, 7.0 * Z / fy, 0.0], np.float32)

    def work(t):
        ex = O.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
        b, k, prev = t % nbase, 0, None
        while time.perf_counter() < stop_at[0]:
            kp, de = ex(frames[(k * nbase + b) % len(frames)])
            last = (kp, de) if prev is None else prev
            pts = np.zeros(len(last[0]), O.LAST_POINT_DTYPE)
            pts["x"] = (last[0]["x"].astype(np.float64) - cx) / fx * Z
            pts["y"] = (last[0]["y"].astype(np.float64) - cy) / fy * Z
            pts["z"], pts["angle"], pts["octave"], pts["valid"], pts["blocks"] = Z, last[0]["angle"], last[0]["octave"], 1, 1
            cur = dict(keys_un=kp, desc=de, u_right=None, claimed=None, bounds=(0.0, 0.0, float(W), float(H)), K=synth.TUM1_K,
                       scale_factors=sf)
            O.search_by_projection_frame(cur, pts, last[1], R, tshift if prev is not None else np.zeros(3, np.float32), False, False,
                                         7.0, True)
            prev = (kp, de)
            k += 1
            if k == nvar:
                k, prev = 0, None
            done[t] += 1

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    stop_at[0] = t0 + seconds_budget
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, sum(done), dt


def _reference_worker(idx, seconds, warm_seconds, shape, ready, go, results, nbase=16, nvar=4):
    """child process: the REFERENCE'S OWN ORBextractor + ORBmatcher (src/ORBextractor.cc, ORBmatcher.cc, Frame.cc, MapPoint.cc compiled
    unmodified into oracle/_ref/*.so, see oracle/cvmini/cvmini.hpp) on the C2 workload: extract, then SearchByProjection(Cur, Last,
    th=7) against the predecessor.  One process per host core (the reference library's allocator is per process)."""
    global W, H, NFEAT
    W, H, NFEAT = shape
    from oracle import oracle_py as O
    from orbx import synth
    fx, fy, cx, cy, bf, bb = synth.TUM1_K
    Z = 4.0
    sf = synth.scale_factors(NLEVELS, SCALE)
    R = np.eye(3, dtype=np.float32)
    tshift = np.array([13.0 * Z / fx, 7.0 * Z / fy, 0.0], np.float32)
    base = synth.g_rect(idx % nbase, W, H)
    frames = [np.roll(base, (7 * k, 13 * k), (0, 1)) if k else base for k in range(nvar)]
    O.ref_extractor_lib(); O.ref_matcher_lib()

    def run(until):
        k, prev, n = 0, None, 0
        while time.perf_counter() < until:
            ex = O.RefExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)      # the handle's arena rewinds when it is closed
            kp, de = ex(frames[k])
            ex.close()
            last = (kp, de) if prev is None else prev
            pts = np.zeros(len(last[0]), O.LAST_POINT_DTYPE)
            pts["x"] = (last[0]["x"].astype(np.float64) - cx) / fx * Z
            pts["y"] = (last[0]["y"].astype(np.float64) - cy) / fy * Z
            pts["z"], pts["angle"], pts["octave"], pts["valid"], pts["blocks"] = Z, last[0]["angle"], last[0]["octave"], 1, 1
            cur = dict(keys_un=kp, desc=de, u_right=None, claimed=None, bounds=(0.0, 0.0, float(W), float(H)), K=synth.TUM1_K,
                       scale_factors=sf)
            t = tshift if prev is not None else np.zeros(3, np.float32)
            O.ref_search_by_projection_frame(cur, pts, last[1], R, t, R, t, 0, 7.0, 0.9, True)
            prev = (kp, de)
            k += 1
            if k == nvar:
                k, prev = 0, None
            n += 1
        return n

    run(time.perf_counter() + warm_seconds)
    ready.put(idx)
    go.wait()
    t0 = time.perf_counter()
    n = run(t0 + seconds)
    results.put((idx, n, time.perf_counter() - t0))


def cpu_path_reference(seconds_budget, procs, warm_seconds=1.0):
    """the C2 workload on the reference's own compiled sources (oracle/_ref), `procs` worker processes.  Returns (fps, frames,
    seconds) or None when oracle/_ref was not built (the reference tree exists only in the build container) or a worker failed."""
    try:
        from oracle import oracle_py as O
        if not (os.path.exists(O.REF_EXTRACTOR_SO) and os.path.exists(O.REF_MATCHER_SO)):
            return None
        import multiprocessing as mp
        ctx = mp.get_context("spawn")                    # the parent may hold a CUDA context
        ready, results, go = ctx.Queue(), ctx.Queue(), ctx.Event()
        ps = [ctx.Process(target=_reference_worker, args=(i, seconds_budget, warm_seconds, (W, H, NFEAT), ready, go, results), daemon=True)
              for i in range(procs)]
        for p_ in ps:
            p_.start()
        for _ in ps:
            ready.get(timeout=120)
        go.set()
        got = [results.get(timeout=seconds_budget + 120) for _ in ps]
        for p_ in ps:
            p_.join(timeout=30)
        return sum(n / dt for _, n, dt in got), sum(n for _, n, _ in got), max(dt for _, _, dt in got)
    except Exception as ex_:                             # noqa: BLE001
        print("reference-library CPU path unavailable (%s): falling back to the oracle port" % str(ex_)[:200], file=sys.stderr)
        return None


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total_budget = max(10.0, min(60.0, 0.2 * args.steps))
    r = cpu_path_reference(total_budget, cores, 3.0 if args.warmup else 0.5)
    if r is not None:
        v, n_frames, total = r
        budget = total / max(args.steps, 1)
        sample = ("%d steps x %.1f s of G-rect VGA frames (extract + SearchByProjection vs predecessor) on %d host processes, the reference's "
                  "own ORBextractor / ORBmatcher / Frame / MapPoint sources (oracle/_ref)" % (args.steps, budget, cores))
        emit({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.batch),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's src/ORBextractor.cc, ORBmatcher.cc, Frame.cc and MapPoint.cc compiled unmodified (oracle/Makefile, target "
                    "ref); the OpenCV primitives underneath (resize, copyMakeBorder, FAST, GaussianBlur) are the stand-in's scalar "
                    "restatements, not OpenCV's SIMD kernels, and every frame pays the shim's Frame / MapPoint construction",
        })
        return
    frames = make_frames(64)
    # K "steps" are K equal slices of ONE timed run with persistent worker threads (a step is a bounded sample of the
    # workload; starting threads and extractors per slice would only measure that overhead)
    cpu_path(frames, 3.0 if args.warmup else 0.5, cores)
    total_budget = max(10.0, min(60.0, 0.2 * args.steps))
    v, n_frames, total = cpu_path(frames, total_budget, cores)
    budget = total / max(args.steps, 1)
    sample = "%d steps x %.1f s of G-rect VGA frames (extract + SearchByProjection vs predecessor) on %d host threads, one oracle extractor per thread" % (args.steps, budget, cores)
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.batch),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference itself cannot be built here (needs OpenCV/Eigen/Pangolin); this is the dependency-free C oracle of its "
                "path with scalar restatements of OpenCV's SIMD primitives",
    })


def bench_stereo(local_rank, with_cpu, n_pairs=32, reps=20):
    """Frame::ComputeStereoMatches on n_pairs rectified VGA pairs, device-resident: one extractor run over
    L0 R0 L1 R1 ..., then the two stereo kernels read keypoints, descriptors and both pyramids in place."""
    import torch
    from orbx import synth
    from orbx.extractor import ORBextractor
    from orbx.stereo import StereoMatcher, StereoSide
    world = synth.stereo_world(0, W, H)
    frames = []
    for p in range(n_pairs):
        frames += [world.render(0.02 * p, 0.01, 0.001 * p), world.render(0.02 * p, 0.01, 0.001 * p, right=True)]
    bf, b = world.bf, world.bf / world.fx
    ex = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=W, max_height=H, max_batch=2 * n_pairs, device=local_rank)
    cap = ex.capacity
    d_img = torch.from_numpy(np.stack(frames)).cuda()
    d_kps = torch.zeros((2 * n_pairs, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((2 * n_pairs, cap, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2 * n_pairs, dtype=torch.int32, device="cuda")
    d_ur = torch.zeros((n_pairs, cap), dtype=torch.float32, device="cuda")
    d_dp = torch.zeros((n_pairs, cap), dtype=torch.float32, device="cuda")
    d_kept = torch.zeros(n_pairs, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    ex.run_device(d_img.data_ptr(), W * H, 2 * n_pairs, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), st.cuda_stream)
    left = StereoSide(d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), 2 * cap, 2, ex._h, 0, 2, cap)
    right = StereoSide(d_kps.data_ptr() + 28 * cap, d_desc.data_ptr() + 32 * cap, d_cnt.data_ptr() + 4, 2 * cap, 2, ex._h, 1, 2, cap)
    sm = StereoMatcher(max_keypoints=cap, max_pairs=n_pairs, device=local_rank)
    run = lambda: sm.matches_device(left, right, n_pairs, bf, b, d_ur.data_ptr(), d_dp.data_ptr(), cap, d_kept.data_ptr(), st.cuda_stream)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        run()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kept = d_kept.cpu().numpy()
    out = {"config": "%d rectified 640x480 pairs of the 3-plane world (z = 2, 4, 8 m), 1000 features per image, device-resident" % n_pairs,
           "pairs_per_s": n_pairs / (ms * 1e-3), "ms_per_batch": ms, "kernel_launches_per_batch": sm.last_launches(),
           "associations_per_pair": float(kept.mean()), "api": "orbx_stereo_matches_device after orbx_extractor_run_device (CUDA events)"}
    if with_cpu:
        from oracle import oracle_py as O
        cnt = d_cnt.cpu().numpy()
        kps = d_kps.cpu().numpy().view(O.KP_DTYPE).reshape(2 * n_pairs, cap)
        desc = d_desc.cpu().numpy()
        sc, isc = ex.GetScaleFactors(), ex.GetInverseScaleFactors()
        t_cpu, n_cpu = 0.0, 0
        for p in range(min(n_pairs, 8)):
            pl = [ex.pyramid_level(l, 2 * p) for l in range(NLEVELS)]
            pr = [ex.pyramid_level(l, 2 * p + 1) for l in range(NLEVELS)]
            a = (kps[2 * p, :cnt[2 * p]], desc[2 * p, :cnt[2 * p]], kps[2 * p + 1, :cnt[2 * p + 1]], desc[2 * p + 1, :cnt[2 * p + 1]])
            t0 = time.perf_counter()
            r = O.stereo_matches(a[0], a[1], a[2], a[3], pl, pr, sc, isc, bf, b)
            t_cpu += time.perf_counter() - t0
            n_cpu += 1
            assert r["kept"] == int(kept[p]), "stereo: GPU and oracle disagree on pair %d" % p
        out["cpu_pairs_per_s"] = n_cpu / t_cpu
        out["cpu"] = "C oracle, 1 thread, %d pairs (pyramids already built); same association counts as the GPU" % n_cpu
    sm.close()
    ex.close()
    return out


def bench_pose(local_rank, with_cpu, n_frames=64, n_obs=500, reps=10):
    """Optimizer::PoseOptimization on n_frames frames of n_obs observations, host buffers in and out."""
    from orbx import synth
    from orbx.optimizer import PoseOptimizer
    probs = [synth.pose_problem(1000 + f, n=n_obs) for f in range(n_frames)]
    po = PoseOptimizer(max_observations=n_frames * n_obs, max_frames=n_frames, device=local_rank)
    rs = po.PoseOptimization(probs)
    t0 = time.perf_counter()
    for _ in range(reps):
        rs = po.PoseOptimization(probs)
    dt = (time.perf_counter() - t0) / reps
    out = {"config": "%d frames x %d observations (60 %% stereo, 15 %% outliers), 4 rounds x optimize(10)" % (n_frames, n_obs),
           "frames_per_s": n_frames / dt, "ms_per_batch": 1e3 * dt, "kernel_launches_per_batch": po.last_launches(),
           "lm_trials_per_frame": float(np.mean([r["trials"] for r in rs])),
           "api": "orbx_pose_optimize_host (host buffers in and out, synchronous; includes packing in Python)"}
    if with_cpu:
        from oracle import oracle_py as O
        t0 = time.perf_counter()
        for f in range(8):
            ref = O.pose_optimize(probs[f])
            assert np.array_equal(ref["outlier"], rs[f]["outlier"]), "pose: GPU and oracle disagree on frame %d" % f
        out["cpu_frames_per_s"] = 8 / (time.perf_counter() - t0)
        out["cpu"] = "C oracle (g2o restated), 1 thread, 8 frames; identical outlier sets"
    po.close()
    return out


def bench_bow(local_rank, with_cpu, n_frames=64, n_feat=1000, reps=20):
    """DBoW2 transform of n_frames x n_feat descriptors against a full 10-ary, 6-level vocabulary (1.1 M nodes, the
    shape of ORBvoc; random descriptors and weights because the reference's vocabulary file does not travel to the GPU box)."""
    import torch
    from orbx.vocabulary import ORBVocabulary, tree_from_parents
    rng = np.random.default_rng(0)
    k, L = 10, 6
    n = sum(k ** l for l in range(1, L + 1))
    parent = (np.arange(1, n + 1, dtype=np.int64) - 1) // k          # breadth-first ids: children of a node are consecutive
    first_leaf = n - k ** L
    is_leaf = np.arange(n) >= first_leaf
    tree = tree_from_parents(parent.astype(np.int32), rng.integers(0, 256, (n, 32), dtype=np.uint8), rng.uniform(0.1, 9.0, n).astype(np.float32),
                             is_leaf, k, L)
    voc = ORBVocabulary(tree, max_features=n_feat, device=local_rank)
    desc = rng.integers(0, 256, (n_frames, n_feat, 32), dtype=np.uint8)
    d_desc = torch.from_numpy(desc).cuda()
    d_cnt = torch.full((n_frames,), n_feat, dtype=torch.int32, device="cuda")
    d_word = torch.zeros((n_frames, n_feat), dtype=torch.int32, device="cuda")
    d_node = torch.zeros((n_frames, n_feat), dtype=torch.int32, device="cuda")
    d_wt = torch.zeros((n_frames, n_feat), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    run = lambda: voc.transform_device(4, d_desc.data_ptr(), d_cnt.data_ptr(), 1, n_feat, n_feat, n_frames, d_word.data_ptr(), d_node.data_ptr(),
                                       d_wt.data_ptr(), st.cuda_stream)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        run()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = {"config": "%d frames x %d descriptors, vocabulary k = 10, L = 6 (%d nodes, %.0f MB of node descriptors), levelsup = 4, device-resident" % (
               n_frames, n_feat, n + 1, 32 * n / 1e6),
           "frames_per_s": n_frames / (ms * 1e-3), "ms_per_batch": ms, "kernel_launches_per_batch": voc.last_launches(),
           "api": "orbx_vocabulary_transform_device (CUDA events)"}
    if with_cpu:
        from oracle import oracle_py as O
        t0 = time.perf_counter()
        for f in range(4):
            w_, n_, _ = O.bow_transform(tree, desc[f], 4)
            assert np.array_equal(w_, d_word[f].cpu().numpy()) and np.array_equal(n_, d_node[f].cpu().numpy()), "bow: GPU and oracle disagree"
        out["cpu_frames_per_s"] = 4 / (time.perf_counter() - t0)
        out["cpu"] = "C oracle, 1 thread, 4 frames; identical word and node ids"
    voc.close()
    return out


def bench_sequence(local_rank, with_cpu, n_frames=40):
    """Config C4 in miniature: single-stream stereo tracking, one frame at a time through the host entry points, as the
    adapter would call them from Tracking: extract L + R -> ComputeStereoMatches -> SearchByProjection(Cur, Last) -> PoseOptimization.
    Latency-bound by design (batch 1, ~20 launches and ~15 small copies per frame); tools/replay.py is the chain."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import replay
    seq = replay.StereoSequence(seed=1)
    imgs = [seq.images(t) for t in range(n_frames)]
    seq.images = lambda t: imgs[t]                        # rendering is not part of the measurement
    gpu = replay.GpuBackend(device=local_rank)
    rng = np.random.default_rng(7)
    last = None
    for t in range(3):                                    # warm-up
        last = replay.track_frame(gpu, seq, t, last, rng)
    for k in gpu.seconds:
        gpu.seconds[k] = 0.0
    t0 = time.perf_counter()
    inl = []
    for t in range(3, n_frames):
        last = replay.track_frame(gpu, seq, t, last, rng)
        inl.append(last["n_inliers"])
    dt = time.perf_counter() - t0
    err = float(np.linalg.norm(last["Tcw"][:3, 3] - seq.true_pose(n_frames - 1)[:3, 3]))
    out = {"config": "%d rendered 640x480 stereo frames, 1000 features per image, camera translating 2 cm per frame; batch 1, host entry points" % (n_frames - 3),
           "frames_per_s": (n_frames - 3) / dt, "ms_per_frame": 1e3 * dt / (n_frames - 3), "inliers_per_frame": float(np.mean(inl)),
           "final_position_error_m": err,
           "ms_per_frame_by_call": {k: 1e3 * v / (n_frames - 3) for k, v in gpu.seconds.items()},
           "api": "orbx_extractor_run_host x2, orbx_stereo_matches_host, orbx_match_projection_frame_host, orbx_pose_optimize_host (+ numpy bookkeeping)"}
    gpu.close()
    out["one_call_per_frame"] = bench_single_stream(local_rank, seq, imgs, n_frames)
    if with_cpu:
        orc = replay.OracleBackend()
        rng = np.random.default_rng(7)
        last = None
        t0 = time.perf_counter()
        for t in range(8):
            last = replay.track_frame(orc, seq, t, last, rng)
        out["cpu_frames_per_s"] = 8 / (time.perf_counter() - t0)
        out["cpu"] = "the same chain on the C oracle, 1 thread (the reference runs the two extractors on two threads), 8 frames"
    return out


def bench_single_stream(local_rank, seq, imgs, n_frames, reps=6):
    """The same single-stream chains with ONE C-ABI call per frame (orbx_sequences_step_host on a handle of one sequence): the frame's
    host image(s) and predicted pose in; keypoints, descriptors, mvuRight / mvDepth, match array, optimised pose and mvbOutlier out;
    the last frame stays on the device.  Strictly serial: a frame is submitted after the previous one has been read on the host."""
    import replay
    from orbx import synth
    from orbx.sequences import Sequences
    K = seq.K
    res = {}
    # (1) stereo chain: extract L + R -> ComputeStereoMatches -> SearchByProjection(Cur, Last) -> PoseOptimization
    sq = Sequences(1, seq.w, seq.h, K, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, stereo=True, th=7.0, mono=False, device=local_rank, pose=True)
    o = sq.alloc_outputs()
    pairs = [np.stack(imgs[t]) for t in range(n_frames)]
    rng = np.random.default_rng(7)
    order = list(range(n_frames)) + list(range(n_frames - 2, 0, -1))       # there and back again: every step moves by one frame
    Tl, n_done, inl, errs = None, 0, [], []

    def quat_to_T(p):
        """Converter::toCvMat(SE3Quat): Eigen's toRotationMatrix, narrowed to float"""
        x, y, z, w_ = p[0], p[1], p[2], p[3]
        T = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w_), 2 * (x * z + y * w_), p[4]],
                      [2 * (x * y + z * w_), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w_), p[5]],
                      [2 * (x * z - y * w_), 2 * (y * z + x * w_), 1 - 2 * (x * x + y * y), p[6]], [0, 0, 0, 1]], np.float32)
        return T
    t0 = None
    for r_ in range(reps):
        for t in order:
            if r_ == 1 and t0 is None:                                       # the first pass is the warm-up
                t0 = time.perf_counter()
            Tcw = seq.true_pose(t).copy()
            Tcw[:3, 3] += rng.normal(0, 0.01, 3).astype(np.float32)          # motion-model guess: the true pose off by ~1 cm / 0.2 deg
            Tcw[:3, :3] = (synth._rot(1, np.deg2rad(rng.normal(0, 0.2))) @ Tcw[:3, :3].astype(np.float64)).astype(np.float32)
            if Tl is not None:
                sq.set_last_poses(Tl[:3])                                    # the optimised pose of the last frame, like Tracking's mLastFrame
            sq.step(pairs[t], Tcw[:3], o)
            Tl = quat_to_T(o["pose"][0]) if n_done else seq.true_pose(t)
            n_done += 1
            if t0 is not None:
                inl.append(int(o["n_inliers"][0]))
                errs.append(float(np.linalg.norm(Tl[:3, 3] - seq.true_pose(t)[:3, 3])))
    dt = time.perf_counter() - t0
    n_timed = (reps - 1) * len(order)
    res["stereo_chain"] = {"frames_per_s": n_timed / dt, "ms_per_frame": 1e3 * dt / n_timed, "inliers_per_frame": float(np.mean(inl)),
                           "odometry_drift_m": float(errs[-1]), "frames_tracked": len(errs),
                           "drift_note": "frame-to-frame odometry without a map: the position error accumulates (~0.1-0.4 mm per frame); the same chain "
                                         "call by call gives the same poses to 1e-6",
                           "kernel_launches_per_frame": sq.last_launches(),
                           "api": "orbx_sequences_step_host, n_sequences = 1, stereo = 1, pose = 1 (+ orbx_sequences_set_last_poses)"}
    sq.close()
    # (2) monocular extract + SearchByProjection(Cur, Last): the chain BASELINE.json's >= 3000 frames/s target names, one frame at a time
    sq = Sequences(1, W, H, synth.TUM1_K, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, stereo=False, th=7.0, mono=False, const_depth=4.0, device=local_rank)
    o = sq.alloc_outputs()
    fx, fy = synth.TUM1_K[0], synth.TUM1_K[1]
    base = synth.g_rect(5, W, H)
    shifts = [0, 1, 2, 3, 4, 3, 2, 1]
    frames = [np.ascontiguousarray(np.roll(base, (7 * k, 13 * k), (0, 1)))[None] for k in shifts]
    poses = []
    for k in shifts:
        T = np.zeros((3, 4), np.float32)
        T[:, :3] = np.eye(3)
        T[:, 3] = (13.0 * k * 4.0 / fx, 7.0 * k * 4.0 / fy, 0.0)
        poses.append(T)
    for i in range(16):
        sq.step(frames[i % 8], poses[i % 8], o)
    n_mono = 400
    t0 = time.perf_counter()
    nm = 0
    for i in range(n_mono):
        sq.step(frames[i % 8], poses[i % 8], o)
        nm += int(o["nmatches"][0])
    dt = time.perf_counter() - t0
    res["mono_extract_match"] = {"frames_per_s": n_mono / dt, "ms_per_frame": 1e3 * dt / n_mono, "matches_per_frame": nm / n_mono,
                                 "kernel_launches_per_frame": sq.last_launches(),
                                 "api": "orbx_sequences_step_host, n_sequences = 1 (host image + pose in; keypoints, descriptors, match array out)"}
    sq.close()
    return res


def bench_c4(local_rank, with_cpu, n_frames=2000):
    """BASELINE.json config 4: a TUM-RGBD-shaped sequence (640x480, 2000 frames) through the Tracking + LocalMapping call mix on a live map
    (tools/replay_c4.py): per frame extract x2, ComputeStereoMatches, SearchByProjection(Cur, Last), PoseOptimization, isInFrustum over
    the local map, SearchByProjection(Frame, local map points), PoseOptimization; every 10th frame a keyframe: new map points, vocabulary
    transform, SearchForTriangulation against up to 10 earlier keyframes, LocalBundleAdjustment over the last 10 keyframes.  Single
    stream, batch 1, host entry points; the numpy map bookkeeping between the calls is inside the wall time."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import replay_c4
    tree = replay_c4.make_vocabulary()
    seq = replay_c4.LoopSequence(seed=1, n_poses=40)
    for t in range(40):
        seq.images(t)                                     # rendering is not part of the measurement
    gpu = replay_c4.GpuBackend(tree, device=local_rank)
    rp = replay_c4.Replay(gpu, seq, kf_every=10)
    for t in range(20):                                   # warm-up (two keyframes)
        rp.step(t)
    for k in gpu.seconds:
        gpu.seconds[k] = 0.0
    t0 = time.perf_counter()
    for t in range(20, 20 + n_frames):
        rp.step(t)
    dt = time.perf_counter() - t0
    out = {"config": "%d frames of a 640x480 stereo sequence (camera moving there and back over 40 rendered views, 2 cm per frame), 1000 features "
                     "per image, a keyframe every 10th frame; batch 1, host entry points" % n_frames,
           "frames_per_s": n_frames / dt, "ms_per_frame": 1e3 * dt / n_frames, "realtime_factor_at_30fps": n_frames / dt / 30.0,
           "ms_per_frame_by_call": {k: 1e3 * v / n_frames for k, v in gpu.seconds.items()},
           "ms_per_frame_in_library_calls": 1e3 * sum(gpu.seconds.values()) / n_frames}
    out.update(rp.summary())
    out["api"] = ("orbx_extractor_run_host x2, orbx_stereo_matches_extractors_host, orbx_match_projection_frame_host, orbx_pose_optimize_host x2, "
                  "orbx_frustum_host, orbx_match_projection_points_host; per keyframe orbx_vocabulary_transform_host, orbx_match_buckets_host "
                  "(SearchForTriangulation) x <= 10, orbx_lba_solve_host")
    gpu.close()
    if with_cpu:
        orc = replay_c4.OracleBackend(tree)
        rq = replay_c4.Replay(orc, seq, kf_every=10)
        t0 = time.perf_counter()
        n_cpu = 41
        for t in range(n_cpu):
            rq.step(t)
        out["cpu_frames_per_s"] = n_cpu / (time.perf_counter() - t0)
        out["cpu"] = "the same replay on the C oracle, 1 thread, %d frames (5 keyframes)" % n_cpu
    return out


def bench_c5(local_rank, rank, world, n_seq_total=64, steps=40, w=1241, h=376, nfeat=2000):
    """BASELINE.json config 5: 64 independent KITTI-stereo-shaped sequences (1241x376, 2000 features; intrinsics of
    Examples/Stereo/KITTI00-02.yaml) sharded over the ranks, sequence s -> rank s mod N, through the per-frame stereo chain (extract L + R ->
    ComputeStereoMatches -> SearchByProjection(Cur, Last)) with HOST buffers in and out (orbx_sequences_step_begin/_end, two handles in
    flight per rank).  Strong scaling: the 64 sequences are fixed, every rank takes its share.  Returns (pairs done, seconds, matches)."""
    import torch
    from orbx import synth
    from orbx.sequences import Sequences
    mine = [s_ for s_ in range(n_seq_total) if s_ % world == rank]
    if not mine:
        return 0, 0.0, 0
    K = (718.856, 718.856, 607.1928, 185.2157, 386.1448, 386.1448 / 718.856)   # Camera.fx, fy, cx, cy, bf (KITTI00-02.yaml:8-25)
    nh = 2 if len(mine) >= 2 else 1
    groups = [mine[i::nh] for i in range(nh)]
    # a sequence = a G-rect scene on a fronto-parallel plane at z = bf / d (disparity d px): the right image is the left shifted by d, the
    # camera translates sideways by 9 px per frame (there and back again)
    SH = [0, 1, 2, 3, 4, 3, 2, 1]
    pool = {}

    def scene(s_):
        if s_ % 8 not in pool:
            pool[s_ % 8] = synth.g_rect(1000 + s_ % 8, w, h, nrect=1800)
        return pool[s_ % 8]

    handles = []
    for grp in groups:
        n = len(grp)
        sq = Sequences(n, w, h, K, nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, stereo=True, th=7.0, mono=False, device=local_rank)
        imgs = torch.empty((len(SH), 2 * n, h, w), dtype=torch.uint8).pin_memory()
        poses = np.zeros((len(SH), n, 3, 4), np.float32)
        for j, s_ in enumerate(grp):
            d = 12 + 4 * (s_ % 5)                                            # disparity in pixels -> depth bf / d
            z = K[4] / d
            sc = scene(s_)
            for t, k in enumerate(SH):
                left = np.roll(sc, 9 * k + 3 * (s_ // 8), 1)
                imgs[t, 2 * j] = torch.from_numpy(left)
                imgs[t, 2 * j + 1] = torch.from_numpy(np.roll(left, -d, 1))
                poses[t, j, :, :3] = np.eye(3)
                poses[t, j, 0, 3] = 9.0 * k * z / K[0]
        handles.append((sq, imgs.numpy(), poses, sq.alloc_outputs(lambda shape, dtype: torch.zeros(shape, dtype={np.uint8: torch.uint8, np.int32: torch.int32, np.float32: torch.float32, np.float64: torch.float64}[dtype]).pin_memory().numpy())))
    busy = [False] * nh
    nm = 0

    def run(n_steps, count):
        nonlocal nm
        for i in range(n_steps):
            for hi, (sq, imgs, poses, o) in enumerate(handles):
                if busy[hi]:
                    sq.end()
                    if count:
                        nm += int(o["nmatches"].sum())
                sq.begin(imgs[i % len(SH)], poses[i % len(SH)], o)
                busy[hi] = True
        for hi, (sq, imgs, poses, o) in enumerate(handles):
            if busy[hi]:
                sq.end()
                busy[hi] = False
                if count:
                    nm += int(o["nmatches"].sum())

    run(3, False)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps, True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for sq, _, _, _ in handles:
        sq.close()
    return steps * len(mine), dt, nm


def cpu_c5(seconds_budget, threads, w=1241, h=376, nfeat=2000):
    """the config-5 chain (extract L + R, ComputeStereoMatches, SearchByProjection(Cur, Last)) on the CPU oracle, `threads` host threads,
    one sequence per thread; returns (pairs/s, pairs, seconds)"""
    from oracle import oracle_py as O
    from orbx import synth
    O.lib()
    K = (718.856, 718.856, 607.1928, 185.2157, 386.1448, 386.1448 / 718.856)
    sf = synth.scale_factors(NLEVELS, SCALE)
    done = [0] * threads
    stop_at = [None]
    scenes = [synth.g_rect(1000 + i, w, h, nrect=1800) for i in range(min(threads, 8))]
    SH = [0, 1, 2, 3, 4, 3, 2, 1]

    def work(t):
        exl, exr = O.Extractor(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH), O.Extractor(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH)
        tb = exl.tables()
        d = 12 + 4 * (t % 5)
        z = K[4] / d
        last, i = None, 0
        while time.perf_counter() < stop_at[0]:
            k = SH[i % 8]
            left = np.roll(scenes[t % len(scenes)], 9 * k, 1)
            kl, dl = exl(left)
            kr, dr = exr(np.roll(left, -d, 1))
            st = O.stereo_matches(kl, dl, kr, dr, [exl.level(l) for l in range(NLEVELS)], [exr.level(l) for l in range(NLEVELS)], tb["scale"],
                                  tb["inv_scale"], K[4], K[5])
            tx = np.float32(9.0 * k * z / K[0])
            if last is not None:
                lk, ld, lz, ltx = last
                pts = np.zeros(len(lk), O.LAST_POINT_DTYPE)
                ok = lz > 0
                pts["x"] = (lk["x"] - np.float32(K[2])) * lz / np.float32(K[0]) - ltx
                pts["y"] = (lk["y"] - np.float32(K[3])) * lz / np.float32(K[1])
                pts["z"], pts["angle"], pts["octave"], pts["valid"], pts["blocks"] = lz, lk["angle"], lk["octave"], ok, ok
                cur = dict(keys_un=kl, desc=dl, u_right=st["u_right"], claimed=None, bounds=(0.0, 0.0, float(w), float(h)), K=K, scale_factors=sf)
                O.search_by_projection_frame(cur, pts, ld, np.eye(3, dtype=np.float32), np.array([tx, 0, 0], np.float32), False, False, 7.0, True)
            last = (kl, dl, st["depth"], tx)
            i += 1
            done[t] += 1

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    stop_at[0] = t0 + seconds_budget
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, sum(done), dt


def lba_host_threads():
    """host threads that submit and collect LocalBA windows: 4 (the same at every N, so that the per-N numbers compare), fewer if this
    rank is bound to fewer cores"""
    try:
        return max(1, min(4, len(os.sched_getaffinity(0))))
    except Exception:                                    # noqa: BLE001
        return 1


def bench_lba_batched(local_rank, rank, NW=16, rounds=8):
    """batched many-window mode: NW independent C3 windows in flight on this rank, one handle + stream each, submitted and
    collected by a few host threads (a window's list building is host work, ~0.3 ms; the C calls release the GIL) -- the shape of a box
    that serves many sequences, each with its own LocalMapping thread.  Every rank runs its own windows (no data-path collective);
    returns (windows, Levenberg trials, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from orbx import synth
    from orbx.optimizer import Optimizer
    ops = [Optimizer(max_keyframes=32, max_points=4096, max_edges=20000, device=local_rank) for _ in range(NW)]
    from orbx.optimizer import pack_problem
    probs = [pack_problem(synth.lba_problem(100 + NW * rank + i, n_kf=20, n_pts=3000, stereo=False, n_fixed=1)) for i in range(NW)]   # packed once: the arrays do not change
    nt = min(lba_host_threads(), NW)

    def work(t, n_rounds):
        n = 0
        for _ in range(n_rounds):
            for i in range(t, NW, nt):
                ops[i].begin(probs[i])
            for i in range(t, NW, nt):
                n += ops[i].end()["trials"]
        return n
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(work, range(nt), [1] * nt))           # warm-up
        t0 = time.perf_counter()
        trials = sum(ex.map(work, range(nt), [rounds] * nt))
        dt = time.perf_counter() - t0
    for o_ in ops:
        o_.close()
    return rounds * NW, trials, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--impl", default="orbx", choices=["orbx", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--width", type=int, default=W, help="frame width (default: the VGA workload BASELINE.json's metric is quoted on)")
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--features", type=int, default=NFEAT)
    ap.add_argument("--profile-steps", type=int, default=0, help="ncu runs: after the warm-up, bracket this many device-resident steps with "
                    "cudaProfilerStart/Stop and exit (use with ncu --profile-from-start off)")
    ap.add_argument("--subs", type=int, default=4, help="sub-batch pipelines inside the handle of the device-resident `value` loop "
                    "(orbx_sequences_config.n_sub; 1 = the whole step on one stream)")
    ap.add_argument("--extract-only", action="store_true", help="other frame shapes: time the extractor + matcher loop only (no LocalBA / stereo / ... sections)")
    args = ap.parse_args()
    default_workload = (args.width, args.height, args.features) == (W, H, NFEAT)
    if not default_workload:
        set_workload(args.width, args.height, args.features)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    from orbx import synth
    from orbx._lib import KP_DTYPE
    from orbx.sequences import Sequences
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the orbx hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    bind_rank_to_cores(local_rank, world)
    npool = 8                                           # 8 time steps x 64 sequences x 300 KB = 157 MB of inputs > 126 MB L2
    # The workload: B independent sequences per handle advance in lockstep.  Sequence j shows G-rect scene j % 16 on a plane
    # at z = 4 m; time step s shifts it by SHIFT[s] x (13, 7) px (np.roll), i.e. the camera translates; every frame is matched against
    # its predecessor (the previous time step of the same sequence) with SearchByProjection(Cur, Last, th = 7), the predecessor's map
    # points being its own keypoints unprojected at z = 4 m (Frame::UnprojectStereo) -- all inside orbx_sequences_step_*.
    SHIFT = [0, 1, 2, 3, 4, 3, 2, 1]                   # a palindrome: cycling through the pool always moves by one step
    fx, fy, cx, cy, bf, bb = synth.TUM1_K
    Z = 4.0
    base = [synth.g_rect(100 * rank + i, W, H) for i in range(16)]
    frames_np = np.empty((npool, B, H, W), np.uint8)
    poses = np.zeros((npool, B, 3, 4), np.float32)
    for s_ in range(npool):
        for j in range(B):
            k = SHIFT[s_] + (j // 16)
            frames_np[s_, j] = np.roll(base[j % 16], (7 * k, 13 * k), (0, 1)) if k else base[j % 16]
            poses[s_, j, :, :3] = np.eye(3)
            poses[s_, j, :, 3] = (13.0 * k * Z / fx, 7.0 * k * Z / fy, 0.0)     # np.roll by (7k, 13k): the scene moves +13k px in x, +7k px in y
    host = torch.from_numpy(frames_np).pin_memory()
    dev = host.cuda()

    def new_handle(n_sub=0):
        return Sequences(B, W, H, synth.TUM1_K, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, stereo=False, th=7.0, check_ori=True, mono=False,
                         const_depth=Z, device=local_rank, n_sub=n_sub)

    sq = new_handle(args.subs)
    cap = sq.capacity
    stream = torch.cuda.Stream()                        # the device-resident loops run on this stream; the timing events are recorded on it

    def step_device(i, h=None, st=None):
        slot = i % npool
        (h or sq).step_device(dev[slot].data_ptr(), W * H, W, poses[slot], (st or stream).cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()                                      # sampled from the warm-up to the end of the end-to-end loop
    for i in range(Wm):
        step_device(i)
    torch.cuda.synchronize()
    launches_per_step = sq.last_launches() + 2          # + the two memset nodes of the match arrays
    if args.profile_steps:
        torch.cuda.cudart().cudaProfilerStart()
        for i in range(args.profile_steps):
            step_device(Wm + i)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        clocks.stop()
        sq.close()
        return
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step_device(Wm + i)
    sq.join(stream.cuda_stream)                         # (a handle with sub-batch pipelines: the stream waits for all of them)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    # host time to enqueue a step: a few steps into an empty queue, no synchronisation (the queue never fills)
    t_enq = time.perf_counter()
    for i in range(8):
        step_device(i)
    t_enq = (time.perf_counter() - t_enq) / 8
    torch.cuda.synchronize()

    # per-stage times come from a second, untimed-for-`value` loop: with stage events on, the extractor keeps its stages on one
    # stream back to back (in the loop above the blur runs on a side stream next to the quadtree kernel)
    # and the whole batch goes through each kernel in one launch (a handle created with n_sub = 1), so a stage time is one kernel's
    # duration over all B frames -- the figure the roofline line is computed from
    KP = min(K, 100)
    sq1 = new_handle(n_sub=1)
    for i in range(2):
        step_device(i, sq1)
    sq1.profile(KP)
    ev_m = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KP)]
    for i in range(KP):
        ev_m[i][0].record(stream)
        step_device(2 + i, sq1)
        ev_m[i][1].record(stream)
    torch.cuda.synchronize()
    runs, stage_ms = sq1.stage_ms()
    # everything of the step that is not the extractor: unprojection of the last frame, the projection search and its memsets
    stage_ms["match"] = sum(a.elapsed_time(b) for a, b in ev_m) * runs / KP - sum(stage_ms.values())
    sq1.profile(0)
    view = sq1.device_view()
    # the same handle without stage events: the step on ONE stream (what `value` is when the handle has no sub-batch pipelines)
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record(stream)
    for i in range(K):
        step_device(i, sq1)
    o1.record(stream)
    torch.cuda.synchronize()
    ms_one_stream = o0.elapsed_time(o1) / K
    # ---- the same step followed by PoseOptimization of every frame from the match arrays, still device-resident (the three
    # calls Tracking::TrackWithMotionModel makes per frame: extract, SearchByProjection(Cur, Last), PoseOptimization) ----
    track = None
    d_is2 = torch.from_numpy(np.float32(1.0) / (synth.scale_factors(NLEVELS, SCALE) ** 2)).cuda()      # mvInvLevelSigma2
    if rank == 0:
        from orbx.optimizer import PoseOptimizer
        pz = PoseOptimizer(max_observations=B * cap, max_frames=B, device=local_rank)
        d_pose = torch.zeros((B, 7), dtype=torch.float64, device="cuda")
        d_inl = torch.zeros(B, dtype=torch.int32, device="cuda")
        d_outkp = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")

        def step_track(i):
            step_device(i, sq1)
            pz.from_matches_device(view.jobs, B, d_is2.data_ptr(), NLEVELS, synth.TUM1_K, d_pose.data_ptr(), d_inl.data_ptr(),
                                   d_outkp.data_ptr(), cap, stream.cuda_stream)

        for i in range(3):
            step_track(i)
        torch.cuda.synchronize()
        KT = min(K, 100)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record(stream)
        for i in range(KT):
            step_track(3 + i)
        t1e.record(stream)
        torch.cuda.synchronize()
        ms_track = t0e.elapsed_time(t1e) / KT
        track = {"config": "the `value` step + Optimizer::PoseOptimization of all %d frames from the match arrays (monocular observations), device-resident" % B,
                 "frames_per_s": B / (ms_track * 1e-3), "ms_per_step": ms_track, "pose_ms_per_step": ms_track - ms_one_stream,
                 "inliers_per_frame": float(d_inl.float().mean().item()), "kernel_launches_per_step": launches_per_step + pz.last_launches(),
                 "api": "orbx_sequences_step_device + orbx_pose_from_matches_device"}
        pz.close()

    # ---- end to end with HOST buffers through the C ABI: every step hands pinned host frames and poses to orbx_sequences_step_begin
    # (upload, kernels, download of keypoints, descriptors, counts and match arrays) and reads the results after
    # orbx_sequences_step_end.  NLANES handles (B sequences each) are in flight, so the copies of one overlap the kernels of the
    # others; every byte of every step crosses the bus inside the timed region. ----
    class Lane:
        pass

    def pinned(shape, dtype):
        return torch.zeros(shape, dtype={np.uint8: torch.uint8, np.int32: torch.int32, np.float32: torch.float32}[dtype]).pin_memory().numpy()

    lanes = []
    NLANES = 4
    for li in range(NLANES):
        L = Lane()
        L.sq = sq1 if li == 0 else new_handle(n_sub=1)       # the end-to-end lanes are one-stream handles: the lanes themselves overlap
        L.sq1 = L.sq
        L.stream = torch.cuda.Stream()
        L.out = L.sq.alloc_outputs(pinned)
        L.t, L.busy = 0, False
        lanes.append(L)
    torch.cuda.synchronize()

    # device-resident again, but with two handles alternating on two streams (for comparison with `value`, which is one stream)
    def step_device_lane(i, one_stream=False):
        L = lanes[i % 2]
        step_device(L.t, L.sq1 if one_stream else L.sq, L.stream)
        L.t += 1

    for i in range(4):
        step_device_lane(i)
    torch.cuda.synchronize()
    t2_0 = torch.cuda.Event(enable_timing=True)
    t2_0.record(torch.cuda.current_stream())
    for L in lanes:
        L.stream.wait_event(t2_0)
    for i in range(K):
        step_device_lane(i)
    t2_1 = []
    for L in lanes[:2]:
        L.sq.join(L.stream.cuda_stream)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(L.stream)
        t2_1.append(ev)
    torch.cuda.synchronize()
    ms_two_lanes = max(t2_0.elapsed_time(ev) for ev in t2_1)
    if track is not None:
        from orbx.optimizer import PoseOptimizer
        for L in lanes[:2]:
            L.pz = PoseOptimizer(max_observations=B * cap, max_frames=B, device=local_rank)
            L.d_pose = torch.zeros((B, 7), dtype=torch.float64, device="cuda")
            L.d_inl = torch.zeros(B, dtype=torch.int32, device="cuda")
            L.d_outkp = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
            L.jobs = L.sq1.device_view().jobs

        def step_track_lane(i):
            step_device_lane(i, one_stream=True)
            L = lanes[i % 2]
            L.pz.from_matches_device(L.jobs, B, d_is2.data_ptr(), NLEVELS, synth.TUM1_K, L.d_pose.data_ptr(), L.d_inl.data_ptr(),
                                     L.d_outkp.data_ptr(), cap, L.stream.cuda_stream)

        for i in range(4):
            step_track_lane(i)
        torch.cuda.synchronize()
        KT2 = min(K, 100)
        t3_0 = torch.cuda.Event(enable_timing=True)
        t3_0.record(torch.cuda.current_stream())
        for L in lanes:
            L.stream.wait_event(t3_0)
        for i in range(KT2):
            step_track_lane(i)
        t3_1 = []
        for L in lanes[:2]:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(L.stream)
            t3_1.append(ev)
        torch.cuda.synchronize()
        ms_t2 = max(t3_0.elapsed_time(ev) for ev in t3_1) / KT2
        track["two_lanes"] = {"frames_per_s": B / (ms_t2 * 1e-3), "ms_per_step": ms_t2,
                              "note": "two handles in flight on two streams, like `value_two_lanes`"}
        for L in lanes[:2]:
            L.pz.close()
    JOB_BYTES = 208                                     # sizeof(orbx_frame_match_job): the per-step job array and last-frame poses the handle uploads
    h2d = B * W * H + B * JOB_BYTES + B * 48
    d2h = B * cap * 28 + B * cap * 32 + 4 * B + 4 * B * cap + 4 * B + 4 * B

    def collect(L):
        """wait for the lane's step in flight; its results are then in the lane's pinned host buffers"""
        if not L.busy:
            return 0
        L.sq.end()
        L.busy = False
        return int(L.out["nmatches"].sum())

    def step_host(i):
        L = lanes[i % NLANES]
        got = collect(L)
        slot = L.t % npool
        L.sq.begin(frames_host[slot], poses[slot], L.out)
        L.t += 1
        L.busy = True
        return got

    frames_host = host.numpy()                          # the pinned pool, as numpy views
    for L in lanes:
        L.sq.reset()
        L.t = 0
    for i in range(max(Wm, 2) * NLANES):
        step_host(i)
    for L in lanes:
        collect(L)
    barrier()
    t0 = time.perf_counter()
    nm_e2e = 0
    for i in range(K):
        nm_e2e += step_host(i)
    for L in lanes:
        nm_e2e += collect(L)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # ---- what the host side of this box can feed: the same bytes per step as plain cudaMemcpyAsync on the same number of streams,
    # no kernels, all ranks at once (the ceiling any end-to-end figure on this box lives under) ----
    ceil_in = [torch.empty((B, H, W), dtype=torch.uint8, device="cuda") for _ in range(NLANES)]
    ceil_out_d = [torch.empty(d2h, dtype=torch.uint8, device="cuda") for _ in range(NLANES)]
    ceil_out_h = [torch.empty(d2h, dtype=torch.uint8).pin_memory() for _ in range(NLANES)]

    def copy_step(i):
        L = lanes[i % NLANES]
        with torch.cuda.stream(L.stream):
            ceil_in[i % NLANES].copy_(host[i % npool], non_blocking=True)
            ceil_out_h[i % NLANES].copy_(ceil_out_d[i % NLANES], non_blocking=True)

    for i in range(2 * NLANES):
        copy_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        if i >= NLANES:
            lanes[i % NLANES].stream.synchronize()     # like collect(): a lane is reused only after its previous step has landed
        copy_step(i)
    torch.cuda.synchronize()
    ceil_s = time.perf_counter() - t0
    del ceil_in, ceil_out_d, ceil_out_h
    clk = clocks.stop()
    barrier()
    kp_per_frame = float(np.mean([L.out["counts"].mean() for L in lanes]))
    matches_per_frame = float(np.mean([L.out["nmatches"].mean() for L in lanes]))
    hnp = make_frames(64, seed0=100 * rank)          # the CPU legs' frames: 16 scenes x 4 shifts, same generator

    # ---- LocalBA (20 keyframes x 3000 points x ~12k edges, 5 + 10 iterations), rank 0 only ----
    lba = None
    side_sections = rank == 0 and default_workload and not args.extract_only
    if side_sections:
        from orbx.optimizer import Optimizer
        prob = synth.lba_problem(0, n_kf=20, n_pts=3000, stereo=False, n_fixed=1)
        op = Optimizer(max_keyframes=32, max_points=4096, max_edges=20000, device=local_rank)
        from orbx.optimizer import pack_problem
        prepared = op.prepare(pack_problem(prob))         # the argument and result structs of orbx_lba_solve_host, built once
        op.call(prepared)                                 # warm-up
        reps, trials = 10, 0
        t0 = time.perf_counter()
        for _ in range(reps):
            trials += op.call(prepared)                   # the C call alone
        lba_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(reps):
            op.LocalBundleAdjustment(prob)
        lba_mirror_s = time.perf_counter() - t0
        launches_lba = op.last_launches()
        lba_phase = op.phase_us()
        ms_build, _, _ = op.build_schur_timed(prob, 100.0, reps=50)
        lba = {"config": "C3: 20 keyframes (1 fixed), 3000 points, %d mono edges, optimize(5) + optimize(10)" % len(prob["e_kf"]),
               "lm_trials_per_s": trials / lba_s, "ms_per_window": 1e3 * lba_s / reps, "lm_trials_per_window": trials / reps,
               "schur_build_us": 1e3 * ms_build / 50, "kernel_launches_per_window": launches_lba,
               "api": "orbx_lba_solve_host (host buffers in and out, synchronous; the C call alone, argument structs built once)",
               "ms_per_window_python_mirror": 1e3 * lba_mirror_s / reps,
               "device_us_per_window_by_phase": {k: round(v, 1) for k, v in lba_phase.items()},
               "phase_note": "the whole-GPU kernel's own timers over the window's 15 trials: first quadratic form, Schur accumulation, collecting the "
                             "reduced system, reduced solve (one CTA), update, residuals + quadratic form at the trial state; the rest of "
                             "ms_per_window is the host's list building, one upload, two cooperative launches and the download",
               "schur_build": "residuals + Jacobians + quadratic form + Schur complement of one Levenberg trial, device time (CUDA events)"}
        # one Levenberg trial's system build moves ~2.59 MB (inputs + Hschur + bschur + D^-1 + Hpl + b_l) and does ~14 MFLOP (f64)
        pk, pk_src = peaks()
        lba["roofline"] = {"bound": "latency", "kernel": "Schur build of one Levenberg trial (k2_build + k2_pairs + k2_final)",
                           "algorithmic_bytes": 2.59e6, "achieved": 2.59e6 / (ms_build / 50 * 1e-3) / 1e9, "unit": "GB/s", "peak": pk,
                           "frac": 2.59e6 / (ms_build / 50 * 1e-3) / 1e9 / pk, "peak_source": pk_src,
                           "f64_tflops": 14e6 / (ms_build / 50 * 1e-3) / 1e12,
                           "note": "neither HBM- nor FMA-bound: three ~10 us launches at 9-12 % occupancy, a chain of dependent L2 round trips; "
                                   "the tensor-core (DMMA) form of the pair accumulation was built and measured 2.8x slower "
                                   "(profiles/r2_n_lba_dmma.txt)"}
        lba["batched"] = None      # filled below from every rank's share
        if not args.no_cpu:
            from oracle import oracle_py as O
            t0 = time.perf_counter()
            trials_c = 0
            tb_, ts_, nb_ = 0.0, 0.0, 0
            for _ in range(3):
                rc_ = O.lba_solve(prob)
                trials_c += rc_["trials"]
                tb_, ts_, nb_ = tb_ + rc_["t_build"], ts_ + rc_["t_schur"], nb_ + rc_["n_builds"]
            cpu_s = time.perf_counter() - t0
            # one system build = residuals + quadratic form (once per iteration) + Schur complement (once per trial)
            lba["cpu_schur_build_us"] = 1e6 * (tb_ / max(nb_, 1) + ts_ / max(trials_c, 1))
            lba["schur_build_speedup_vs_cpu_thread"] = lba["cpu_schur_build_us"] / lba["schur_build_us"]
            lba["cpu_lm_trials_per_s"] = trials_c / cpu_s
            lba["cpu_ms_per_window"] = 1e3 * cpu_s / 3
            lba["cpu"] = "C oracle (g2o restated), 1 thread, as g2o runs in the reference (OpenMP off)"
        op.close()

    # LocalBA windows sharded over the ranks: every rank solves its own 16 windows in flight
    lba_local = (0, 0, 0.0)
    if default_workload and not args.extract_only:
        barrier()
        lba_local = bench_lba_batched(local_rank, rank)
    c5_local = (0, 0.0, 0)
    if default_workload and not args.extract_only:
        barrier()
        c5_local = bench_c5(local_rank, rank, world)
    stereo = pose = bow = sequence = c4 = None
    if side_sections:
        with_cpu = not args.no_cpu and world == 1      # CPU legs are timed on rank 0 at N = 1 only
        stereo = bench_stereo(local_rank, with_cpu)
        pose = bench_pose(local_rank, with_cpu)
        bow = bench_bow(local_rank, with_cpu)
        sequence = bench_sequence(local_rank, with_cpu)
        c4 = bench_c4(local_rank, with_cpu)

    # the only collectives of the run: max of the timers, all-gather of per-rank counters
    from orbx import shard
    ms_total, e2e_s, ms_two_lanes, ceil_s = shard.max_over_ranks([ms_total, e2e_s, ms_two_lanes, ceil_s], device="cuda")
    counters = shard.gather_counters([B * K, int(round(kp_per_frame * B)), int(round(matches_per_frame * B)), lba_local[0], lba_local[1],
                                      int(lba_local[2] * 1e6), c5_local[0], int(c5_local[1] * 1e6), c5_local[2]], device="cuda")
    frames_total = sum(c[0] for c in counters)
    value = frames_total / (ms_total * 1e-3)
    e2e = frames_total / e2e_s

    if rank == 0 and lba is not None:
        lw, lt, ls = sum(c[3] for c in counters), sum(c[4] for c in counters), max(c[5] for c in counters) * 1e-6
        lba["batched"] = {"windows_in_flight_per_gpu": 16, "windows_per_s": lw / ls if ls > 0 else 0.0, "lm_trials_per_s": lt / ls if ls > 0 else 0.0,
                          "n_gpus": world, "windows_per_rank": [c[3] for c in counters],
                          "host_threads_per_rank": lba_host_threads(),
                          "api": "orbx_lba_solve_begin / orbx_lba_solve_end, one handle per window, submitted and collected by host_threads_per_rank threads; "
                                 "every rank solves its own windows, time = max over ranks"}
    c5 = None
    if rank == 0 and sum(c[6] for c in counters) > 0:
        pairs, secs = sum(c[6] for c in counters), max(c[7] for c in counters) * 1e-6
        c5 = {"config": "BASELINE.json config 5: 64 independent KITTI-stereo-shaped synthetic sequences (1241x376 rectified pairs, 2000 features, "
                        "Examples/Stereo/KITTI00-02.yaml intrinsics), sequence s -> rank s mod N; per pair: extract L + R, ComputeStereoMatches, "
                        "SearchByProjection(Cur, Last); host images + poses in, host results out",
              "pairs_per_s": pairs / secs, "n_gpus": world, "scaling": "strong (the 64 sequences are fixed)", "pairs_per_rank": [c[6] for c in counters],
              "seconds": secs, "matches_per_pair": sum(c[8] for c in counters) / max(pairs, 1),
              "h2d_bytes_per_pair": 2 * 1241 * 376,
              "api": "orbx_sequences_step_begin / orbx_sequences_step_end, stereo = 1, two handles in flight per rank; time = max over ranks"}
    if rank == 0:
        peak, peak_src = peaks()
        dom = max(stage_ms, key=stage_ms.get)
        dom_ms = stage_ms[dom] / max(runs, 1)
        n_launch_dom = 8 if dom == "pyramid" else 1
        achieved = ALGO_BYTES[dom] * B / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        tr = ncu_traffic() if B == 64 and default_workload else None
        traffic_dom = tr["kernels"].get(dom, {}).get("dram_bytes") if tr else None
        traffic_step = tr.get("step_dram_bytes") if tr else None
        traffic_src = tr.get("source") if tr else None
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": workload_config(B),
            "run": {"batch_per_gpu": B, "parallelism": "%d independent sequences per GPU in lockstep, sharded over %d GPU(s), no collective on the data path" % (B, world),
                    "host_enqueue_ms_per_step": 1e3 * t_enq,
                    "value_handle": "one orbx_sequences handle per GPU with n_sub = %d sub-batch pipelines (streams of their own, joined at the end of the timed region)" % args.subs,
                    "keypoints_per_frame": kp_per_frame, "matches_per_frame": matches_per_frame, "frames_per_rank": [c[0] for c in counters],
                    "per_step_working_set_mb": B * 3.3, "binary": binary_identity()},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "orbx_sequences_step_begin / orbx_sequences_step_end (include/orbx.h): pinned HOST frames + poses in, H2D, "
                           "extract + unprojection of the last frame + SearchByProjection(Cur, Last) on the device, D2H of keypoints, descriptors, "
                           "counts and match arrays into HOST buffers; %d handles (%d sequences each) in flight so copies of one overlap kernels of "
                           "the others; every step's results are read on the host" % (NLANES, B),
                    "matches_per_step": nm_e2e / K,
                    "gbs_per_gpu": (h2d + d2h) * K / e2e_s / 1e9,
                    "host_ceiling": {"frames_per_s": frames_total / ceil_s, "gbs_per_gpu": (B * W * H + d2h) * K / ceil_s / 1e9,
                                     "e2e_fraction_of_ceiling": ceil_s / e2e_s,
                                     "what": "the same bytes per step (frames in, results out) as plain cudaMemcpyAsync on the same %d streams, "
                                             "no kernels, all %d rank(s) copying at once, ranks bound to disjoint host cores" % (NLANES, world)}},
            "value_one_stream": {"value": B * world / (ms_one_stream * 1e-3), "unit": "frames/s", "ms_per_step": ms_one_stream,
                                 "note": "rank 0's handle with n_sub = 1 (the whole step on one stream), times the number of ranks"},
            "value_two_lanes": {"value": frames_total / (ms_two_lanes * 1e-3), "unit": "frames/s",
                                "note": "device-resident like `value`, but two handles in flight on two streams like `e2e`; `value` itself is "
                                        "one handle on one stream"},
            "gpu_launches": launches_per_step * K,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_dom, "traffic_source": traffic_src,
                         "peak_source": peak_src, "launches_per_step": n_launch_dom,
                         "algorithmic_bytes_per_frame": ALGO_BYTES[dom], "ms_per_step": dom_ms,
                         "whole_step": {"algorithmic_bytes_per_frame": FRAME_ALGO_BYTES,
                                        "achieved": FRAME_ALGO_BYTES * B / (ms_total / K * 1e-3) / 1e9,
                                        "frac": FRAME_ALGO_BYTES * B / (ms_total / K * 1e-3) / 1e9 / peak,
                                        "traffic": traffic_step}},
            "stage_ms_per_step": {k: v / max(runs, 1) for k, v in stage_ms.items()},
            "lba": lba,
            "c4": c4,
            "c5": c5,
            "stereo": stereo,
            "pose": pose,
            "bow": bow,
            "sequence": sequence,
            "track": track,
        }
        if not args.no_cpu and world == 1 and c5 is not None:
            cores_ = os.cpu_count() or 1
            cps, cn, cdt = cpu_c5(6.0, cores_)
            c5["cpu_pairs_per_s"] = cps
            c5["cpu"] = "the same chain on the C oracle, %d host threads (one sequence per thread), %d pairs in %.1f s" % (cores_, cn, cdt)
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            fps, n, dt = cpu_path(hnp[:64], 12.0, cores)
            ref_run = cpu_path_reference(10.0, cores) if default_workload else None
            cv_note = None
            try:                                          # context only: OpenCV's own SIMD ORB (not the reference's quadtree extractor)
                import cv2
                cv2.setNumThreads(1)
                orb = cv2.ORB_create(nfeatures=NFEAT, scaleFactor=SCALE, nlevels=NLEVELS, scoreType=cv2.ORB_FAST_SCORE, fastThreshold=INI_TH)
                bf_ = cv2.BFMatcher(cv2.NORM_HAMMING)
                prev_ = None
                t0 = time.perf_counter()
                ncv = 0
                while time.perf_counter() - t0 < 3.0:
                    k_, d_ = orb.detectAndCompute(hnp[ncv % 64], None)
                    if prev_ is not None and d_ is not None:
                        bf_.match(d_, prev_)
                    prev_ = d_
                    ncv += 1
                cv_note = {"frames_per_s_1_thread": ncv / (time.perf_counter() - t0),
                           "what": "cv2.ORB (FAST score, no quadtree distribution) detectAndCompute + brute-force Hamming match against the previous "
                                   "frame, one thread, cv2 %s; a different algorithm with SIMD kernels, shown for scale only" % cv2.__version__}
            except Exception as ex_:                      # noqa: BLE001
                cv_note = {"unavailable": str(ex_)[:100]}
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "opencv_orb": cv_note,
                                   "sample": "%d G-rect VGA frames (extract + SearchByProjection vs predecessor) in %.1f s on %d host threads (C oracle, one extractor per thread)" % (n, dt, cores)}
            if ref_run is not None:                     # the reference's own compiled sources (oracle/_ref), one process per core
                out["cpu_baseline"]["reference_sources"] = {
                    "value": ref_run[0], "unit": "frames/s", "cores": cores, "kind": "reference",
                    "sample": "%d frames in %.1f s on %d host processes: src/ORBextractor.cc + ORBmatcher.cc + Frame.cc + MapPoint.cc compiled "
                              "unmodified over the OpenCV stand-in (scalar primitives; shim builds Frame / MapPoint objects per frame)" % (ref_run[1], ref_run[2], cores)}
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    for L in lanes:
        L.sq.close()
    sq.close()


if __name__ == "__main__":
    main()

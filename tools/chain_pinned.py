#!/usr/bin/env python
"""one-call-per-frame stereo chain (orbx_sequences_step_host, 1 sequence, stereo + pose) with pageable and with pinned caller buffers"""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orb-slam2_b200"), os.path.join(ROOT, "tools")]
import numpy as np
import torch
import bench
from orbx import synth
from orbx.sequences import Sequences

world = synth.stereo_world(3, 640, 480)
imgs = [np.stack([world.render(0.02 * t, 0.0, 0.0), world.render(0.02 * t, 0.0, 0.0, right=True)]) for t in range(8)]
K = (world.fx, world.fy, world.cx, world.cy, world.bf)
T = np.zeros((8, 3, 4), np.float32)
for t in range(8):
    T[t, :, :3] = np.eye(3); T[t, 0, 3] = -0.02 * t
pin = lambda shape, dtype: torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True).numpy()
out = {}
for name, alloc, pin_in in (("pageable", None, False), ("pinned_out", pin, False), ("pinned_both", pin, True)):
    sq = Sequences(1, 640, 480, K, 1000, 1.2, 8, 20, 7, stereo=True, th=7.0, mono=False, device=0, pose=True)
    o = sq.alloc_outputs(alloc)
    fr = imgs
    if pin_in:
        fr = []
        for im in imgs:
            p = pin(im.shape, im.dtype); p[...] = im; fr.append(p)
    order = [0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4, 3, 2, 1]
    for i in range(28):
        sq.step(fr[order[i % 14]], T[order[i % 14]], o)
    t0 = time.perf_counter()
    for i in range(560):
        sq.step(fr[order[i % 14]], T[order[i % 14]], o)
    out[name] = 1e3 * (time.perf_counter() - t0) / 560
    sq.close()
print(json.dumps(out))

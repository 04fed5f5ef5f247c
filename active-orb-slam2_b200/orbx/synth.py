"""Seeded synthetic inputs for the parity tests and the bench.

Pure numpy (no cv2), so the same frames are produced in this container and on the GPU box.
"""
import numpy as np


def _smooth3(img):
    """cheap separable [1 2 1]/4 smoothing in integer arithmetic (keeps generators cv2-free)"""
    a = img.astype(np.int32)
    p = np.pad(a, 1, mode="reflect")
    h = p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]
    v = h[:-2] + 2 * h[1:-1] + h[2:]
    return ((v + 8) >> 4).astype(np.uint8)


def g_rect(seed, w=640, h=480, nrect=900, noise=3.0):
    """G-rect: background 110 + filled axis-aligned rectangles of random grey + N(0, noise^2)."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 110, np.float32)
    cx = rng.uniform(0, w, nrect)
    cy = rng.uniform(0, h, nrect)
    side = rng.uniform(4, 40, nrect)
    asp = rng.uniform(0.4, 1.6, nrect)
    grey = rng.uniform(0, 255, nrect)
    for i in range(nrect):
        hw, hh = 0.5 * side[i] * asp[i], 0.5 * side[i]
        x0, x1 = int(max(cx[i] - hw, 0)), int(min(cx[i] + hw, w))
        y0, y1 = int(max(cy[i] - hh, 0)), int(min(cy[i] + hh, h))
        img[y0:y1, x0:x1] = grey[i]
    img += rng.normal(0, noise, (h, w)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def g_noise(seed, w=640, h=480):
    """G-noise: uniform noise, lightly smoothed (dense corners everywhere)."""
    rng = np.random.default_rng(seed)
    return _smooth3(rng.integers(0, 256, (h, w), dtype=np.uint8))


def g_sparse(seed, w=640, h=480):
    """G-sparse: few rectangles, low noise -> empty cells and the minThFAST retry path."""
    return g_rect(seed, w, h, nrect=40, noise=1.0)


def g_flat(w=640, h=480, value=128):
    """G-flat: constant image -> zero keypoints (reference ORBextractor.cc:1064-1065)."""
    return np.full((h, w), value, np.uint8)


GENERATORS = {"rect": g_rect, "noise": g_noise, "sparse": g_sparse}


def frame(kind, seed, w=640, h=480):
    if kind == "flat":
        return g_flat(w, h)
    return GENERATORS[kind](seed, w, h)


# ---- stereo / sequence world: textured fronto-parallel planes seen by a moving rig ----
class PlaneWorld:
    """Three textured fronto-parallel planes; renders rectified left/right views for a camera pose.

    The renderer is a nearest-texel lookup with per-pixel plane selection by image row band, which keeps
    it exact integer work (deterministic across machines) while giving real disparity structure.
    """

    def __init__(self, seed, w=640, h=480, fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, bf=40.0):
        self.w, self.h, self.fx, self.fy, self.cx, self.cy, self.bf = w, h, fx, fy, cx, cy, bf
        self.depths = (2.0, 4.0, 8.0)
        self.tex = [g_rect(seed * 16 + i, 2048, 2048, nrect=9000) for i in range(3)]
        self.texscale = 220.0  # texels per metre

    def plane_of_row(self, v):
        b = (v * 3) // self.h
        return np.clip(b, 0, 2)

    def render(self, tx, ty, yaw, right=False):
        """view from camera at (tx, ty, 0) with yaw (rad) about y; right camera is offset by baseline."""
        h, w = self.h, self.w
        vs, us = np.mgrid[0:h, 0:w]
        pl = self.plane_of_row(vs)
        z = np.choose(pl, self.depths)
        base = self.bf / self.fx
        xc = (us - self.cx) / self.fx * z + (base if right else 0.0)
        yc = (vs - self.cy) / self.fy * z
        cs, sn = np.cos(yaw), np.sin(yaw)
        xw = cs * xc + sn * z + tx
        yw = yc + ty
        out = np.zeros((h, w), np.uint8)
        for i in range(3):
            m = pl == i
            tu = np.mod(np.rint(xw[m] * self.texscale + 1024).astype(np.int64), 2048)
            tv = np.mod(np.rint(yw[m] * self.texscale + 1024).astype(np.int64), 2048)
            out[m] = self.tex[i][tv, tu]
        return out

"""Host-side mirror of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111) over the orbx C ABI.

Same constructor arguments, same getters, same call result (keypoints in level-major order as cv::KeyPoint
records + an N x 32 uint8 descriptor matrix), same silent return on an empty image.  All compute happens in
liborbx.so (sm_100a CUDA); this module only moves buffers.
"""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib


class ORBextractor:
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)  -- ORBextractor.cc:410.

    max_width / max_height / max_batch size the device buffers once (the reference reallocates per call)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width=640, max_height=480,
                 max_batch=1, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        self.nfeatures, self.scaleFactor, self.nlevels = nfeatures, scaleFactor, nlevels
        self.iniThFAST, self.minThFAST = iniThFAST, minThFAST
        self.max_width, self.max_height, self.max_batch = max_width, max_height, max_batch
        check(self._L.orbx_extractor_create(C.byref(self._h), nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                                            max_width, max_height, max_batch, device))
        self.capacity = self._L.orbx_extractor_capacity(self._h)
        n = nlevels
        self._scale, self._inv_scale, self._sigma2, self._inv_sigma2 = (np.zeros(n, np.float32) for _ in range(4))
        self._quota = np.zeros(n, np.int32)
        check(self._L.orbx_extractor_tables(self._h, self._scale.ctypes.data, self._inv_scale.ctypes.data,
                                            self._sigma2.ctypes.data, self._inv_sigma2.ctypes.data, self._quota.ctypes.data))
        self._last_batch = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_extractor_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # ---- ORBextractor.h:62-82 ----
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self.scaleFactor

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv_scale.copy()

    def GetScaleSigmaSquares(self):
        return self._sigma2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._inv_sigma2.copy()

    def features_per_level(self):
        return self._quota.copy()

    # ---- operator(), ORBextractor.cc:1043 ----
    def __call__(self, image, mask=None):
        """-> (keypoints[KP_DTYPE], descriptors[N,32] uint8).  `mask` is ignored, like the reference."""
        if image is None or image.size == 0:
            self._last_batch = 0
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        kps, desc = self.extract_batch([image])
        return kps[0], desc[0]

    def extract_batch(self, images):
        """operator() on a batch of same-sized 8-bit single-channel frames -> (list of kps, list of desc)."""
        imgs = [np.asarray(im) for im in images]
        b = len(imgs)
        if b == 0:
            return [], []
        h, w = imgs[0].shape
        for im in imgs:
            if im.dtype != np.uint8 or im.ndim != 2 or im.shape != (h, w) or im.strides[1] != 1:
                raise ValueError("frames must be 2-D uint8 arrays of identical shape with unit column stride")
        stride = imgs[0].strides[0]
        if any(im.strides[0] != stride for im in imgs):
            imgs = [np.ascontiguousarray(im) for im in imgs]
            stride = w
        ptrs = (C.c_void_p * b)(*[im.ctypes.data for im in imgs])
        kps = np.empty((b, self.capacity), KP_DTYPE)
        desc = np.empty((b, self.capacity, 32), np.uint8)
        counts = np.zeros(b, np.int32)
        check(self._L.orbx_extractor_run_host(self._h, ptrs, b, w, h, stride, kps.ctypes.data, desc.ctypes.data,
                                              counts.ctypes.data))
        self._last_batch = b
        return [kps[i, :counts[i]].copy() for i in range(b)], [desc[i, :counts[i]].copy() for i in range(b)]

    def run_device(self, d_images, frame_pitch, batch, width, height, stride, d_kps, d_desc, d_counts, stream=0):
        """device-resident variant: all arguments are raw device pointers (ints); only enqueues on `stream`."""
        check(self._L.orbx_extractor_run_device(self._h, d_images, frame_pitch, batch, width, height, stride, d_kps, d_desc,
                                                d_counts, stream))
        self._last_batch = batch

    # ---- mvImagePyramid, ORBextractor.h:85 ----
    def level_shape(self, level, batch_idx=0):
        p, w, h, pitch = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        check(self._L.orbx_extractor_pyramid(self._h, batch_idx, level, C.byref(p), C.byref(w), C.byref(h), C.byref(pitch)))
        return w.value, h.value, pitch.value, p.value

    def pyramid_level(self, level, batch_idx=0, with_border=False):
        w, h, _, _ = self.level_shape(level, batch_idx)
        pad = 19 if with_border else 0
        out = np.empty((h + 2 * pad, w + 2 * pad), np.uint8)
        check(self._L.orbx_extractor_pyramid_host(self._h, batch_idx, level, int(with_border), out.ctypes.data, out.strides[0]))
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    # ---- stage access (parity tests) ----
    def blurred_level(self, level, batch_idx=0):
        w, h, _, _ = self.level_shape(level, batch_idx)
        out = np.empty((h, w), np.uint8)
        check(self._L.orbx_extractor_blurred_host(self._h, batch_idx, level, out.ctypes.data, out.strides[0]))
        return out

    def _packed(self, fn, level, batch_idx):
        n = C.c_int()
        check(fn(self._h, batch_idx, level, None, 0, C.byref(n)))
        buf = np.zeros(max(n.value, 1), np.uint32)
        check(fn(self._h, batch_idx, level, buf.ctypes.data, n.value, C.byref(n)))
        buf = buf[:n.value]
        return (buf & 0xfff).astype(np.int32), ((buf >> 12) & 0xfff).astype(np.int32), (buf >> 24).astype(np.int32)

    def candidates(self, level, batch_idx=0):
        """FAST candidates before DistributeOctTree: (x, y, score), coordinates relative to the 16-px border; unordered."""
        return self._packed(self._L.orbx_extractor_candidates_host, level, batch_idx)

    def level_keypoints(self, level, batch_idx=0):
        """keypoints kept by DistributeOctTree for one level, in the reference's list order."""
        return self._packed(self._L.orbx_extractor_level_keypoints_host, level, batch_idx)

    def last_launches(self):
        return self._L.orbx_extractor_last_launches(self._h)

    STAGES = ("pyramid", "fast", "quadtree", "blur", "describe")

    def profile(self, slots):
        check(self._L.orbx_extractor_profile(self._h, slots))

    def stage_ms(self):
        """-> (runs, {stage: summed milliseconds over those runs}) since the last call"""
        n = C.c_int()
        ms = np.zeros(5, np.float32)
        check(self._L.orbx_extractor_stage_ms(self._h, C.byref(n), ms.ctypes.data))
        return n.value, dict(zip(self.STAGES, ms.tolist()))

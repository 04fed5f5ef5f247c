"""Host-side mirror of the batched many-sequence mode of the orbx C ABI (include/orbx.h: orbx_sequences_*): n
independent sequences, one new frame (or rectified pair) per sequence per step, through the chain Tracking::TrackWithMotionModel
runs per frame (reference src/Tracking.cc:857-880, :900-965): ORBextractor::operator(), Frame::ComputeStereoMatches,
ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) against the last frame's keypoints unprojected with their depth
(Frame::UnprojectStereo, src/Frame.cc:695-709).  Host images and poses in, host results out; the last frame stays on the device."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib


class SequencesConfig(C.Structure):
    """orbx_sequences_config (include/orbx.h)"""
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32), ("ini_th", C.c_int32), ("min_th", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("n_sequences", C.c_int32), ("stereo", C.c_int32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("th", C.c_float), ("check_ori", C.c_int32), ("mono", C.c_int32), ("const_depth", C.c_float), ("device", C.c_int32),
                ("pose", C.c_int32), ("n_sub", C.c_int32)]


class SequencesOutputs(C.Structure):
    """orbx_sequences_outputs (include/orbx.h)"""
    _fields_ = [("kps", C.c_void_p), ("desc", C.c_void_p), ("counts", C.c_void_p), ("match", C.c_void_p), ("nmatches", C.c_void_p),
                ("u_right", C.c_void_p), ("depth", C.c_void_p), ("pose", C.c_void_p), ("n_inliers", C.c_void_p), ("outlier", C.c_void_p)]


class SequencesDevice(C.Structure):
    """orbx_sequences_device (include/orbx.h): device buffers of the last step"""
    _fields_ = [("extractor", C.c_void_p), ("kps", C.c_void_p), ("desc", C.c_void_p), ("counts", C.c_void_p), ("match", C.c_void_p),
                ("nmatches", C.c_void_p), ("u_right", C.c_void_p), ("depth", C.c_void_p), ("jobs", C.c_void_p),
                ("pose", C.c_void_p), ("n_inliers", C.c_void_p), ("outlier", C.c_void_p), ("stream", C.c_void_p)]


class Sequences:
    """One handle = n_sequences sequences in lockstep.  step() / begin() + end(); output arrays are allocated once by the caller of
    outputs() (numpy, or pinned torch tensors viewed through .numpy()) and reused every step."""

    def __init__(self, n_sequences, width, height, K, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, stereo=False,
                 th=7.0, check_ori=True, mono=False, const_depth=0.0, device=0, n_sub=0, pose=False):
        self._L = lib()
        c = SequencesConfig()
        c.nfeatures, c.scale_factor, c.nlevels, c.ini_th, c.min_th = nfeatures, scale_factor, nlevels, ini_th, min_th
        c.width, c.height, c.n_sequences, c.stereo = width, height, n_sequences, int(stereo)
        c.fx, c.fy, c.cx, c.cy, c.bf = (float(v) for v in K[:5])
        c.th, c.check_ori, c.mono, c.const_depth, c.device, c.n_sub, c.pose = th, int(check_ori), int(mono), const_depth, device, n_sub, int(pose)
        self.config = c
        self._h = C.c_void_p()
        check(self._L.orbx_sequences_create(C.byref(self._h), C.byref(c)))
        self.capacity = self._L.orbx_sequences_capacity(self._h)
        self.n_sequences, self.n_images = n_sequences, n_sequences * (2 if stereo else 1)
        self._out = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_sequences_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def alloc_outputs(self, alloc=None, descriptors=True):
        """dict of host arrays for one step's results; alloc(shape, dtype) -> array (default numpy.zeros)"""
        alloc = alloc or (lambda shape, dtype: np.zeros(shape, dtype))
        cap, ns, ni = self.capacity, self.n_sequences, self.n_images
        o = dict(kps=alloc((ni, cap, 28), np.uint8), counts=alloc((ni,), np.int32), match=alloc((ns, cap), np.int32), nmatches=alloc((ns,), np.int32))
        if descriptors:
            o["desc"] = alloc((ni, cap, 32), np.uint8)
        if self.config.stereo:
            o["u_right"], o["depth"] = alloc((ns, cap), np.float32), alloc((ns, cap), np.float32)
        if self.config.pose:
            o["pose"], o["n_inliers"], o["outlier"] = alloc((ns, 7), np.float64), alloc((ns,), np.int32), alloc((ns, cap), np.uint8)
        return o

    @staticmethod
    def _pack(o):
        S = SequencesOutputs()
        for k in ("kps", "desc", "counts", "match", "nmatches", "u_right", "depth", "pose", "n_inliers", "outlier"):
            if o.get(k) is not None:
                setattr(S, k, o[k].ctypes.data)
        return S

    def begin(self, images, Tcw, out):
        """images: uint8 [n_images, height, width(+padding)] C-contiguous per image; Tcw: float32 [n_sequences, 3, 4]; out: alloc_outputs()"""
        assert images.dtype == np.uint8 and images.shape[0] == self.n_images and images.strides[2] == 1
        Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(self.n_sequences, 12)
        self._keep = (images, Tcw, out, self._pack(out))
        check(self._L.orbx_sequences_step_begin(self._h, images.ctypes.data, images.strides[0], images.strides[1], Tcw.ctypes.data,
                                                C.byref(self._keep[3])))

    def end(self):
        check(self._L.orbx_sequences_step_end(self._h))
        out = self._keep[2]
        self._keep = None
        return out

    def step(self, images, Tcw, out=None):
        out = out if out is not None else self.alloc_outputs()
        self.begin(images, Tcw, out)
        return self.end()

    def step_device(self, d_images, frame_pitch, stride, Tcw, stream=None):
        """the step with the new images already on the device and the results left there; only enqueues (raw device pointer in)"""
        Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(self.n_sequences, 12)
        check(self._L.orbx_sequences_step_device(self._h, d_images, frame_pitch, stride, Tcw.ctypes.data, stream))

    def set_last_poses(self, Tcw_last):
        """the (optimised) poses of the frames of the last step, [n_sequences, 3, 4]: the next step unprojects the last frame with them"""
        T = np.ascontiguousarray(Tcw_last, np.float32).reshape(self.n_sequences, 12)
        check(self._L.orbx_sequences_set_last_poses(self._h, T.ctypes.data))

    def join(self, stream=None):
        """n_sub > 1: `stream` waits for every sub-batch's last step"""
        check(self._L.orbx_sequences_join(self._h, stream))

    def device_view(self):
        v = SequencesDevice()
        check(self._L.orbx_sequences_device_view(self._h, C.byref(v)))
        return v

    def profile(self, slots):
        """per-stage device timing of the handle's extractor (orbx_extractor_profile)"""
        check(self._L.orbx_extractor_profile(self.device_view().extractor, slots))

    def stage_ms(self):
        runs, ms = C.c_int(), np.zeros(5, np.float32)
        check(self._L.orbx_extractor_stage_ms(self.device_view().extractor, C.byref(runs), ms.ctypes.data))
        return runs.value, dict(zip(("pyramid", "fast", "quadtree", "blur", "describe"), ms.tolist()))

    def extractor_launches(self):
        return self._L.orbx_extractor_last_launches(self.device_view().extractor)

    def reset(self):
        check(self._L.orbx_sequences_reset(self._h))

    def last_launches(self):
        return self._L.orbx_sequences_last_launches(self._h)


def keypoints_of(out, image):
    """(keypoints, descriptors) of one image of a step's outputs, trimmed to its count"""
    n = int(out["counts"][image])
    return out["kps"][image].view(KP_DTYPE).reshape(-1)[:n], (out["desc"][image][:n] if "desc" in out else None)

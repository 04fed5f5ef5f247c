"""Host-side mirror of Frame::ComputeStereoMatches (reference src/Frame.cc:495-669) over the orbx C ABI.

The reference method reads mvKeys / mvKeysRight, both descriptor matrices and both extractors' mvImagePyramid and fills
mvuRight / mvDepth.  Here the pyramids stay where the two ORBextractor handles left them on the device; all compute
happens in liborbx.so (sm_100a CUDA)."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib


class StereoSide(C.Structure):
    """orbx_stereo_side (include/orbx.h)"""
    _fields_ = [("keys", C.c_void_p), ("desc", C.c_void_p), ("counts", C.c_void_p), ("pitch", C.c_int32), ("count_step", C.c_int32),
                ("extractor", C.c_void_p), ("first_slot", C.c_int32), ("slot_step", C.c_int32), ("max_count", C.c_int32)]


class StereoMatcher:
    def __init__(self, max_keypoints=4096, max_pairs=1, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.orbx_stereo_create(C.byref(self._h), max_keypoints, max_pairs, device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_stereo_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def ComputeStereoMatches(self, left, right, keys_l, desc_l, keys_r, desc_r, bf, b, left_slot=0, right_slot=0):
        """left / right: ORBextractor mirrors that have just run on the two images (their pyramids are read on the device).
        -> (mvuRight, mvDepth, number kept after the median cut)"""
        kl, kr = np.ascontiguousarray(keys_l, KP_DTYPE), np.ascontiguousarray(keys_r, KP_DTYPE)
        dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
        n = len(kl)
        ur, dp = np.full(max(n, 1), -1, np.float32), np.full(max(n, 1), -1, np.float32)
        kept = C.c_int32()
        check(self._L.orbx_stereo_matches_host(self._h, left._h, left_slot, right._h, right_slot, kl.ctypes.data, dl.ctypes.data, n,
                                               kr.ctypes.data, dr.ctypes.data, len(kr), bf, b, ur.ctypes.data, dp.ctypes.data,
                                               C.byref(kept)))
        return ur[:n], dp[:n], kept.value

    def ComputeStereoMatchesFromExtractors(self, left, right, n_left, bf, b, left_slot=0, right_slot=0):
        """the same, for the pair the two extractor mirrors just processed through their host entry point: keypoints and descriptors
        are read from the extractors' device buffers instead of being uploaded again"""
        ur, dp = np.full(max(n_left, 1), -1, np.float32), np.full(max(n_left, 1), -1, np.float32)
        kept = C.c_int32()
        check(self._L.orbx_stereo_matches_extractors_host(self._h, left._h, left_slot, right._h, right_slot, n_left, bf, b, ur.ctypes.data,
                                                          dp.ctypes.data, C.byref(kept)))
        return ur[:n_left], dp[:n_left], kept.value

    def matches_device(self, left: StereoSide, right: StereoSide, n_pairs, bf, b, d_u_right, d_depth, out_pitch, d_kept=None, stream=0):
        check(self._L.orbx_stereo_matches_device(self._h, C.byref(left), C.byref(right), n_pairs, bf, b, d_u_right, d_depth, out_pitch,
                                                 d_kept, stream))

    def last_launches(self):
        return self._L.orbx_stereo_last_launches(self._h)

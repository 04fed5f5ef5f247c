"""ctypes loader of liborbx.so (the C ABI of include/orbx.h).  There is no CPU fallback: if the CUDA library is
missing or no sm_100 device is present, every entry point raises."""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "liborbx.so")
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])   # cv::KeyPoint, 28 bytes
assert KP_DTYPE.itemsize == 28

ORBX_OK = 0
STATUS_NAMES = {0: "OK", -1: "INVALID", -2: "CUDA", -3: "NO_DEVICE", -4: "CAPACITY", -5: "UNSUPPORTED", -6: "NOMEM",
                -7: "ABORTED"}


class OrbxError(RuntimeError):
    def __init__(self, status, text):
        super().__init__("orbx: %s (%d): %s" % (STATUS_NAMES.get(status, "?"), status, text))
        self.status = status


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the orbx hot path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        _declare(L)
        _LIB = L
    return _LIB


def check(status):
    if status != ORBX_OK:
        raise OrbxError(status, lib().orbx_last_error().decode("utf-8", "replace"))


def _declare(L):
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_last_error.argtypes = []
    L.orbx_version.argtypes = []
    L.orbx_extractor_create.argtypes = [C.POINTER(vp), i, f, i, i, i, i, i, i, i]
    L.orbx_extractor_destroy.restype = None
    L.orbx_extractor_destroy.argtypes = [vp]
    L.orbx_extractor_capacity.argtypes = [vp]
    L.orbx_extractor_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.orbx_extractor_run_host.argtypes = [vp, vp, i, i, i, i, vp, vp, vp]
    L.orbx_extractor_run_device.argtypes = [vp, vp, sz, i, i, i, i, vp, vp, vp, vp]
    L.orbx_extractor_pyramid.argtypes = [vp, i, i, C.POINTER(vp), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.orbx_extractor_pyramid_host.argtypes = [vp, i, i, i, vp, i]
    L.orbx_extractor_candidates_host.argtypes = [vp, i, i, vp, i, C.POINTER(i)]
    L.orbx_extractor_level_keypoints_host.argtypes = [vp, i, i, vp, i, C.POINTER(i)]
    L.orbx_extractor_blurred_host.argtypes = [vp, i, i, vp, i]
    L.orbx_extractor_last_launches.argtypes = [vp]
    L.orbx_hamming256.argtypes = [vp, vp]
    L.orbx_matcher_create.argtypes = [C.POINTER(vp), i, i, i, i]
    L.orbx_matcher_destroy.restype = None
    L.orbx_matcher_destroy.argtypes = [vp]
    L.orbx_match_projection_points_host.argtypes = [vp, vp, i, vp, vp, f, f, vp, vp]
    L.orbx_match_projection_frame_host.argtypes = [vp, vp, i, vp, vp, vp, vp, i, i, f, i, vp, vp]
    L.orbx_match_projection_frame_device.argtypes = [vp, vp, i, vp]
    L.orbx_match_projection_keyframe_host.argtypes = [vp, vp, i, vp, vp, vp, vp, f, i, i, vp, vp]
    L.orbx_match_buckets_host.argtypes = [vp, vp, vp, vp]
    L.orbx_match_initialization_host.argtypes = [vp, vp, vp, vp, i, f, i, vp, vp]
    L.orbx_match_window_host.argtypes = [vp, vp, i, vp, vp, i, vp, i, vp, vp, vp]
    L.orbx_matcher_last_launches.argtypes = [vp]
    L.orbx_matcher_last_sweeps.argtypes = [vp, vp, i]
    L.orbx_lba_create.argtypes = [C.POINTER(vp), i, i, i, i]
    L.orbx_lba_destroy.restype = None
    L.orbx_lba_destroy.argtypes = [vp]
    L.orbx_lba_solve_host.argtypes = [vp, vp, i, i, vp]
    L.orbx_lba_solve_begin.argtypes = [vp, vp, i, i]
    L.orbx_lba_solve_end.argtypes = [vp, vp, vp]
    L.orbx_lba_build_schur_timed.argtypes = [vp, vp, C.c_double, i, vp, vp, vp]
    L.orbx_lba_last_launches.argtypes = [vp]
    L.orbx_lba_phase_ns.argtypes = [vp, vp]
    L.orbx_stereo_create.argtypes = [C.POINTER(vp), i, i, i]
    L.orbx_stereo_destroy.restype = None
    L.orbx_stereo_destroy.argtypes = [vp]
    L.orbx_stereo_matches_device.argtypes = [vp, vp, vp, i, f, f, vp, vp, i, vp, vp]
    L.orbx_stereo_matches_host.argtypes = [vp, vp, i, vp, i, vp, vp, i, vp, vp, i, f, f, vp, vp, vp]
    L.orbx_stereo_matches_extractors_host.argtypes = [vp, vp, i, vp, i, i, f, f, vp, vp, vp]
    L.orbx_stereo_last_launches.argtypes = [vp]
    L.orbx_pose_create.argtypes = [C.POINTER(vp), i, i, i]
    L.orbx_pose_destroy.restype = None
    L.orbx_pose_destroy.argtypes = [vp]
    L.orbx_pose_optimize_host.argtypes = [vp, vp, i, vp]
    L.orbx_pose_from_matches_device.argtypes = [vp, vp, i, vp, i, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp, vp, i, vp]
    L.orbx_pose_last_launches.argtypes = [vp]
    L.orbx_vocabulary_create.argtypes = [C.POINTER(vp), i, vp, vp, vp, vp, vp, i, i, i]
    L.orbx_vocabulary_destroy.restype = None
    L.orbx_vocabulary_destroy.argtypes = [vp]
    L.orbx_vocabulary_transform_host.argtypes = [vp, vp, i, i, vp, vp, vp]
    L.orbx_vocabulary_transform_device.argtypes = [vp, i, vp, vp, i, i, i, i, vp, vp, vp, vp]
    L.orbx_vocabulary_last_launches.argtypes = [vp]
    L.orbx_mappoints_create.argtypes = [C.POINTER(vp), i, i, i]
    L.orbx_mappoints_destroy.restype = None
    L.orbx_mappoints_destroy.argtypes = [vp]
    L.orbx_mappoints_distinctive_host.argtypes = [vp, i, vp, vp, vp, vp]
    L.orbx_mappoints_last_launches.argtypes = [vp]
    L.orbx_frustum_host.argtypes = [vp, i, vp, vp, vp, i]
    L.orbx_frustum_device.argtypes = [vp, i, vp, vp, vp, vp]
    L.orbx_sequences_create.argtypes = [C.POINTER(vp), vp]
    L.orbx_sequences_destroy.restype = None
    L.orbx_sequences_destroy.argtypes = [vp]
    L.orbx_sequences_capacity.argtypes = [vp]
    L.orbx_sequences_reset.argtypes = [vp]
    L.orbx_sequences_step_begin.argtypes = [vp, vp, sz, i, vp, vp]
    L.orbx_sequences_step_end.argtypes = [vp]
    L.orbx_sequences_step_host.argtypes = [vp, vp, sz, i, vp, vp]
    L.orbx_sequences_last_launches.argtypes = [vp]
    L.orbx_sequences_step_device.argtypes = [vp, vp, sz, i, vp, vp]
    L.orbx_sequences_device_view.argtypes = [vp, vp]
    L.orbx_sequences_join.argtypes = [vp, vp]
    L.orbx_sequences_set_last_poses.argtypes = [vp, vp]
    L.orbx_extractor_profile.argtypes = [vp, i]
    L.orbx_extractor_stage_ms.argtypes = [vp, C.POINTER(i), vp]

"""Host-side mirror of MapPoint::ComputeDistinctiveDescriptors (reference src/MapPoint.cc:275-340) over the orbx C ABI,
batched over map points.  All compute happens in liborbx.so (sm_100a CUDA)."""
import ctypes as C

import numpy as np

from ._lib import check, lib


class MapPointOps:
    def __init__(self, max_points=8192, max_descriptors=262144, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.orbx_mappoints_create(C.byref(self._h), max_points, max_descriptors, device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_mappoints_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def ComputeDistinctiveDescriptors(self, descriptor_sets):
        """descriptor_sets: list of [N_p, 32] uint8 arrays (vDescriptors of every map point) -> (best index per point, its median)"""
        n = len(descriptor_sets)
        start = np.zeros(n + 1, np.int32)
        for p, d in enumerate(descriptor_sets):
            start[p + 1] = start[p] + len(d)
        desc = np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in descriptor_sets]) if n and start[n] else np.zeros((0, 32), np.uint8)
        desc = np.ascontiguousarray(desc)
        best, med = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        check(self._L.orbx_mappoints_distinctive_host(self._h, n, start.ctypes.data, desc.ctypes.data, best.ctypes.data, med.ctypes.data))
        return best[:n], med[:n]

    def last_launches(self):
        return self._L.orbx_mappoints_last_launches(self._h)

"""Host-side mirror of `class CudaCoarseTracker`
(tandem/libdr/cuda_coarse_tracker/include/public/cuda_coarse_tracker.h:9-82) over the C ABI.
Eigen types of the reference signature become numpy arrays (Vec6 -> (6,), Mat88 -> (8,8), Vec8 -> (8,))."""
import ctypes

import numpy as np

from ._lib import check, lib

_fp = ctypes.POINTER(ctypes.c_float)
_dp = ctypes.POINTER(ctypes.c_double)


def _f(a):
    return np.ascontiguousarray(a, np.float32)


class CudaCoarseTracker:
    def __init__(self, w, h, setting_huberTH=9.0, setting_coarseCutoffTH=20.0, device=0):
        self.w, self.h = w, h
        self._args = (w, h, float(setting_huberTH), float(setting_coarseCutoffTH), device)
        self._h = None

    def init(self, n_max=0):
        if self._h is not None:
            raise RuntimeError("Cannot call CudaCoarseTracker::init more than once.")
        w, h, hub, cut, dev = self._args
        self._h = ctypes.c_void_p()
        check(lib().tdm_tracker_create(w, h, hub, cut, n_max, dev, ctypes.byref(self._h)))

    def free(self):
        if self._h is not None:
            lib().tdm_tracker_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def setK(self, w, h, fx, fy, cx, cy):
        check(lib().tdm_tracker_set_k(self._h, w, h, fx, fy, cx, cy))

    def setReference(self, n, pc_u, pc_v, pc_idepth, pc_color, ref_exposure, ref_aff_g2l):
        u, v, i, c = _f(pc_u), _f(pc_v), _f(pc_idepth), _f(pc_color)
        aff = np.ascontiguousarray(ref_aff_g2l, np.float64)
        check(lib().tdm_tracker_set_reference(self._h, n, u.ctypes.data_as(_fp), v.ctypes.data_as(_fp),
                                              i.ctypes.data_as(_fp), c.ctypes.data_as(_fp), float(ref_exposure),
                                              aff.ctypes.data_as(_dp)))

    def setNew(self, dInew):
        d = _f(dInew)
        assert d.size == self.w * self.h * 3
        check(lib().tdm_tracker_set_new(self._h, d.ctypes.data_as(_fp)))

    def calcRes(self, refToNew, new_exposure, aff_g2l, cutoffTH):
        T = np.ascontiguousarray(refToNew, np.float64)
        aff = np.ascontiguousarray(aff_g2l, np.float64)
        res = np.zeros(6, np.float64)
        check(lib().tdm_tracker_calc_res(self._h, T.ctypes.data_as(_dp), float(new_exposure), aff.ctypes.data_as(_dp),
                                         float(cutoffTH), res.ctypes.data_as(_dp)))
        return res

    def calcG(self, new_exposure, aff_g2l):
        aff = np.ascontiguousarray(aff_g2l, np.float64)
        H = np.zeros((8, 8), np.float64)
        b = np.zeros(8, np.float64)
        check(lib().tdm_tracker_calc_g(self._h, float(new_exposure), aff.ctypes.data_as(_dp), H.ctypes.data_as(_dp),
                                       b.ctypes.data_as(_dp)))
        return H, b

    def calcResAndG(self, refToNew, new_exposure, aff_g2l, cutoffTH):
        T = np.ascontiguousarray(refToNew, np.float64)
        aff = np.ascontiguousarray(aff_g2l, np.float64)
        res = np.zeros(6, np.float64)
        H = np.zeros((8, 8), np.float64)
        b = np.zeros(8, np.float64)
        check(lib().tdm_tracker_calc_res_g(self._h, T.ctypes.data_as(_dp), float(new_exposure), aff.ctypes.data_as(_dp),
                                           float(cutoffTH), res.ctypes.data_as(_dp), H.ctypes.data_as(_dp),
                                           b.ctypes.data_as(_dp)))
        return res, H, b

    def synchronize(self):
        check(lib().tdm_tracker_synchronize(self._h))

    def run_resident(self, iters):
        ms = ctypes.c_float()
        check(lib().tdm_tracker_run_resident(self._h, iters, ctypes.byref(ms)))
        return ms.value

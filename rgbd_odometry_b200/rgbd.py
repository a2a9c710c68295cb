"""RGBDAligner: Python host-side mirror of the dvo_rgbd_* C-ABI (RGBDOdometry, the semi-dense photometric Gauss-Newton)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import RgbdConfig, RgbdInfo, RgbdParams, check

MEM_HOST, MEM_DEVICE = 0, 1
FRAME_REF, FRAME_NOW = 0, 1


def rgbd_params(iterations=3, gradient_threshold=5, min_points=100, max_points=50000, eps_norm_exit=200.0):
    """Defaults are the reference's constants (src/RGBDOdometry.cpp:32-34, :541, :556)."""
    return RgbdParams(iterations, gradient_threshold, min_points, max_points, eps_norm_exit)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RGBDAligner:
    def __init__(self, width=640, height=480, levels=4, max_batch=1, device=0, intrinsics=(525.0, 525.0, 319.5, 239.5)):
        self.lib = _lib.load()
        self.cfg = RgbdConfig(width, height, levels, max_batch, device)
        self.h = C.c_void_p()
        check(self.lib.dvo_rgbd_create(C.byref(self.cfg), C.byref(self.h)), "dvo_rgbd_create")
        self.levels, self.max_batch = levels, max_batch
        if intrinsics is not None:
            self.set_intrinsics(*intrinsics)

    def close(self):
        if self.h:
            self.lib.dvo_rgbd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        check(self.lib.dvo_rgbd_set_stream(self.h, C.c_void_p(stream_ptr) if stream_ptr else None), "dvo_rgbd_set_stream")

    def synchronize(self):
        check(self.lib.dvo_rgbd_synchronize(self.h), "dvo_rgbd_synchronize")

    def launch_count(self):
        return int(self.lib.dvo_rgbd_launch_count(self.h))

    def set_intrinsics(self, fx, fy, cx, cy):
        check(self.lib.dvo_rgbd_set_intrinsics(self.h, fx, fy, cx, cy), "dvo_rgbd_set_intrinsics")

    def set_frames(self, frame, bgr, depth, first=0, count=None, device=False):
        if not device:
            bgr = np.ascontiguousarray(bgr, np.uint8); depth = np.ascontiguousarray(depth, np.uint16)
            count = bgr.shape[0] if count is None else count
            check(self.lib.dvo_rgbd_set_frames(self.h, frame, first, count, _ptr(bgr), _ptr(depth), MEM_HOST), "dvo_rgbd_set_frames")
        else:
            check(self.lib.dvo_rgbd_set_frames(self.h, frame, first, count, C.c_void_p(bgr), C.c_void_p(depth), MEM_DEVICE), "dvo_rgbd_set_frames")

    def compute_jacobians(self, count, first=0, gradient_threshold=5):
        check(self.lib.dvo_rgbd_compute_jacobians(self.h, first, count, gradient_threshold), "dvo_rgbd_compute_jacobians")

    def set_pose(self, count, T=None, first=0):
        if T is not None:
            T = np.ascontiguousarray(T, np.float64).reshape(count, 16)
        check(self.lib.dvo_rgbd_set_pose(self.h, first, count, _ptr(T)), "dvo_rgbd_set_pose")

    def gauss_newton(self, count, level, params=None, first=0):
        params = params or rgbd_params()
        check(self.lib.dvo_rgbd_gauss_newton(self.h, first, count, level, C.byref(params)), "dvo_rgbd_gauss_newton")

    def get_poses(self, count, first=0):
        T = np.empty((count, 4, 4), np.float64)
        info = (RgbdInfo * count)()
        check(self.lib.dvo_rgbd_get_poses(self.h, first, count, _ptr(T), C.cast(info, C.c_void_p)), "dvo_rgbd_get_poses")
        return T, info

    def level_dims(self, level):
        r, c = C.c_int(), C.c_int()
        check(self.lib.dvo_rgbd_level_dims(self.h, level, C.byref(r), C.byref(c)), "dvo_rgbd_level_dims")
        return r.value, c.value

    def get_level(self, slot, frame, level):
        r, c = self.level_dims(level)
        gray = np.empty((r, c), np.uint8); depth = np.empty((r, c), np.uint16)
        check(self.lib.dvo_rgbd_get_level(self.h, slot, frame, level, _ptr(gray), _ptr(depth)), "dvo_rgbd_get_level")
        return gray, depth

    def get_A(self, slot, level):
        A = np.empty(36, np.float64); n = C.c_int()
        check(self.lib.dvo_rgbd_get_A(self.h, slot, level, _ptr(A), C.byref(n)), "dvo_rgbd_get_A")
        return A.reshape(6, 6), n.value

    def eval(self, slot, level, T, gradient_threshold=5):
        r, c = self.level_dims(level)
        cap = r * c
        T = np.ascontiguousarray(T, np.float64).reshape(16)
        ij = np.empty((cap, 2), np.int32); J = np.empty((cap, 6), np.float64); eps = np.empty(cap, np.float64); uv = np.empty((cap, 2), np.int32)
        b = np.empty(6, np.float64); ss = C.c_double(); nv = C.c_int(); n = C.c_int()
        check(self.lib.dvo_rgbd_eval(self.h, slot, level, _ptr(T), gradient_threshold, cap, C.byref(n), _ptr(ij), _ptr(J), _ptr(eps), _ptr(uv), _ptr(b),
                                     C.byref(ss), C.byref(nv)), "dvo_rgbd_eval")
        k = n.value
        return {"ij": ij[:k].copy(), "J": J[:k].copy(), "eps": eps[:k].copy(), "uv": uv[:k].copy(), "b": b, "sumsq": ss.value, "nvis": nv.value}

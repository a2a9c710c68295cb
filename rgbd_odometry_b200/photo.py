"""PhotoEstimator: Python host-side mirror of the EPoseEstimator / PyramidalStorageStruct C-ABI (dvo_photo_*)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import PhotoConfig, PhotoInfo, check

WHICH = {"bgr": (0, np.uint8), "gray": (1, np.uint8), "depth": (2, np.uint16), "X": (3, np.float64), "Y": (4, np.float64),
         "Z": (5, np.float64), "J": (6, np.float64), "grayVals": (7, np.float64), "redVals": (8, np.float64),
         "greenVals": (9, np.float64), "blueVals": (10, np.float64)}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PhotoEstimator:
    def __init__(self, width=640, height=480, levels=5, max_batch=1, device=0, intrinsics=(525.0, 525.0, 319.5, 239.5)):
        self.lib = _lib.load()
        self.cfg = PhotoConfig(width, height, levels, max_batch, device)
        self.h = C.c_void_p()
        check(self.lib.dvo_photo_create(C.byref(self.cfg), C.byref(self.h)), "dvo_photo_create")
        self.width, self.height, self.levels, self.max_batch = width, height, levels, max_batch
        if intrinsics is not None:
            self.set_intrinsics(*intrinsics)

    def close(self):
        if self.h:
            self.lib.dvo_photo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, ptr):
        check(self.lib.dvo_photo_set_stream(self.h, C.c_void_p(ptr) if ptr else None), "dvo_photo_set_stream")

    def synchronize(self):
        check(self.lib.dvo_photo_synchronize(self.h), "dvo_photo_synchronize")

    def launch_count(self):
        return int(self.lib.dvo_photo_launch_count(self.h))

    def set_intrinsics(self, fx, fy, cx, cy):
        check(self.lib.dvo_photo_set_intrinsics(self.h, fx, fy, cx, cy), "dvo_photo_set_intrinsics")

    def set_frames(self, frame, bgr, depth=None, first=0, count=None, device=False):
        """setRefFrame (frame 0) / setNowFrame (frame 1)."""
        if not device:
            bgr = np.ascontiguousarray(bgr, np.uint8)
            count = bgr.shape[0] if count is None else count
            if depth is not None:
                depth = np.ascontiguousarray(depth, np.uint16)
            check(self.lib.dvo_photo_set_frames(self.h, frame, first, count, _ptr(bgr), _ptr(depth), 0), "dvo_photo_set_frames")
        else:
            check(self.lib.dvo_photo_set_frames(self.h, frame, first, count, C.c_void_p(bgr), C.c_void_p(depth) if depth else None, 1),
                  "dvo_photo_set_frames")

    def prepare_ref(self, count, first=0, compat=False):
        check(self.lib.dvo_photo_prepare_ref(self.h, first, count, int(compat)), "dvo_photo_prepare_ref")

    def set_pose(self, count, poses=None, first=0):
        if poses is not None:
            poses = np.ascontiguousarray(poses, np.float64).reshape(count, 12)
        check(self.lib.dvo_photo_set_pose(self.h, first, count, _ptr(poses)), "dvo_photo_set_pose")

    def estimate(self, count, level, iters=3, first=0, compat=False, huber_k=0.0, lambda0=0.0):
        check(self.lib.dvo_photo_estimate(self.h, first, count, level, iters, int(compat), huber_k, lambda0), "dvo_photo_estimate")

    def get_poses(self, count, first=0):
        poses = np.empty((count, 12), np.float64)
        info = (PhotoInfo * count)()
        check(self.lib.dvo_photo_get_poses(self.h, first, count, _ptr(poses), C.cast(info, C.c_void_p)), "dvo_photo_get_poses")
        return poses, info

    def get_level(self, slot, frame, level, which, compat=False):
        """PyramidalStorageStruct::getLevel member `which` of one slot / level."""
        code, dt = WHICH[which]
        h, w = self.height >> level, self.width >> level
        shape = (h, w, 3) if which == "bgr" else ((h * w, 6) if which == "J" else (h, w))
        out = np.empty(shape, dt)
        check(self.lib.dvo_photo_get_level(self.h, slot, frame, level, code, _ptr(out), out.nbytes, int(compat)), "dvo_photo_get_level")
        return out

    def get_A(self, slot, level):
        A = np.empty((6, 6), np.float64)
        check(self.lib.dvo_photo_get_A(self.h, slot, level, _ptr(A)), "dvo_photo_get_A")
        return A

    def eval(self, slot, level, R, T, compat=False, huber_k=0.0, want_canvas=False):
        pose = np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(T, np.float64).reshape(3)])
        b, A = np.empty(6), np.empty((6, 6))
        sumsq, nre, nus = C.c_double(), C.c_int(), C.c_int()
        cv = np.empty((self.height >> level, self.width >> level), np.float64) if want_canvas else None
        check(self.lib.dvo_photo_eval(self.h, slot, level, _ptr(pose), int(compat), huber_k, _ptr(b), _ptr(A), C.byref(sumsq),
                                      C.byref(nre), C.byref(nus), _ptr(cv)), "dvo_photo_eval")
        return {"b": b, "A": A, "sumsq": sumsq.value, "nreproj": nre.value, "nused": nus.value, "canvas": cv}

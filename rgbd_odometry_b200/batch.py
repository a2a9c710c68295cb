"""BatchAligner: Python host-side mirror of the batch C-ABI (one object = one dvo_ctx on one GPU)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Config, KeyframePolicy, PairInfo, SolverParams, check

SUBGRAD_REF, GN, LM = 0, 1, 2
JAC_REFERENCE, JAC_EXACT = 0, 1
W_REF_CAUCHY, W_HUBER, W_NONE = 0, 1, 2
ARITH_EXACT, ARITH_FAST = 0, 1
RES_DT_FLOOR, RES_DT_INTERP = 0, 1
FRAME_REF, FRAME_NOW = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
BUF = {"gray": (0, np.uint8), "depth": (1, np.uint16), "edge": (2, np.uint8), "d2": (3, np.int32),
       "dtn": (4, np.float32), "gx": (5, np.float32), "gy": (6, np.float32)}


def solver_params(solver=SUBGRAD_REF, jacobian=JAC_REFERENCE, weight=W_REF_CAUCHY, arithmetic=ARITH_EXACT, huber_k=1.345,
                  lm_lambda0=1e-3, iters=(50, 50, 50, 50), residual=RES_DT_FLOOR):
    p = SolverParams()
    p.solver, p.jacobian, p.weight, p.arithmetic, p.residual = solver, jacobian, weight, arithmetic, residual
    p.huber_k, p.lm_lambda0 = huber_k, lm_lambda0
    for i in range(_lib.MAX_LEVELS):
        p.iters[i] = int(iters[i]) if i < len(iters) else 0
    return p


def keyframe_policy(keyframe_every=5, use_quality_gates=False, laplacian_thresh=3.0, visible_ratio_thresh=0.8, min_reprojections=50):
    """Defaults are the reference's constants (src/SolveDVO.cpp:22-23, :2145, :2156); the gates ship disabled there."""
    return KeyframePolicy(keyframe_every, int(use_quality_gates), laplacian_thresh, visible_ratio_thresh, min_reprojections)


def undistort(img, K, D):
    """cv::undistort of a batch of images (N, H, W[, 3]) u8 or (N, H, W) u16 on the GPU (dvo_undistort, host buffers)."""
    lib = _lib.load()
    img = np.ascontiguousarray(img)
    n, h, w = img.shape[:3]
    typ = 2 if img.dtype == np.uint16 else (1 if img.ndim == 4 and img.shape[3] == 3 else 0)
    out = np.empty_like(img)
    K4 = np.array(K, np.float64); D5 = np.array(D, np.float64)
    check(lib.dvo_undistort(img.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), w, h, typ, n, K4.ctypes.data_as(C.c_void_p),
                            D5.ctypes.data_as(C.c_void_p), MEM_HOST, None), "dvo_undistort")
    return out


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class BatchAligner:
    """Owns a dvo_ctx.  Arrays are numpy (host) unless `device=True`, in which case raw device pointers (ints) are passed."""

    def __init__(self, width=640, height=480, levels=4, max_batch=1, device=0, keep_now_depth=False, trace_iters=0,
                 intrinsics=(525.0, 525.0, 319.5, 239.5)):
        self.lib = _lib.load()
        self.cfg = Config(width, height, levels, max_batch, device, int(keep_now_depth), trace_iters)
        self.h = C.c_void_p()
        check(self.lib.dvo_create(C.byref(self.cfg), C.byref(self.h)), "dvo_create")
        self.levels, self.max_batch, self.width, self.height = levels, max_batch, width, height
        if intrinsics is not None:
            self.set_intrinsics(*intrinsics)

    def close(self):
        if self.h:
            self.lib.dvo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        check(self.lib.dvo_set_stream(self.h, C.c_void_p(stream_ptr) if stream_ptr else None), "dvo_set_stream")

    def synchronize(self):
        check(self.lib.dvo_synchronize(self.h), "dvo_synchronize")

    def set_intrinsics(self, fx, fy, cx, cy):
        check(self.lib.dvo_set_intrinsics(self.h, fx, fy, cx, cy), "dvo_set_intrinsics")

    def set_frames(self, frame, gray, depth=None, first=0, count=None, device=False):
        if not device:
            gray = np.ascontiguousarray(gray, np.uint8)
            count = gray.shape[0] if count is None else count
            if depth is not None:
                depth = np.ascontiguousarray(depth, np.uint16)
        check(self.lib.dvo_set_frames(self.h, frame, first, count, _ptr(gray), _ptr(depth), MEM_DEVICE if device else MEM_HOST),
              "dvo_set_frames")

    def set_frames_raw(self, frame, bgr, depth_m=None, first=0, count=None, device=False):
        """Raw sensor frames: bgr (n, H, W, 3) u8 and depth_m (n, H, W) f32 metres (dvo_set_frames_raw)."""
        if not device:
            bgr = np.ascontiguousarray(bgr, np.uint8)
            count = bgr.shape[0] if count is None else count
            if depth_m is not None:
                depth_m = np.ascontiguousarray(depth_m, np.float32)
        check(self.lib.dvo_set_frames_raw(self.h, frame, first, count, _ptr(bgr), _ptr(depth_m), MEM_DEVICE if device else MEM_HOST),
              "dvo_set_frames_raw")

    def build_pyramids(self, count, first=0, frames_mask=3):
        check(self.lib.dvo_build_pyramids(self.h, first, count, frames_mask), "dvo_build_pyramids")

    def prepare(self, count, first=0, frames_mask=3):
        check(self.lib.dvo_prepare(self.h, first, count, frames_mask), "dvo_prepare")

    def promote_now_to_ref(self, count, first=0):
        check(self.lib.dvo_promote_now_to_ref(self.h, first, count), "dvo_promote_now_to_ref")

    def set_initial_pose(self, count, poses=None, first=0):
        if poses is not None:
            poses = np.ascontiguousarray(poses, np.float64).reshape(count, 12)
        check(self.lib.dvo_set_initial_pose(self.h, first, count, _ptr(poses), MEM_HOST), "dvo_set_initial_pose")

    def run(self, count, params, first=0):
        check(self.lib.dvo_run(self.h, first, count, C.byref(params)), "dvo_run")

    def process(self, count, params, first=0, poses_out=None):
        """build_pyramids + prepare + run on resident frames; large ranges run as two staggered half batches (dvo_process).
        poses_out: optional raw device pointer (count x 12 doubles) the poses are copied to without joining."""
        check(self.lib.dvo_process(self.h, first, count, C.byref(params), C.c_void_p(poses_out) if poses_out else None), "dvo_process")

    def join(self):
        check(self.lib.dvo_join(self.h), "dvo_join")

    def join_stream(self, stream_ptr):
        check(self.lib.dvo_join_stream(self.h, C.c_void_p(stream_ptr)), "dvo_join_stream")

    def get_poses(self, count, first=0, want_info=True):
        poses = np.empty((count, 12), np.float64)
        info = (PairInfo * count)() if want_info else None
        check(self.lib.dvo_get_poses(self.h, first, count, _ptr(poses), C.cast(info, C.c_void_p) if want_info else None, MEM_HOST),
              "dvo_get_poses")
        return poses, info

    def get_poses_device(self, count, dst_ptr, first=0):
        check(self.lib.dvo_get_poses(self.h, first, count, C.c_void_p(dst_ptr), None, MEM_DEVICE), "dvo_get_poses")

    def align_batch(self, ref_gray, ref_depth, now_gray, params, now_depth=None, want_info=True):
        ref_gray = np.ascontiguousarray(ref_gray, np.uint8)
        ref_depth = np.ascontiguousarray(ref_depth, np.uint16)
        now_gray = np.ascontiguousarray(now_gray, np.uint8)
        n = ref_gray.shape[0]
        poses = np.empty((n, 12), np.float64)
        info = (PairInfo * n)() if want_info else None
        check(self.lib.dvo_align_batch(self.h, n, _ptr(ref_gray), _ptr(ref_depth), _ptr(now_gray), _ptr(now_depth), C.byref(params),
                                       _ptr(poses), C.cast(info, C.c_void_p) if want_info else None), "dvo_align_batch")
        return poses, info

    def run_sequences(self, gray, depth, params, keyframe_every=5):
        """SolveDVO::loop over nseq sequences in lock step.  gray/depth: (nseq, nframes, H, W)."""
        gray = np.ascontiguousarray(gray, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        nseq, nframes = gray.shape[:2]
        rel = np.empty((nseq, nframes, 12), np.float64)
        kind = np.empty((nseq, nframes), np.int32)
        glob = np.empty((nseq, nframes, 19), np.float64)
        check(self.lib.dvo_run_sequences(self.h, nseq, nframes, _ptr(gray), _ptr(depth), C.byref(params), keyframe_every, _ptr(rel),
                                         _ptr(kind), _ptr(glob)), "dvo_run_sequences")
        return rel, kind, glob

    def run_sequences_gated(self, gray, depth, params, policy):
        """SolveDVO::loop with the key-frame decision (periodic rule and / or the quality gates) taken per sequence on the device."""
        gray = np.ascontiguousarray(gray, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        nseq, nframes = gray.shape[:2]
        rel = np.empty((nseq, nframes, 12), np.float64)
        kind = np.empty((nseq, nframes), np.int32)
        reason = np.empty((nseq, nframes), np.int32)
        glob = np.empty((nseq, nframes, 19), np.float64)
        check(self.lib.dvo_run_sequences_gated(self.h, nseq, nframes, _ptr(gray), _ptr(depth), C.byref(params), C.byref(policy), _ptr(rel),
                                               _ptr(kind), _ptr(reason), _ptr(glob)), "dvo_run_sequences_gated")
        return rel, kind, reason, glob

    def run_sequences_mem(self, gray, depth, nseq, nframes, params, policy, device=False, want_global=True):
        """dvo_run_sequences_mem: gray / depth are numpy arrays (host; pinned arrays upload asynchronously) or, with device=True,
        raw device pointers to [nseq][nframes][H][W] buffers."""
        rel = np.empty((nseq, nframes, 12), np.float64)
        kind = np.empty((nseq, nframes), np.int32)
        reason = np.empty((nseq, nframes), np.int32)
        glob = np.empty((nseq, nframes, 19), np.float64) if want_global else None
        check(self.lib.dvo_run_sequences_mem(self.h, nseq, nframes, _ptr(gray), _ptr(depth), MEM_DEVICE if device else MEM_HOST,
                                             C.byref(params), C.byref(policy), _ptr(rel), _ptr(kind), _ptr(reason), _ptr(glob)),
              "dvo_run_sequences_mem")
        return rel, kind, reason, glob

    def level_dims(self, level):
        w, h = C.c_int(), C.c_int()
        check(self.lib.dvo_level_dims(self.h, level, C.byref(w), C.byref(h)), "dvo_level_dims")
        return w.value, h.value

    def get_level_buffer(self, slot, frame, level, which):
        code, dt = BUF[which]
        w, h = self.level_dims(level)
        out = np.empty((h, w), dt)
        check(self.lib.dvo_get_level_buffer(self.h, slot, frame, level, code, _ptr(out), out.nbytes), "dvo_get_level_buffer")
        return out

    def get_points(self, slot, level):
        w, h = self.level_dims(level)
        cap = w * h
        X, Y, Z = (np.empty(cap, np.float32) for _ in range(3))
        n = C.c_int()
        check(self.lib.dvo_get_points(self.h, slot, level, _ptr(X), _ptr(Y), _ptr(Z), cap, C.byref(n)), "dvo_get_points")
        return X[: n.value].copy(), Y[: n.value].copy(), Z[: n.value].copy()

    def eval_normal_equations(self, slot, level, R, T, jacobian=JAC_REFERENCE, weight=W_REF_CAUCHY, arithmetic=ARITH_EXACT,
                              huber_k=1.345, per_point=False, npts=None, residual=RES_DT_FLOOR):
        pose = np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(T, np.float64).reshape(3)])
        H = np.empty(36, np.float64)
        g = np.empty(6, np.float64)
        sumsq, nvis = C.c_double(), C.c_int()
        pp = {}
        if per_point:
            pp = {"eps": np.zeros(npts, np.float32), "w": np.zeros(npts, np.float32), "u": np.zeros(npts, np.float32),
                  "v": np.zeros(npts, np.float32), "J": np.zeros((npts, 6), np.float32)}
        prm = solver_params(jacobian=jacobian, weight=weight, arithmetic=arithmetic, huber_k=huber_k, residual=residual)
        check(self.lib.dvo_eval_normal_equations_ex(self.h, slot, level, _ptr(pose), C.byref(prm), _ptr(H), _ptr(g),
                                                    C.byref(sumsq), C.byref(nvis), _ptr(pp.get("eps")), _ptr(pp.get("w")),
                                                    _ptr(pp.get("u")), _ptr(pp.get("v")), _ptr(pp.get("J"))), "dvo_eval_normal_equations_ex")
        out = {"H": H.reshape(6, 6), "g": g, "sumsq": sumsq.value, "nvis": nvis.value}
        out.update(pp)
        return out

    def get_trace(self, slot, level):
        n = self.cfg.trace_iters
        tr = np.zeros((n, 56), np.float64)
        check(self.lib.dvo_get_trace(self.h, slot, level, _ptr(tr)), "dvo_get_trace")
        return {"g": tr[:, 0:6], "H": tr[:, 6:42].reshape(n, 6, 6), "energy": tr[:, 42], "nvis": tr[:, 43],
                "R": tr[:, 44:53].reshape(n, 3, 3), "T": tr[:, 53:56]}

    def get_energies(self, slot, level, n=128):
        out = np.zeros(n, np.float32)
        check(self.lib.dvo_get_energies(self.h, slot, level, _ptr(out), n), "dvo_get_energies")
        return out

    def enable_timing(self, on=True):
        check(self.lib.dvo_enable_timing(self.h, int(on)), "dvo_enable_timing")

    def stage_ms(self):
        ms = (C.c_float * len(_lib.STAGES))()
        check(self.lib.dvo_get_stage_ms(self.h, ms), "dvo_get_stage_ms")
        return dict(zip(_lib.STAGES, list(ms)))

    def launch_count(self):
        return int(self.lib.dvo_launch_count(self.h))

    def gop_compose(self, kind, rel):
        kind = np.ascontiguousarray(kind, np.int32)
        nseq, nframes = kind.shape
        rel = np.ascontiguousarray(rel, np.float64).reshape(nseq, nframes, 12)
        out = np.empty((nseq, nframes, 19), np.float64)
        check(self.lib.dvo_gop_compose(self.h, nseq, nframes, _ptr(kind), _ptr(rel), _ptr(out), MEM_HOST), "dvo_gop_compose")
        return out

"""ctypes access to the C shim of the C++ host classes (rgbd_odometry_b200/libdvo_host.so)."""
import ctypes as C

import numpy as np

from rgbd_odometry_b200 import build as _build

_h = None


def lib():
    global _h
    if _h is None:
        _h = C.CDLL(_build.build_host())
    return _h


_hcv = None


def build_cv_shim(force=False):
    """tests/cpp/host_capi_cv.cpp compiled against the stand-in OpenCV / Eigen headers (tests/standin_include): exercises the
    `#ifdef DVO_HAVE_OPENCV / DVO_HAVE_EIGEN` overloads of the host classes.  Returns the path of the built library."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cpp", "host_capi_cv.cpp")
    out = os.path.join(root, "tests", "cpp", "libdvo_host_cv.so")
    host = _build.build_host()
    deps = [src, host] + [os.path.join(root, "rgbd_odometry_b200", "host", f) for f in os.listdir(os.path.join(root, "rgbd_odometry_b200", "host")) if f.endswith(".h")]
    if force or not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "rgbd_odometry_b200", "host"),
                               "-I", os.path.join(root, "tests", "standin_include"), "-o", out, src, "-L", os.path.join(root, "rgbd_odometry_b200"),
                               "-ldvo_host", "-ldvo_b200", "-Wl,-rpath,$ORIGIN/../../rgbd_odometry_b200"])
    return out


def cvlib():
    global _hcv
    if _hcv is None:
        _hcv = C.CDLL(build_cv_shim())
    return _hcv


def cv_eposeestimator(ref_bgr, ref_depth, now_bgr, now_depth, K, level, iters=6, huber_k=10.0, lambda0=1e-3):
    H, W = ref_depth.shape
    h, w = H >> level, W >> level
    R = np.zeros(9); T = np.zeros(3); J = np.zeros((6, h * w)); ok = C.c_int()
    c = lambda a: np.ascontiguousarray(a).copy()
    n = cvlib().cvapi_eposeestimator(_p(c(ref_bgr)), _p(c(ref_depth)), _p(c(now_bgr)), _p(c(now_depth)), W, H, C.c_double(K[0]), C.c_double(K[1]),
                                     C.c_double(K[2]), C.c_double(K[3]), level, iters, C.c_double(huber_k), C.c_double(lambda0), _p(R), _p(T), _p(J),
                                     C.byref(ok))
    return {"levels": n, "R": R.reshape(3, 3), "T": T, "J": J.T.copy(), "roundtrip_ok": bool(ok.value)}       # J arrives column-major (Eigen)


def cv_solvedvo_run_iterations(ref_gray, ref_depth, now_gray, now_depth, levels, level, max_iter, K):
    H, W = ref_gray.shape
    R = np.zeros(9); T = np.zeros(3); en = np.zeros(max_iter, np.float32); cap = W * H
    eps = np.zeros(cap, np.float32); ru = np.zeros(cap, np.float32); bi = C.c_int(); vr = C.c_float(); gop = np.zeros(12)
    c = lambda a: np.ascontiguousarray(a).copy()
    n = cvlib().cvapi_solvedvo_run_iterations(_p(c(ref_gray)), _p(c(ref_depth)), _p(c(now_gray)), _p(c(now_depth)), W, H, levels, C.c_float(K[0]),
                                              C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3]), level, max_iter, _p(R), _p(T), _p(en), _p(eps), _p(ru),
                                              C.byref(bi), C.byref(vr), _p(gop))
    return {"R": R.reshape(3, 3), "T": T, "energies": en, "eps": eps[:n], "u": ru[:n], "best_index": bi.value, "visible_ratio": vr.value, "gop": gop}


def _p(a, t=None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def gop_replay(kind, reason, rel, use_float=False):
    n = len(kind)
    kind = np.ascontiguousarray(kind, np.int32); reason = np.ascontiguousarray(reason, np.int32)
    rel = np.ascontiguousarray(rel, np.float64).reshape(n, 12)
    out = np.empty((n, 19)); is_key = np.empty(n, np.int32); r_out = np.empty(n, np.int32)
    m = lib().hostapi_gop_replay(n, _p(kind), _p(reason), _p(rel), _p(out), _p(is_key), _p(r_out), int(use_float))
    assert m == n
    return out, is_key, r_out


def solvedvo_sequence(gray, depth, levels, iters, K):
    n, H, W = gray.shape
    gray = np.ascontiguousarray(gray, np.uint8); depth = np.ascontiguousarray(depth, np.uint16)
    it = np.array(iters, np.int32)
    out = np.empty((n, 19)); is_key = np.empty(n, np.int32); reason = np.empty(n, np.int32)
    rc = lib().hostapi_solvedvo_sequence(_p(gray), _p(depth), n, W, H, levels, C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]),
                                         C.c_float(K[3]), _p(it), _p(out), _p(is_key), _p(reason))
    assert rc == 0
    return out, is_key, reason


def solvedvo_run_iterations(ref_gray, ref_depth, now_gray, now_depth, levels, level, max_iter, K, R0=None, T0=None):
    H, W = ref_gray.shape
    R = np.eye(3).reshape(9).copy() if R0 is None else np.array(R0, np.float64).reshape(9).copy()
    T = np.zeros(3) if T0 is None else np.array(T0, np.float64).copy()
    en = np.zeros(max_iter, np.float32)
    cap = W * H
    eps, ru, rv = (np.zeros(cap, np.float32) for _ in range(3))
    bi, npts = C.c_int(), C.c_int()
    vr, bcap = C.c_float(), C.c_float()
    rc = lib().hostapi_solvedvo_run_iterations(_p(np.ascontiguousarray(ref_gray)), _p(np.ascontiguousarray(ref_depth)),
                                               _p(np.ascontiguousarray(now_gray)), _p(np.ascontiguousarray(now_depth)), W, H, levels,
                                               C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3]), level, max_iter, _p(R), _p(T),
                                               _p(en), _p(eps), _p(ru), _p(rv), C.byref(bi), C.byref(vr), C.byref(npts), C.byref(bcap))
    assert rc == 0
    n = npts.value
    return {"R": R.reshape(3, 3), "T": T, "energies": en, "eps": eps[:n], "u": ru[:n], "v": rv[:n], "best_index": bi.value,
            "visible_ratio": vr.value, "b_cap": bcap.value}


def eposeestimator(ref_bgr, ref_depth, now_bgr, now_depth, K, compat, level, iters=3, huber_k=0.0, lambda0=0.0, R0=None, T0=None):
    H, W = ref_depth.shape
    R = np.eye(3).reshape(9).copy() if R0 is None else np.array(R0, np.float64).reshape(9).copy()
    T = np.zeros(3) if T0 is None else np.array(T0, np.float64).copy()
    A = np.empty((6, 6)); vis = C.c_float(); st = C.c_int()
    h, w = H >> level, W >> level
    J = np.empty((h * w, 6)); X = np.empty((h, w)); g = np.empty((h, w), np.uint8)
    nlev = lib().hostapi_eposeestimator(_p(np.ascontiguousarray(ref_bgr)), _p(np.ascontiguousarray(ref_depth)), _p(np.ascontiguousarray(now_bgr)),
                                        _p(np.ascontiguousarray(now_depth)), W, H, C.c_double(K[0]), C.c_double(K[1]), C.c_double(K[2]),
                                        C.c_double(K[3]), int(compat), level, iters, C.c_double(huber_k), C.c_double(lambda0), _p(R), _p(T), _p(A),
                                        C.byref(vis), C.byref(st), _p(J), _p(X), _p(g))
    return {"R": R.reshape(3, 3), "T": T, "A": A, "visible": vis.value, "status": st.value, "J": J, "X": X, "gray": g, "levels": nlev}


def pydstore_add_level(bgrA, depthA, bgrB, depthB, now_bgr, now_depth, K, level, iters=6, huber_k=10.0, lambda0=1e-3):
    H, W = depthA.shape
    R = np.zeros(9); T = np.zeros(3); A = np.zeros((6, 6)); sizes = np.zeros(3, np.int32)
    c = np.ascontiguousarray
    bad = lib().hostapi_pydstore_add_level(_p(c(bgrA)), _p(c(depthA)), _p(c(bgrB)), _p(c(depthB)), _p(c(now_bgr)), _p(c(now_depth)), W, H,
                                           C.c_double(K[0]), C.c_double(K[1]), C.c_double(K[2]), C.c_double(K[3]), level, iters,
                                           C.c_double(huber_k), C.c_double(lambda0), _p(R), _p(T), _p(A), _p(sizes))
    return bad, R.reshape(3, 3), T, A, sizes


# ---- FrameIO ----
def store_frame_xml(path, mono_levels, depth_levels):
    H, W = mono_levels[0].shape
    mono = np.concatenate([np.ascontiguousarray(m, np.uint8).ravel() for m in mono_levels])
    depth = np.concatenate([np.ascontiguousarray(d, np.uint16).ravel() for d in depth_levels])
    return lib().hostapi_store_frame_xml(str(path).encode(), _p(mono), _p(depth), W, H, len(mono_levels))


def load_frame_xml(path, levels, capacity=1 << 22):
    mono = np.zeros(capacity, np.uint8); depth = np.zeros(capacity, np.uint16); dims = np.zeros(2 * levels, np.int32)
    rc = lib().hostapi_load_frame_xml(str(path).encode(), levels, _p(mono), _p(depth), C.c_size_t(capacity), _p(dims))
    if rc != 0:
        return rc, None, None
    ms, ds, off = [], [], 0
    for l in range(levels):
        r, c = int(dims[2 * l]), int(dims[2 * l + 1])
        ms.append(mono[off:off + r * c].reshape(r, c).copy()); ds.append(depth[off:off + r * c].reshape(r, c).copy())
        off += r * c
    return 0, ms, ds


def read_xml_matrix(path, name, capacity=1 << 20):
    out = np.zeros(capacity, np.float64)
    r, c, ch, el = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib().hostapi_read_xml_matrix(str(path).encode(), name.encode(), _p(out), C.c_size_t(capacity), C.byref(r), C.byref(c), C.byref(ch), C.byref(el))
    if rc != 0:
        return rc, None, None
    shape = (r.value, c.value, ch.value) if ch.value > 1 else (r.value, c.value)
    return 0, out[: r.value * c.value * ch.value].reshape(shape).copy(), el.value


def write_pose_file(path, q7):
    q7 = np.ascontiguousarray(q7, np.float64).reshape(-1, 7)
    return lib().hostapi_write_pose_file(str(path).encode(), _p(q7), len(q7))


def read_pose_file(path, capacity=100000):
    out = np.zeros((capacity, 7), np.float64)
    n = lib().hostapi_read_pose_file(str(path).encode(), _p(out), capacity)
    return None if n < 0 else out[:n].copy()


def load_gt_path(path, skip=350, capacity=100000):
    out = np.zeros((capacity, 7), np.float64)
    n = lib().hostapi_load_gt_path(str(path).encode(), int(skip), _p(out), capacity)
    return None if n < 0 else out[:n].copy()


def solvedvo_from_files(ref_xml, now_xml, W, H, levels, iters, K):
    it = np.array(iters, np.int32)
    R = np.zeros(9); T = np.zeros(3)
    rc = lib().hostapi_solvedvo_from_files(str(ref_xml).encode(), str(now_xml).encode(), W, H, levels, C.c_float(K[0]), C.c_float(K[1]),
                                           C.c_float(K[2]), C.c_float(K[3]), _p(it), _p(R), _p(T))
    assert rc == 0, rc
    return R.reshape(3, 3), T


def solvedvo_loop_xml(folder, start, end, W, H, levels, iters, K, dry=False, capacity=256):
    it = np.array(iters, np.int32)
    out = np.zeros((capacity, 19)); is_key = np.zeros(capacity, np.int32); reason = np.zeros(capacity, np.int32)
    n = lib().hostapi_solvedvo_loop_xml(str(folder).encode(), start, end, int(dry), W, H, levels, C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]),
                                        C.c_float(K[3]), _p(it), _p(out), _p(is_key), _p(reason), capacity)
    return n, out[:max(0, min(n, capacity))], is_key[:max(0, min(n, capacity))], reason[:max(0, min(n, capacity))]


def solvedvo_loop_callback(gray, depth, levels, iters, K, skip_every=0):
    n, H, W = gray.shape
    it = np.array(iters, np.int32)
    out = np.zeros((n, 19))
    m = lib().hostapi_solvedvo_loop_callback(_p(np.ascontiguousarray(gray, np.uint8)), _p(np.ascontiguousarray(depth, np.uint16)), n, skip_every, W, H,
                                             levels, C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3]), _p(it), _p(out))
    return m, out


def solvedvo_casual(folder, ref_index, now_index, iterations, W, H, levels, K):
    R = np.zeros(9); T = np.zeros(3); en = np.zeros(max(iterations, 1), np.float32)
    n = lib().hostapi_solvedvo_casual(str(folder).encode(), ref_index, now_index, iterations, W, H, levels, C.c_float(K[0]), C.c_float(K[1]),
                                      C.c_float(K[2]), C.c_float(K[3]), _p(R), _p(T), _p(en), len(en))
    return n, R.reshape(3, 3), T, en


def solvedvo_loop_from_file(folder, start, end, iterations, W, H, levels, K, capacity=64):
    out = np.zeros((capacity, 12))
    n = lib().hostapi_solvedvo_loop_from_file(str(folder).encode(), start, end, iterations, W, H, levels, C.c_float(K[0]), C.c_float(K[1]),
                                              C.c_float(K[2]), C.c_float(K[3]), _p(out), capacity)
    return n, out[:max(0, min(n, capacity))]


# ---- RGBDOdometry host class ----
def rgbdodometry_sequence(bgr, depth, K, ref_every=10000):
    n, H, W = depth.shape
    bgr = np.ascontiguousarray(bgr, np.uint8); depth = np.ascontiguousarray(depth, np.uint16)
    out = np.empty((n, 4, 4)); npts = np.empty(n, np.int32); eps = np.empty(n, np.float64)
    rc = lib().hostapi_rgbdodometry_sequence(_p(bgr), _p(depth), n, W, H, C.c_double(K[0]), C.c_double(K[1]), C.c_double(K[2]), C.c_double(K[3]),
                                             int(ref_every), _p(out), _p(npts), _p(eps))
    assert rc == 0
    return out, npts, eps


def rgbdodometry_materialise(ref_bgr, ref_depth, now_bgr, now_depth, K, level, T):
    H, W = ref_depth.shape
    h, w = H >> level, W >> level
    cap = h * w
    J = np.zeros((cap, 6)); marks = np.zeros((h, w), np.int32); eps = np.zeros(cap); roi = np.zeros((h, w), np.int32)
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    n = lib().hostapi_rgbdodometry_materialise(_p(np.ascontiguousarray(ref_bgr)), _p(np.ascontiguousarray(ref_depth)), _p(np.ascontiguousarray(now_bgr)),
                                               _p(np.ascontiguousarray(now_depth)), W, H, C.c_double(K[0]), C.c_double(K[1]), C.c_double(K[2]),
                                               C.c_double(K[3]), level, _p(T), _p(J), _p(marks), _p(eps), _p(roi), cap)
    assert n >= 0
    return J[:n].copy(), marks, eps[:n].copy(), roi

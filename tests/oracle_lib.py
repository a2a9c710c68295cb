"""ctypes bindings for the CPU oracle (oracle/libdvo_oracle.so) and the synthetic renderer.

Test infrastructure: imported only from tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke().  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_SO = os.path.join(ROOT, "oracle", "libdvo_oracle.so")
_SYNTH_SO = os.path.join(ROOT, "rgbd_odometry_b200", "synth", "libdvo_synth.so")


def build_oracle(force=False):
    src = [os.path.join(ROOT, "oracle", f) for f in ("dvo_oracle_c.cpp", "dvo_oracle.hpp")]
    if force or not os.path.exists(_ORACLE_SO) or any(os.path.getmtime(s) > os.path.getmtime(_ORACLE_SO) for s in src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    return _ORACLE_SO


def build_synth(force=False):
    src = os.path.join(ROOT, "rgbd_odometry_b200", "synth", "synth.cpp")
    if force or not os.path.exists(_SYNTH_SO) or os.path.getmtime(src) > os.path.getmtime(_SYNTH_SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _SYNTH_SO, src])
    return _SYNTH_SO


class SolverCfg(C.Structure):
    _fields_ = [("solver", C.c_int), ("jacobian", C.c_int), ("weight", C.c_int), ("huber_k", C.c_float),
                ("lm_lambda0", C.c_double), ("residual", C.c_int)]


SUBGRAD_REF, GN, LM = 0, 1, 2
JAC_REFERENCE, JAC_EXACT = 0, 1
W_REF_CAUCHY, W_HUBER, W_NONE = 0, 1, 2


RES_DT_FLOOR, RES_DT_INTERP = 0, 1


def cfg(solver=SUBGRAD_REF, jacobian=JAC_REFERENCE, weight=W_REF_CAUCHY, huber_k=1.345, lm_lambda0=1e-3, residual=RES_DT_FLOOR):
    return SolverCfg(solver, jacobian, weight, huber_k, lm_lambda0, residual)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


_lib = None
_syn = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_edt_d2.restype = C.c_size_t
        _lib.orc_dt_normalize.restype = C.c_float
        _lib.orc_align_batch.restype = C.c_double
        _lib.orc_weight_ref.restype = C.c_float
        _lib.orc_weight_ref.argtypes = [C.c_float]
        _lib.orc_laplacian_b.restype = C.c_float
    return _lib


def synth():
    global _syn
    if _syn is None:
        _syn = C.CDLL(build_synth())
    return _syn


K640 = (525.0, 525.0, 319.5, 239.5)
K1280 = (1050.0, 1050.0, 639.5, 359.5)


def synth_pair(seed, W=640, H=480, K=K640, max_angle_deg=1.0, max_trans_m=0.02, bgr=False):
    P = W * H
    out = {
        "ref_gray": np.empty((H, W), np.uint8), "ref_depth": np.empty((H, W), np.uint16),
        "now_gray": np.empty((H, W), np.uint8), "now_depth": np.empty((H, W), np.uint16),
        "R": np.empty(9, np.float64), "T": np.empty(3, np.float64),
    }
    rb = np.empty((H, W, 3), np.uint8) if bgr else None
    nb = np.empty((H, W, 3), np.uint8) if bgr else None
    synth().dvo_synth_pair(C.c_uint64(seed), W, H, C.c_double(K[0]), C.c_double(K[1]), C.c_double(K[2]), C.c_double(K[3]),
                           C.c_double(max_angle_deg), C.c_double(max_trans_m), _p(rb, C.c_uint8), _p(out["ref_gray"], C.c_uint8),
                           _p(out["ref_depth"], C.c_uint16), _p(nb, C.c_uint8), _p(out["now_gray"], C.c_uint8),
                           _p(out["now_depth"], C.c_uint16), _p(out["R"], C.c_double), _p(out["T"], C.c_double))
    if bgr:
        out["ref_bgr"], out["now_bgr"] = rb, nb
    out["R"] = out["R"].reshape(3, 3)
    return out


def synth_batch(seed0, count, W=640, H=480, K=K640, max_angle_deg=1.0, max_trans_m=0.02, nthreads=None, bgr=False,
                now_depth=False):
    nthreads = nthreads or os.cpu_count() or 1
    out = {
        "ref_gray": np.empty((count, H, W), np.uint8), "ref_depth": np.empty((count, H, W), np.uint16),
        "now_gray": np.empty((count, H, W), np.uint8),
        "R": np.empty((count, 3, 3), np.float64), "T": np.empty((count, 3), np.float64),
    }
    nd = np.empty((count, H, W), np.uint16) if now_depth else None
    rb = np.empty((count, H, W, 3), np.uint8) if bgr else None
    nb = np.empty((count, H, W, 3), np.uint8) if bgr else None
    synth().dvo_synth_batch(C.c_uint64(seed0), count, W, H, C.c_double(K[0]), C.c_double(K[1]), C.c_double(K[2]),
                            C.c_double(K[3]), C.c_double(max_angle_deg), C.c_double(max_trans_m), _p(rb, C.c_uint8),
                            _p(out["ref_gray"], C.c_uint8), _p(out["ref_depth"], C.c_uint16), _p(nb, C.c_uint8),
                            _p(out["now_gray"], C.c_uint8), _p(nd, C.c_uint16), _p(out["R"], C.c_double),
                            _p(out["T"], C.c_double), nthreads)
    if now_depth:
        out["now_depth"] = nd
    if bgr:
        out["ref_bgr"], out["now_bgr"] = rb, nb
    return out


def synth_sequence(seed, nframes, W=640, H=480, K=K640, max_angle_deg=0.5, max_trans_m=0.01):
    gray = np.empty((nframes, H, W), np.uint8)
    depth = np.empty((nframes, H, W), np.uint16)
    R = np.empty((nframes, 3, 3), np.float64)
    T = np.empty((nframes, 3), np.float64)
    synth().dvo_synth_sequence(C.c_uint64(seed), nframes, W, H, C.c_double(K[0]), C.c_double(K[1]), C.c_double(K[2]),
                               C.c_double(K[3]), C.c_double(max_angle_deg), C.c_double(max_trans_m), _p(gray, C.c_uint8),
                               _p(depth, C.c_uint16), None, _p(R, C.c_double), _p(T, C.c_double))
    return gray, depth, R, T


# ---------------------------------------------------------------- oracle stage wrappers
def level_dim(dim, level):
    return lib().orc_level_dim(dim, level)


def pyr_nearest(img, level):
    H, W = img.shape[:2]
    h, w = level_dim(H, level), level_dim(W, level)
    img = np.ascontiguousarray(img)
    if img.dtype == np.uint16:
        out = np.empty((h, w), np.uint16)
        lib().orc_pyr_nearest_u16(_p(img, C.c_uint16), W, H, level, _p(out, C.c_uint16))
    else:
        ch = 1 if img.ndim == 2 else img.shape[2]
        out = np.empty((h, w) if ch == 1 else (h, w, ch), np.uint8)
        lib().orc_pyr_nearest_u8(_p(img, C.c_uint8), W, H, level, _p(out, C.c_uint8), ch)
    return out


def pyr_area(img, level):
    H, W = img.shape[:2]
    s = 1 << level
    h, w = H // s, W // s
    img = np.ascontiguousarray(img)
    if img.dtype == np.uint16:
        out = np.empty((h, w), np.uint16)
        lib().orc_pyr_area_u16(_p(img, C.c_uint16), W, H, level, _p(out, C.c_uint16))
    else:
        ch = 1 if img.ndim == 2 else img.shape[2]
        out = np.empty((h, w) if ch == 1 else (h, w, ch), np.uint8)
        lib().orc_pyr_area_u8(_p(img, C.c_uint8), W, H, level, _p(out, C.c_uint8), ch)
    return out


def bgr2gray(bgr):
    bgr = np.ascontiguousarray(bgr)
    out = np.empty(bgr.shape[:2], np.uint8)
    lib().orc_bgr2gray(_p(bgr, C.c_uint8), C.c_size_t(out.size), _p(out, C.c_uint8))
    return out


def depth_m_to_mm(m):
    m = np.ascontiguousarray(m, np.float32)
    out = np.empty(m.shape, np.uint16)
    lib().orc_depth_m_to_mm(_p(m, C.c_float), C.c_size_t(m.size), _p(out, C.c_uint16))
    return out


def canny(img):
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_canny(_p(img, C.c_uint8), W, H, _p(out, C.c_uint8))
    return out


def edt_d2(edge, brute=False):
    edge = np.ascontiguousarray(edge, np.uint8)
    H, W = edge.shape
    out = np.empty((H, W), np.int32)
    if brute:
        lib().orc_edt_d2_brute(_p(edge, C.c_uint8), W, H, _p(out, C.c_int32))
    else:
        lib().orc_edt_d2(_p(edge, C.c_uint8), W, H, _p(out, C.c_int32))
    return out


def dt_normalize(d2):
    d2 = np.ascontiguousarray(d2, np.int32)
    out = np.empty(d2.shape, np.float32)
    s = lib().orc_dt_normalize(_p(d2, C.c_int32), C.c_size_t(d2.size), _p(out, C.c_float))
    return out, s


def gradient(img):
    img = np.ascontiguousarray(img, np.float32)
    H, W = img.shape
    gx = np.empty((H, W), np.float32)
    gy = np.empty((H, W), np.float32)
    lib().orc_gradient(_p(img, C.c_float), W, H, _p(gx, C.c_float), _p(gy, C.c_float))
    return gx, gy


def select_points(edge, depth, level, K=K640):
    edge = np.ascontiguousarray(edge, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = edge.shape
    cap = W * H
    X, Y, Z, u, v = (np.empty(cap, np.float32) for _ in range(5))
    n = lib().orc_select_points(_p(edge, C.c_uint8), _p(depth, C.c_uint16), W, H, level, C.c_float(K[0]), C.c_float(K[1]),
                                C.c_float(K[2]), C.c_float(K[3]), _p(X, C.c_float), _p(Y, C.c_float), _p(Z, C.c_float),
                                _p(u, C.c_float), _p(v, C.c_float), cap)
    return X[:n].copy(), Y[:n].copy(), Z[:n].copy(), u[:n].copy(), v[:n].copy()


def preprocess_level(gray, depth, level):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    h, w = level_dim(H, level), level_dim(W, level)
    o = {"gray": np.empty((h, w), np.uint8), "edge": np.empty((h, w), np.uint8), "d2": np.empty((h, w), np.int32),
         "dtn": np.empty((h, w), np.float32), "gx": np.empty((h, w), np.float32), "gy": np.empty((h, w), np.float32)}
    dl = None
    if depth is not None:
        depth = np.ascontiguousarray(depth, np.uint16)
        dl = np.empty((h, w), np.uint16)
        o["depth"] = dl
    lib().orc_preprocess_level(_p(gray, C.c_uint8), _p(depth, C.c_uint16), W, H, level, _p(o["gray"], C.c_uint8),
                               _p(dl, C.c_uint16), _p(o["edge"], C.c_uint8), _p(o["d2"], C.c_int32), _p(o["dtn"], C.c_float),
                               _p(o["gx"], C.c_float), _p(o["gy"], C.c_float))
    return o


def evaluate(X, Y, Z, dtn, gx, gy, level, R, T, K=K640, jac=JAC_REFERENCE, weight=W_REF_CAUCHY, huber_k=1.345,
             per_point=False, residual=RES_DT_FLOOR):
    N = len(X)
    H, W = dtn.shape
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    T = np.ascontiguousarray(T, np.float64).reshape(3)
    g = np.empty(6, np.float64)
    Hm = np.empty(36, np.float64)
    sumsq = C.c_double()
    nvis = C.c_int()
    pp = {}
    if per_point:
        pp = {"eps": np.empty(N, np.float32), "w": np.empty(N, np.float32), "u": np.empty(N, np.float32),
              "v": np.empty(N, np.float32), "J": np.empty((N, 6), np.float32)}
    lib().orc_evaluate(_p(np.ascontiguousarray(X), C.c_float), _p(np.ascontiguousarray(Y), C.c_float),
                       _p(np.ascontiguousarray(Z), C.c_float), N, _p(np.ascontiguousarray(dtn), C.c_float),
                       _p(np.ascontiguousarray(gx), C.c_float), _p(np.ascontiguousarray(gy), C.c_float), W, H, level,
                       C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3]), _p(R, C.c_double), _p(T, C.c_double),
                       jac, weight, C.c_float(huber_k), _p(g, C.c_double), _p(Hm, C.c_double), C.byref(sumsq), C.byref(nvis),
                       _p(pp.get("eps"), C.c_float), _p(pp.get("w"), C.c_float), _p(pp.get("u"), C.c_float),
                       _p(pp.get("v"), C.c_float), _p(pp.get("J"), C.c_float), int(residual))
    out = {"g": g, "H": Hm.reshape(6, 6), "sumsq": sumsq.value, "nvis": nvis.value}
    out.update(pp)
    return out


def parse_trace(tr):
    """tr: (..., 56) doubles -> dict of arrays."""
    return {"g": tr[..., 0:6], "H": tr[..., 6:42].reshape(tr.shape[:-1] + (6, 6)), "energy": tr[..., 42],
            "nvis": tr[..., 43], "R": tr[..., 44:53].reshape(tr.shape[:-1] + (3, 3)), "T": tr[..., 53:56]}


def run_iterations(X, Y, Z, dtn, gx, gy, level, max_iter, R0=None, T0=None, K=K640, scfg=None, trace=False):
    N = len(X)
    H, W = dtn.shape
    R = np.eye(3).reshape(9).copy() if R0 is None else np.array(R0, np.float64).reshape(9).copy()
    T = np.zeros(3) if T0 is None else np.array(T0, np.float64).reshape(3).copy()
    scfg = scfg or cfg()
    energies = np.zeros(max_iter, np.float32)
    bi, ir = C.c_int(), C.c_int()
    vr = C.c_float()
    be, bu, bv = (np.zeros(N, np.float32) for _ in range(3))
    tr = np.zeros((max_iter, 56), np.float64) if trace else None
    lib().orc_run_iterations(_p(np.ascontiguousarray(X), C.c_float), _p(np.ascontiguousarray(Y), C.c_float),
                             _p(np.ascontiguousarray(Z), C.c_float), N, _p(np.ascontiguousarray(dtn), C.c_float),
                             _p(np.ascontiguousarray(gx), C.c_float), _p(np.ascontiguousarray(gy), C.c_float), W, H, level,
                             max_iter, C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3]), C.byref(scfg),
                             _p(R, C.c_double), _p(T, C.c_double), _p(energies, C.c_float), C.byref(bi), C.byref(vr),
                             C.byref(ir), _p(be, C.c_float), _p(bu, C.c_float), _p(bv, C.c_float), _p(tr, C.c_double))
    out = {"R": R.reshape(3, 3), "T": T, "energies": energies, "best_index": bi.value, "visible_ratio": vr.value,
           "iterations_run": ir.value, "best_eps": be, "best_u": bu, "best_v": bv}
    if trace:
        out["trace"] = parse_trace(tr[: ir.value])
    return out


def align_pair(ref_gray, ref_depth, now_gray, levels=4, iters=(50, 50, 50, 50), K=K640, scfg=None, R0=None, T0=None,
               trace=False, stats=False):
    H, W = ref_gray.shape
    scfg = scfg or cfg()
    it = np.array(iters, np.int32)
    R, T = np.empty(9, np.float64), np.empty(3, np.float64)
    npts, bi, ir = (np.zeros(levels, np.int32) for _ in range(3))
    be, vr = np.zeros(levels, np.float32), np.zeros(levels, np.float32)
    mi = int(max(iters))
    tr = np.zeros((levels, mi, 56), np.float64) if trace else None
    R0p = None if R0 is None else np.ascontiguousarray(R0, np.float64).reshape(9)
    T0p = None if T0 is None else np.ascontiguousarray(T0, np.float64).reshape(3)
    bcap = C.c_float()
    st = lib().orc_align_pair(_p(np.ascontiguousarray(ref_gray), C.c_uint8), _p(np.ascontiguousarray(ref_depth), C.c_uint16),
                              _p(np.ascontiguousarray(now_gray), C.c_uint8), W, H, levels, C.c_float(K[0]), C.c_float(K[1]),
                              C.c_float(K[2]), C.c_float(K[3]), _p(it, C.c_int32), C.byref(scfg), _p(R0p, C.c_double),
                              _p(T0p, C.c_double), _p(R, C.c_double), _p(T, C.c_double), _p(npts, C.c_int32), _p(bi, C.c_int32),
                              _p(ir, C.c_int32), _p(be, C.c_float), _p(vr, C.c_float), _p(tr, C.c_double), mi,
                              C.byref(bcap) if stats else None)
    out = {"R": R.reshape(3, 3), "T": T, "npts": npts, "best_index": bi, "iterations_run": ir, "best_energy": be,
           "visible_ratio": vr, "status": st}
    if stats:
        out["b_cap"] = bcap.value
    if trace:
        out["trace"] = parse_trace(tr)
    return out


def align_batch(ref_gray, ref_depth, now_gray, levels=4, iters=(50, 50, 50, 50), K=K640, scfg=None, nthreads=1):
    n, H, W = ref_gray.shape
    scfg = scfg or cfg()
    it = np.array(iters, np.int32)
    poses = np.empty((n, 12), np.float64)
    status = np.zeros(n, np.int32)
    secs = lib().orc_align_batch(_p(np.ascontiguousarray(ref_gray), C.c_uint8), _p(np.ascontiguousarray(ref_depth), C.c_uint16),
                                 _p(np.ascontiguousarray(now_gray), C.c_uint8), n, W, H, levels, C.c_float(K[0]),
                                 C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3]), _p(it, C.c_int32), C.byref(scfg),
                                 _p(poses, C.c_double), _p(status, C.c_int32), nthreads)
    return poses, status, secs


def se3_exp(psi):
    psi = np.ascontiguousarray(psi, np.float64)
    R, t = np.empty(9), np.empty(3)
    lib().orc_se3_exp(_p(psi, C.c_double), _p(R, C.c_double), _p(t, C.c_double))
    return R.reshape(3, 3), t


def se3_log(R, t):
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    t = np.ascontiguousarray(t, np.float64)
    psi = np.empty(6)
    lib().orc_se3_log(_p(R, C.c_double), _p(t, C.c_double), _p(psi, C.c_double))
    return psi


def rotationize(R):
    R = np.array(R, np.float64).reshape(9).copy()
    lib().orc_rotationize(_p(R, C.c_double))
    return R.reshape(3, 3)


def rot_to_quat(R):
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    q = np.empty(4)
    lib().orc_rot_to_quat(_p(R, C.c_double), _p(q, C.c_double))
    return q


def gop_replay(kind, reason, rel):
    n = len(kind)
    kind = np.ascontiguousarray(kind, np.int32)
    reason = np.ascontiguousarray(reason, np.int32)
    rel = np.ascontiguousarray(rel, np.float64).reshape(n, 12)
    out = np.empty((n, 19))
    is_key, r_out = np.empty(n, np.int32), np.empty(n, np.int32)
    lib().orc_gop_replay(n, _p(kind, C.c_int32), _p(reason, C.c_int32), _p(rel, C.c_double), _p(out, C.c_double),
                         _p(is_key, C.c_int32), _p(r_out, C.c_int32))
    return out, is_key, r_out


# ---------------------------------------------------------------- EPoseEstimator (photometric) oracle
def photo_build_ref_level(bgr, depth, level, K=K640, compat=True):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    s = 1 << level
    h, w = H // s, W // s
    K4 = np.array(K, np.float64)
    o = {"gray": np.empty((h, w), np.uint8), "bgr": np.empty((h, w, 3), np.uint8), "depth": np.empty((h, w), np.uint16),
         "X": np.empty((h, w)), "Y": np.empty((h, w)), "Z": np.empty((h, w)), "gx": np.empty((h, w)), "gy": np.empty((h, w)),
         "J": np.empty((h * w, 6)), "A": np.empty((6, 6))}
    lib().orc_photo_build_ref_level(_p(bgr, C.c_uint8), _p(depth, C.c_uint16), W, H, level, _p(K4, C.c_double), int(compat),
                                    _p(o["gray"], C.c_uint8), _p(o["bgr"], C.c_uint8), _p(o["depth"], C.c_uint16), _p(o["X"], C.c_double),
                                    _p(o["Y"], C.c_double), _p(o["Z"], C.c_double), _p(o["gx"], C.c_double), _p(o["gy"], C.c_double),
                                    _p(o["J"], C.c_double), _p(o["A"], C.c_double))
    return o


def photo_now_level(bgr, level):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    H, W = bgr.shape[:2]
    s = 1 << level
    out = np.empty((H // s, W // s), np.uint8)
    lib().orc_photo_now_level(_p(bgr, C.c_uint8), W, H, level, _p(out, C.c_uint8))
    return out


def _tr16(R, T):
    Tr = np.eye(4)
    Tr[:3, :3] = np.asarray(R, np.float64).reshape(3, 3)
    Tr[:3, 3] = np.asarray(T, np.float64).reshape(3)
    return np.ascontiguousarray(Tr)


def photo_evaluate(now_gray_l, R, T, K=K640, compat=True, huber_k=0.0, want_canvas=False):
    """Uses the level built by the last photo_build_ref_level call."""
    now_gray_l = np.ascontiguousarray(now_gray_l, np.uint8)
    K4 = np.array(K, np.float64)
    Tr = _tr16(R, T)
    b, A = np.empty(6), np.empty((6, 6))
    sumsq, nre, nus = C.c_double(), C.c_int(), C.c_int()
    cv = np.empty(now_gray_l.shape, np.float64) if want_canvas else None
    lib().orc_photo_evaluate(_p(now_gray_l, C.c_uint8), _p(K4, C.c_double), _p(Tr, C.c_double), int(compat), C.c_double(huber_k),
                             _p(b, C.c_double), _p(A, C.c_double), C.byref(sumsq), C.byref(nre), C.byref(nus), _p(cv, C.c_double))
    return {"b": b, "A": A, "sumsq": sumsq.value, "nreproj": nre.value, "nused": nus.value, "canvas": cv}


def photo_estimate(now_gray_l, R0, T0, iters, K=K640, compat=False, huber_k=0.0, lambda0=0.0):
    now_gray_l = np.ascontiguousarray(now_gray_l, np.uint8)
    K4 = np.array(K, np.float64)
    R0 = np.ascontiguousarray(R0, np.float64).reshape(9)
    T0 = np.ascontiguousarray(T0, np.float64).reshape(3)
    R, T = np.empty(9), np.empty(3)
    s0, s1, vis = C.c_double(), C.c_double(), C.c_double()
    ir, st = C.c_int(), C.c_int()
    lib().orc_photo_estimate(_p(now_gray_l, C.c_uint8), _p(K4, C.c_double), _p(R0, C.c_double), _p(T0, C.c_double), iters, int(compat),
                             C.c_double(huber_k), C.c_double(lambda0), _p(R, C.c_double), _p(T, C.c_double), C.byref(s0), C.byref(s1),
                             C.byref(ir), C.byref(st), C.byref(vis))
    return {"R": R.reshape(3, 3), "T": T, "sumsq_first": s0.value, "sumsq_last": s1.value, "iters_run": ir.value, "status": st.value,
            "visible": vis.value}


# ---- RGBDOdometry (semi-dense photometric GN) ----
def rgbd_level(bgr, depth, level):
    H, W = depth.shape
    h, w = level_dim(H, level), level_dim(W, level)
    g = np.empty((h, w), np.uint8); d = np.empty((h, w), np.uint16)
    lib().orc_rgbd_level(_p(np.ascontiguousarray(bgr), C.c_uint8), _p(np.ascontiguousarray(depth), C.c_uint16), W, H, level, _p(g, C.c_uint8), _p(d, C.c_uint16))
    return g, d


def rgbd_jacobian(bgr, depth, level, K, thresh=5, max_j=1 << 30, min_pts=100):
    H, W = depth.shape
    cap = level_dim(H, level) * level_dim(W, level)
    J = np.empty((cap, 6), np.float64); ij = np.empty((cap, 2), np.int32); A = np.empty(36, np.float64); st = C.c_int()
    K4 = np.array(K, np.float64)
    lib().orc_rgbd_jacobian.restype = C.c_int
    n = lib().orc_rgbd_jacobian(_p(np.ascontiguousarray(bgr), C.c_uint8), _p(np.ascontiguousarray(depth), C.c_uint16), W, H, level, _p(K4, C.c_double),
                                thresh, max_j, min_pts, _p(J, C.c_double), _p(ij, C.c_int32), cap, _p(A, C.c_double), C.byref(st))
    return {"J": J[:n].copy(), "ij": ij[:n].copy(), "A": A.reshape(6, 6), "status": st.value}


def rgbd_epsilon(ref_bgr, ref_depth, now_bgr, now_depth, level, K, T, thresh=5):
    H, W = ref_depth.shape
    cap = level_dim(H, level) * level_dim(W, level)
    eps = np.empty(cap, np.float64); uv = np.empty((cap, 2), np.int32); b = np.empty(6, np.float64); ss = C.c_double(); nv = C.c_int()
    K4 = np.array(K, np.float64); T16 = np.ascontiguousarray(T, np.float64).reshape(16)
    lib().orc_rgbd_epsilon.restype = C.c_int
    n = lib().orc_rgbd_epsilon(_p(np.ascontiguousarray(ref_bgr), C.c_uint8), _p(np.ascontiguousarray(ref_depth), C.c_uint16),
                               _p(np.ascontiguousarray(now_bgr), C.c_uint8), _p(np.ascontiguousarray(now_depth), C.c_uint16), W, H, level,
                               _p(K4, C.c_double), thresh, _p(T16, C.c_double), _p(eps, C.c_double), _p(uv, C.c_int32), cap, _p(b, C.c_double),
                               C.byref(ss), C.byref(nv))
    return {"eps": eps[:n].copy(), "uv": uv[:n].copy(), "b": b, "sumsq": ss.value, "nvis": nv.value}


def rgbd_exponential_map(psi):
    out = np.empty(16, np.float64)
    lib().orc_rgbd_exponential_map(_p(np.ascontiguousarray(psi, np.float64), C.c_double), _p(out, C.c_double))
    return out.reshape(4, 4)


def rgbd_qr_solve6(A, b):
    x = np.empty(6, np.float64)
    lib().orc_rgbd_qr_solve6(_p(np.ascontiguousarray(A, np.float64).reshape(36), C.c_double), _p(np.ascontiguousarray(b, np.float64), C.c_double), _p(x, C.c_double))
    return x


def rgbd_gauss_newton(ref_bgr, ref_depth, now_bgr, now_depth, K, levels=(3, 2), iters=3, eps_exit=200.0, thresh=5, T0=None):
    H, W = ref_depth.shape
    T = np.eye(4).reshape(16).copy() if T0 is None else np.array(T0, np.float64).reshape(16).copy()
    lv = np.array(levels, np.int32); info = np.zeros((len(levels), 6), np.float64); K4 = np.array(K, np.float64)
    lib().orc_rgbd_gauss_newton(_p(np.ascontiguousarray(ref_bgr), C.c_uint8), _p(np.ascontiguousarray(ref_depth), C.c_uint16),
                                _p(np.ascontiguousarray(now_bgr), C.c_uint8), _p(np.ascontiguousarray(now_depth), C.c_uint16), W, H, _p(K4, C.c_double),
                                thresh, _p(lv, C.c_int32), len(levels), iters, C.c_double(eps_exit), _p(T, C.c_double), _p(info, C.c_double))
    return T.reshape(4, 4), info


# ---- publisher front-end ----
def undistort(img, K, D):
    img = np.ascontiguousarray(img)
    H, W = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    K4 = np.array(K, np.float64); D5 = np.array(D, np.float64)
    lib().orc_undistort(img.ctypes.data_as(C.c_void_p), W, H, cn, 0 if img.dtype == np.uint8 else 1, _p(K4, C.c_double), _p(D5, C.c_double),
                        out.ctypes.data_as(C.c_void_p))
    return out

"""numpy transliteration of the reference's Eigen expressions, written FROM THE REFERENCE SOURCE (not from oracle/dvo_oracle.hpp), so
that the C++ oracle's restatement of these functions is cross-checked by an independent reading:

  * SolveDVO::enlistRefEdgePts          /root/reference/src/SolveDVO.cpp:224-264
  * SolveDVO::computeJacobianOfNowFrame /root/reference/src/SolveDVO.cpp:306-414   (+ to_se_3 :1104-1114)
  * SolveDVO::getReprojectedEpsilons    /root/reference/src/SolveDVO.cpp:425-462   (+ getWeightOf :1047-1053)
  * the weighted sub-gradient g         /root/reference/src/SolveDVO.cpp:714-720, :777
  * EPoseEstimator::evaluate3d          /root/reference/src/EPoseEstimator.cpp:439-477
  * EPoseEstimator::evaluateJacobian    /root/reference/src/EPoseEstimator.cpp:320-430

Matrix products go through numpy (BLAS order / FMA differ from Eigen's kernels and the reference itself was built with -ffast-math),
so the SolveDVO quantities are compared with a tolerance; the EPoseEstimator arrays are element-wise fp64 expressions evaluated
in the source's order and are expected to agree to the last bit.  Test infrastructure only.
"""
import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------ SolveDVO
def enlist_ref_edge_pts(level, mask, ref_depth, fx, fy, cx, cy):
    """:224-264.  mask (rows, cols) int, ref_depth (rows, cols) float32 mm.  Returns _3d (3, N), _2d (2, N) float32."""
    scaleFac = f32(2.0 ** (-level))                                  # (float)pow(2,-level)
    tmpfx = f32(1.0 / float(scaleFac * f32(fx)))                     # float tmpfx = 1./(scaleFac*fx)  (double division, float store)
    tmpfy = f32(1.0 / float(scaleFac * f32(fy)))
    tmpcx = scaleFac * f32(cx)
    tmpcy = scaleFac * f32(cy)
    xs, ys = np.nonzero(mask.T > 0)                                  # xx outer, yy inner
    Z = ref_depth[ys, xs].astype(f32) / f32(1000.0)
    X = Z * (xs.astype(f32) - tmpcx) * tmpfx
    Y = Z * (ys.astype(f32) - tmpcy) * tmpfy
    return np.stack([X, Y, Z]).astype(f32), np.stack([xs, ys]).astype(f32)


def to_se_3(w):
    """:1104-1114"""
    wx = np.zeros((3, 3), f32)
    wx[1, 2] = -w[0]; wx[0, 2] = w[1]; wx[0, 1] = -w[2]
    wx[2, 1] = w[0]; wx[2, 0] = -w[1]; wx[1, 0] = w[2]
    return wx


def compute_jacobian_of_now_frame(level, cR, cT, _3d, nowDist, dGx, dGy, fx, fy, cx, cy):
    """:306-414.  cR (3,3), cT (3,) float32.  Returns Jcbian (N, 6), reprojections (3, N), visible mask."""
    cR = cR.astype(f32); cT = cT.astype(f32)
    N = _3d.shape[1]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], f32)
    cTRep = np.repeat(cT.reshape(3, 1), N, axis=1)                   # igl::repmat(cT,1,_3d.cols(),cTRep)
    t3 = (cR.T @ (_3d - cTRep)).astype(f32)                          # _3d_transformed
    scaleFac = f32(2.0 ** (-level))
    scaleMatrix = np.eye(3, dtype=f32); scaleMatrix[0, 0] = scaleFac; scaleMatrix[1, 1] = scaleFac
    lastRow_inv = (f32(1.0) / t3[2]).astype(f32)
    for i in range(3):
        t3[i] = t3[i] * lastRow_inv
    rep = (scaleMatrix @ K @ t3).astype(f32)                         # _2d_reprojected
    nCols, nRows = nowDist.shape[1], nowDist.shape[0]
    J = np.zeros((N, 6), f32)
    with np.errstate(invalid="ignore"):
        skip = (rep[0] < 0) | (rep[0] > nCols) | (rep[1] < 0) | (rep[1] > nRows)      # :371
    vis = ~skip & np.isfinite(rep[0]) & np.isfinite(rep[1])
    # u == nCols / v == nRows passes the reference's test and then indexes out of range; treated as not visible (measure zero)
    vis &= (rep[0] < nCols) & (rep[1] < nRows)
    A2 = np.zeros((3, 6), f32)
    A2[:, 0:3] = -cR.T
    for i in np.nonzero(vis)[0]:
        xx = int(rep[0, i]); yy = int(rep[1, i])
        X, Y, Z = t3[0, i], t3[1, i], t3[2, i]
        G = np.array([dGx[yy, xx], dGy[yy, xx]], f32)
        A1 = np.zeros((2, 3), f32)
        A1[0, 0] = scaleFac * f32(fx) / Z
        A1[0, 2] = -scaleFac * f32(fx) * X / (Z * Z)
        A1[1, 1] = scaleFac * f32(fy) / Z
        A1[1, 2] = -scaleFac * f32(fy) * Y / (Z * Z)
        tmp = (cR.T @ t3[:, i]).astype(f32)
        A2[:, 3:6] = to_se_3(tmp)
        J[i] = (G @ A1 @ A2).astype(f32)
    return J, rep, vis


def get_weight_of(r):
    """:1047-1053: return 6.0 / (6.0 + r*r/.25) -- r*r in float, the rest in double, stored as float"""
    rr = (r.astype(f32) * r.astype(f32)).astype(np.float64)
    return (6.0 / (6.0 + rr / 0.25)).astype(f32)


def get_reprojected_epsilons(reprojections, nowDist, vis):
    """:425-462 (the shipped build: epsilon = _nowDist(floor(v), floor(u)))"""
    N = reprojections.shape[1]
    eps = np.zeros(N, f32); w = np.zeros(N, f32)
    idx = np.nonzero(vis)[0]
    eps[idx] = nowDist[np.floor(reprojections[1, idx]).astype(int), np.floor(reprojections[0, idx]).astype(int)]
    w[idx] = get_weight_of(eps[idx])
    ratio = f32(len(idx)) / f32(N)
    return eps, w, ratio


def weighted_subgradient(J, w, eps):
    """:714-720, :777: JTW = J' with column i scaled by w(i) (float); g = JTW.cast<double>() * epsilon.cast<double>()"""
    JTW = (J.T * w[None, :]).astype(f32)
    return JTW.astype(np.float64) @ eps.astype(np.float64)


# ------------------------------------------------------------------------------------------------ EPoseEstimator
def filter2d_forward(im, axis):
    """cv::filter2D(im_r, g, CV_64F, kern) with kernX = [0 -1 1] (centre row) / kernY its transpose, BORDER_REFLECT_101 (:331-339)"""
    a = im.astype(np.float64)
    if axis == 1:
        nxt = np.concatenate([a[:, 1:], a[:, -2:-1]], axis=1)        # column W maps to W-2
    else:
        nxt = np.concatenate([a[1:, :], a[-2:-1, :]], axis=0)
    return nxt - a


def evaluate3d(dim_r, fx, fy, cx, cy, scaleFactor):
    """:439-477.  X = Z/fx * (U - scaleFactor*cx) with U = ROW number, Y = Z/fy * (V - scaleFactor*cy) with V = COLUMN number; /1000."""
    Z = dim_r.astype(np.float64)
    rows, cols = Z.shape
    U = np.repeat(np.arange(rows, dtype=np.float64)[:, None], cols, axis=1)
    V = np.repeat(np.arange(cols, dtype=np.float64)[None, :], rows, axis=0)
    X = Z / fx * (U - scaleFactor * cx)
    Y = Z / fy * (V - scaleFactor * cy)
    return X / 1000.0, Y / 1000.0, Z / 1000.0


def evaluate_jacobian(im_r, dim_r, fx, fy, cx, cy, scaleFactor):
    """:320-430.  Returns J (rows*cols, 6): columns pJ1, pJ2, pJ3, pJ4, pJ4, pJ6 (:415), each the COLUMN-major flattening (:405-410)."""
    egx = filter2d_forward(im_r, 1)
    egy = filter2d_forward(im_r, 0)
    X, Y, Z = evaluate3d(dim_r, fx, fy, cx, cy, scaleFactor)
    Z_inv = 1.0 / Z
    Z2_inv = 1.0 / (Z * Z)
    J1 = fx * egx * Z_inv
    J2 = fy * egy * Z_inv
    J3 = -fy * egy * Y * Z2_inv - fx * egx * X * Z2_inv
    J4 = egy * (-fy * Y * Y * Z2_inv - fy) - fx * egx * X * Y * Z2_inv
    J5 = egx * (fx * X * X * Z2_inv + fx) + fy * egy * X * Y * Z2_inv
    J6 = fy * egy * X * Z_inv - fx * egy * Y * Z_inv
    flat = lambda a: a.flatten(order="F")                            # Eigen::Map over column-major storage
    J = np.stack([flat(J1), flat(J2), flat(J3), flat(J4), flat(J4), flat(J6)], axis=1)
    return J, (X, Y, Z), flat(J5)

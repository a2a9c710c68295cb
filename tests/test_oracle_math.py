"""Closed-form pieces of the oracle (SE(3), rotationize, quaternion, GOP, weights) and solver sanity."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle_lib as O


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def test_se3_exp_log_roundtrip_and_small_angles():
    rng = np.random.default_rng(0)
    for scale in (1e-12, 1e-7, 1e-3, 0.3, 2.0):
        for _ in range(20):
            psi = rng.normal(size=6) * scale
            R, t = O.se3_exp(psi)
            assert np.allclose(R @ R.T, np.eye(3), atol=1e-13) and abs(np.linalg.det(R) - 1) < 1e-13
            assert np.allclose(R, Rotation.from_rotvec(psi[3:]).as_matrix(), atol=1e-13)
            back = O.se3_log(R, t)
            if np.linalg.norm(psi[3:]) < 3.0:          # principal branch: log inverts exp
                assert np.allclose(back, psi, rtol=1e-9, atol=1e-13 + 1e-9 * scale)
            R2, t2 = O.se3_exp(back)                   # always: exp(log(.)) reproduces the group element
            assert np.allclose(R2, R, atol=1e-11) and np.allclose(t2, t, atol=1e-10)
    R, t = O.se3_exp(np.zeros(6))
    assert np.array_equal(R, np.eye(3)) and np.array_equal(t, np.zeros(3))
    assert np.array_equal(O.se3_log(np.eye(3), np.zeros(3)), np.zeros(6))


def test_se3_exp_translation_uses_V():
    psi = np.array([0.1, -0.2, 0.3, 0.4, 0.5, -0.6])
    R, t = O.se3_exp(psi)
    w = psi[3:]
    th = np.linalg.norm(w)
    Om = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    assert np.allclose(t, V @ psi[:3], atol=1e-14)


def test_rotationize_is_polar_factor():
    rng = np.random.default_rng(1)
    for _ in range(20):
        R = Rotation.from_rotvec(rng.normal(size=3)).as_matrix()
        A = R + rng.normal(size=(3, 3)) * 1e-3
        U, _, Vt = np.linalg.svd(A)
        assert np.allclose(O.rotationize(A), U @ Vt, atol=1e-12)
        assert np.allclose(O.rotationize(R), R, atol=1e-14)


def test_quaternion_matches_eigen_convention():
    rng = np.random.default_rng(2)
    for _ in range(50):
        R = Rotation.from_rotvec(rng.normal(size=3) * 2).as_matrix()
        q = O.rot_to_quat(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert rot_angle(Rotation.from_quat(q).as_matrix(), R) < 1e-7
    # trace <= 0 branch
    R = Rotation.from_rotvec([np.pi * 0.99, 0, 0]).as_matrix()
    assert rot_angle(Rotation.from_quat(O.rot_to_quat(R)).as_matrix(), R) < 1e-7


def test_weight_formula():
    # getWeightOf: 6 / (6 + r^2 / 0.25)  (src/SolveDVO.cpp:1051)
    for r in (0.0, 0.5, 1.0, 7.25, 255.0):
        assert O.lib().orc_weight_ref(r) == np.float32(6.0 / (6.0 + float(np.float32(r) * np.float32(r)) / 0.25))


def test_gop_composition():
    # src/GOP.cpp:138-196: global = key * relative; keyframes / updateMostRecentToKeyFrame re-anchor
    rng = np.random.default_rng(3)
    n = 12
    kind = np.array([1, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0], np.int32)
    rel = np.zeros((n, 12))
    Ts = []
    for i in range(n):
        R, t = O.se3_exp(rng.normal(size=6) * 0.1)
        rel[i, :9], rel[i, 9:] = R.reshape(9), t
        M = np.eye(4); M[:3, :3] = R; M[:3, 3] = t
        Ts.append(M)
    out, is_key, reason = O.gop_replay(kind, np.full(n, 5, np.int32), rel)
    key = np.eye(4)
    for i in range(n):
        Gm = key @ Ts[i]
        assert np.allclose(out[i, :9].reshape(3, 3), Gm[:3, :3], atol=1e-13) and np.allclose(out[i, 9:12], Gm[:3, 3], atol=1e-13)
        assert np.allclose(out[i, 12:15], Gm[:3, 3], atol=1e-13)
        assert rot_angle(Rotation.from_quat(out[i, 15:19]).as_matrix(), Gm[:3, :3]) < 1e-7
        assert bool(is_key[i]) == (kind[i] != 0) and reason[i] == (5 if kind[i] else -1)
        if kind[i]:
            key = Gm


def test_solver_reduces_energy_and_approaches_truth():
    d = O.synth_pair(0)
    r = O.align_pair(d["ref_gray"], d["ref_depth"], d["now_gray"], trace=True)
    assert r["status"] == 0 and (r["npts"] > 0).all()
    e0 = r["trace"]["energy"][0]          # level 0 energies
    assert r["best_energy"][0] <= e0[0]
    # the estimate is closer to the true motion than the identity initial guess
    err0 = np.linalg.norm(d["T"]) + rot_angle(np.eye(3), d["R"])
    err1 = np.linalg.norm(r["T"] - d["T"]) + rot_angle(r["R"], d["R"])
    assert err1 < err0
    # trust region: no executed step moves the pose by more than 0.003 (+ rounding)
    T = r["trace"]["T"][0][: r["iterations_run"][0]]
    assert (np.linalg.norm(np.diff(T, axis=0), axis=1) <= 0.003 + 1e-9).all()


@pytest.mark.parametrize("solver", [O.GN, O.LM])
def test_gn_lm_extension_runs(solver):
    d = O.synth_pair(2)
    r = O.align_pair(d["ref_gray"], d["ref_depth"], d["now_gray"], iters=(10, 10, 10, 10), scfg=O.cfg(solver, O.JAC_EXACT))
    assert r["status"] == 0
    assert np.linalg.norm(r["T"] - d["T"]) + rot_angle(r["R"], d["R"]) < np.linalg.norm(d["T"]) + rot_angle(np.eye(3), d["R"])


def test_normal_equations_linearity_and_symmetry():
    d = O.synth_pair(4, 320, 240, (262.5, 262.5, 159.5, 119.5))
    ref = O.preprocess_level(d["ref_gray"], d["ref_depth"], 0)
    now = O.preprocess_level(d["now_gray"], None, 0)
    X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], 0, (262.5, 262.5, 159.5, 119.5))
    K = (262.5, 262.5, 159.5, 119.5)
    full = O.evaluate(X, Y, Z, now["dtn"], now["gx"], now["gy"], 0, np.eye(3), np.zeros(3), K, per_point=True)
    h = len(X) // 2
    a = O.evaluate(X[:h], Y[:h], Z[:h], now["dtn"], now["gx"], now["gy"], 0, np.eye(3), np.zeros(3), K)
    b = O.evaluate(X[h:], Y[h:], Z[h:], now["dtn"], now["gx"], now["gy"], 0, np.eye(3), np.zeros(3), K)
    assert np.allclose(a["H"] + b["H"], full["H"], rtol=1e-12) and np.allclose(a["g"] + b["g"], full["g"], rtol=1e-12)
    assert a["nvis"] + b["nvis"] == full["nvis"]
    # H = J^T W J and g = J^T W eps from the per-point outputs
    J, w, e = full["J"].astype(np.float64), full["w"], full["eps"].astype(np.float64)
    Jw = (full["J"] * w[:, None]).astype(np.float32).astype(np.float64)
    assert np.allclose(Jw.T @ J, full["H"], rtol=1e-9) and np.allclose(Jw.T @ e, full["g"], rtol=1e-9)
    assert np.allclose(full["H"], full["H"].T, rtol=1e-6)


def test_interpolated_dt_lookup_properties():
    """SolveDVO::interpolate restatement (src/SolveDVO.cpp:1285-1308): exact at integer positions, the stated
    squared-bilinear formula in fp32 elsewhere, ceil indices clamped at the border."""
    rng = np.random.default_rng(5)
    H, W = 12, 16
    F = (rng.random((H, W)) * 255).astype(np.float32)
    X = np.array([0.0], np.float32); Y = np.array([0.0], np.float32); Z = np.array([1.0], np.float32)
    zeros = np.zeros((H, W), np.float32)

    def lookup(u, v):      # a camera with fx = fy = 1 and principal point (u, v) sends the point (0,0,1) to pixel (u, v)
        o = O.evaluate(X, Y, Z, F, zeros, zeros, 0, np.eye(3), np.zeros(3), K=(1.0, 1.0, u, v), per_point=True, residual=O.RES_DT_INTERP)
        assert o["nvis"] == 1 and o["u"][0] == np.float32(u) and o["v"][0] == np.float32(v)
        return o["eps"][0]

    for (u, v) in ((3.0, 4.0), (0.0, 0.0), (15.0, 11.0)):
        assert lookup(u, v) == F[int(v), int(u)]
    f32 = np.float32
    u, v = f32(5.25), f32(7.5)
    ix, iy = f32(0.25), f32(0.5)
    f1 = np.sqrt(((f32(1) - ix) * F[7, 5]) * F[7, 5] + (ix * F[7, 6]) * F[7, 6], dtype=np.float32)
    f2 = np.sqrt(((f32(1) - ix) * F[8, 5]) * F[8, 5] + (ix * F[8, 6]) * F[8, 6], dtype=np.float32)
    want = np.sqrt(((f32(1) - iy) * f1) * f1 + (iy * f2) * f2, dtype=np.float32)
    assert lookup(float(u), float(v)) == want
    # within one pixel of the right / bottom border the ceil index is clamped: the lookup degenerates to the last column / row
    assert lookup(15.5, 3.0) == F[3, 15]
    assert lookup(2.0, 11.75) == F[11, 2]


def test_rgbd_exponential_map_and_qr_solve():
    """RGBDOdometry::exponentialMap (src/RGBDOdometry.cpp:707-745) against scipy's matrix exponential of the twist, and the
    pivoted Householder QR solve (:563) against numpy on well- and ill-scaled SPD systems."""
    from scipy.linalg import expm
    rng = np.random.default_rng(11)
    for _ in range(20):
        psi = np.concatenate([rng.standard_normal(3), rng.standard_normal(3) * 0.3])
        w = psi[3:]
        tw = np.zeros((4, 4)); tw[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]; tw[:3, 3] = psi[:3]
        assert np.allclose(O.rgbd_exponential_map(psi), expm(tw), atol=1e-12)
    assert np.array_equal(O.rgbd_exponential_map(np.array([1.0, 2.0, 3.0, 0, 0, 1e-13])), np.eye(4))     # theta < 1e-12 -> identity, translation dropped (:718-722)
    for scale in (1.0, 1e6):
        J = rng.standard_normal((200, 6)) * np.array([scale, 1, 1, 1e-3, 1, 1])
        A = J.T @ J; b = rng.standard_normal(6)
        x = O.rgbd_qr_solve6(A, b)
        assert np.allclose(A @ x, b, rtol=1e-8, atol=1e-8 * np.abs(b).max())
        assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-6)


def test_exact_jacobian_is_the_derivative_of_the_reprojection():
    """JAC_EXACT (the oracle's analytic Jacobian wrt the right-multiplicative update cT += cR*ups, cR <- cR*exp(om)) against
    central finite differences of the oracle's own reprojection: with gradient images (gx, gy) = (1, 0) / (0, 1) the
    Jacobian row is du/dpsi / dv/dpsi.  Ties the Jacobian convention to the pose update the solver actually applies."""
    rng = np.random.default_rng(21)
    H, W, level = 60, 80, 1
    N = 40
    Z = rng.uniform(1.0, 4.0, N).astype(np.float32)
    X = (rng.uniform(-0.3, 0.3, N) * Z).astype(np.float32); Y = (rng.uniform(-0.2, 0.2, N) * Z).astype(np.float32)
    K = (120.0, 118.0, 79.5, 59.5)             # level-0 intrinsics; level 1 halves them
    R0 = O.se3_exp(np.array([0, 0, 0, 0.02, -0.015, 0.01]))[0]; T0 = np.array([0.01, -0.02, 0.015])
    ones, zeros = np.ones((H, W), np.float32), np.zeros((H, W), np.float32)

    def reproj(R, T):
        o = O.evaluate(X, Y, Z, zeros, zeros, zeros, level, R, T, K=K, jac=O.JAC_EXACT, weight=O.W_NONE, per_point=True)
        return o["u"].astype(np.float64), o["v"].astype(np.float64)

    Ju = O.evaluate(X, Y, Z, zeros, ones, zeros, level, R0, T0, K=K, jac=O.JAC_EXACT, weight=O.W_NONE, per_point=True)
    Jv = O.evaluate(X, Y, Z, zeros, zeros, ones, level, R0, T0, K=K, jac=O.JAC_EXACT, weight=O.W_NONE, per_point=True)
    assert Ju["nvis"] == N
    h = 1e-3
    for k in range(6):
        d = np.zeros(6); d[k] = h
        Rp, tp = O.se3_exp(d); Rm, tm = O.se3_exp(-d)
        up, vp = reproj(R0 @ Rp, T0 + R0 @ tp)
        um, vm = reproj(R0 @ Rm, T0 + R0 @ tm)
        fu, fv = (up - um) / (2 * h), (vp - vm) / (2 * h)
        assert np.allclose(Ju["J"][:, k], fu, rtol=2e-2, atol=2e-2), (k, np.abs(Ju["J"][:, k] - fu).max())
        assert np.allclose(Jv["J"][:, k], fv, rtol=2e-2, atol=2e-2), (k, np.abs(Jv["J"][:, k] - fv).max())

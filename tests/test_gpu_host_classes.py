"""The C++ host classes with the reference's names (SolveDVO, EPoseEstimator, PyramidalStorageStruct) driven through their
C shim: same call sequences a C++ user of the reference makes, results against the oracle."""
import numpy as np
import pytest

import host_lib as Hh
import oracle_lib as O

pytestmark = pytest.mark.gpu

W, H, L = 320, 240, 4
K = (262.5, 262.5, 159.5, 119.5)


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def test_solvedvo_run_iterations_out_parameters():
    d = O.synth_pair(21, W, H, K)
    for level, iters in ((3, 20), (1, 12)):
        ref = O.preprocess_level(d["ref_gray"], d["ref_depth"], level)
        now = O.preprocess_level(d["now_gray"], None, level)
        X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], level, K)
        o = O.run_iterations(X, Y, Z, now["dtn"], now["gx"], now["gy"], level, iters, K=K)
        g = Hh.solvedvo_run_iterations(d["ref_gray"], d["ref_depth"], d["now_gray"], d["now_depth"], L, level, iters, K)
        assert rot_angle(g["R"], o["R"]) < 1e-5 and np.linalg.norm(g["T"] - o["T"]) < 1e-5
        assert g["best_index"] == o["best_index"] and abs(g["visible_ratio"] - o["visible_ratio"]) < 1e-6
        assert np.allclose(g["energies"], o["energies"], rtol=1e-6)
        assert len(g["eps"]) == len(X)
        # residuals / reprojections of the best iterate (evaluated at the returned pose)
        assert np.allclose(g["eps"], o["best_eps"], atol=1e-4) and np.allclose(g["u"], o["best_u"], atol=1e-3)
        assert abs(g["b_cap"] - o["best_eps"].mean()) < 1e-3


def test_solvedvo_loop_matches_sequence_schedule():
    from test_gpu_sequence import oracle_sequence
    nframes = 8
    gray, depth, Rw, Tw = O.synth_sequence(91, nframes, W, H, K, max_angle_deg=0.4, max_trans_m=0.008)
    iters = (6, 6, 6, 6)
    out, is_key, reason = Hh.solvedvo_sequence(gray, depth, L, iters, K)
    orel, okind, oglob = oracle_sequence(gray, depth, L, iters, K)
    assert list(is_key) == [1 if k else 0 for k in okind]
    assert reason[0] == 1 and reason[4] == 5
    for t in range(nframes):
        assert rot_angle(out[t, :9].reshape(3, 3), oglob[t, :9].reshape(3, 3)) < 1e-5
        assert np.linalg.norm(out[t, 9:12] - oglob[t, 9:12]) < 1e-5
        assert np.allclose(out[t, 12:15], out[t, 9:12])


@pytest.mark.parametrize("compat", [True, False])
def test_eposeestimator_class(compat):
    d = O.synth_pair(33, W, H, K, bgr=True)
    level = 2
    g = Hh.eposeestimator(d["ref_bgr"], d["ref_depth"], d["now_bgr"], d["now_depth"], K, compat, level, iters=3 if compat else 6,
                          huber_k=0.0 if compat else 10.0, lambda0=0.0 if compat else 1e-3)
    assert g["levels"] == 5                                        # setRefFrame always builds levels 0..4
    o = O.photo_build_ref_level(d["ref_bgr"], d["ref_depth"], level, K, compat=compat)
    assert np.array_equal(g["gray"], o["gray"]) and np.array_equal(g["X"], o["X"]) and np.array_equal(g["J"], o["J"])
    assert np.allclose(g["A"], o["A"], rtol=1e-10, atol=1e-10 * np.abs(o["A"]).max())
    if compat:
        assert g["status"] == 2 and np.allclose(g["R"], np.eye(3)) and np.allclose(g["T"], 0)
        assert g["visible"] in (0.0, 1.0)                          # integer division of the reference (:171)
    else:
        oe = O.photo_estimate(O.photo_now_level(d["now_bgr"], level), np.eye(3), np.zeros(3), 6, K, compat=False, huber_k=10.0, lambda0=1e-3)
        assert g["status"] == oe["status"]
        assert rot_angle(g["R"], oe["R"]) < 1e-7 and np.linalg.norm(g["T"] - oe["T"]) < 1e-9


def test_rgbdodometry_class_matches_oracle():
    """RGBDOdometry (host class over dvo_rgbd_*): eventLoop's schedule over a short sequence and the materialised
    computeJacobian / computeEpsilon outputs against the oracle's restatement of src/RGBDOdometry.cpp."""
    Wd, Hd, Kd = 320, 240, (262.5, 262.5, 159.5, 119.5)
    gray, depth, _, _ = O.synth_sequence(77, 4, Wd, Hd, Kd)
    bgr = np.repeat(gray[..., None], 3, axis=3)                      # BGR2GRAY(g, g, g) == g exactly
    out, npts, eps = Hh.rgbdodometry_sequence(bgr, depth, Kd)
    T = np.eye(4)
    for t in range(4):
        T, info = O.rgbd_gauss_newton(bgr[0], depth[0], bgr[t], depth[t], Kd, levels=(3, 2), T0=T)
        assert np.allclose(out[t], T, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(T).max())), f"frame {t}: {np.abs(out[t] - T).max()}"
        assert npts[t] == int(info[1, 0]) and abs(eps[t] - info[1, 5]) <= 1e-9 * max(1.0, info[1, 5])
    assert np.abs(out[3] - np.eye(4)).max() > 1e-6
    J, marks, e, roi = Hh.rgbdodometry_materialise(bgr[0], depth[0], bgr[2], depth[2], Kd, 2, out[2])
    oj = O.rgbd_jacobian(bgr[0], depth[0], 2, Kd)
    oe = O.rgbd_epsilon(bgr[0], depth[0], bgr[2], depth[2], 2, Kd, out[2])
    assert np.array_equal(J, oj["J"]) and np.array_equal(e, oe["eps"])
    want = np.zeros_like(marks); want[oj["ij"][:, 0], oj["ij"][:, 1]] = 1
    assert np.array_equal(marks, want)
    seen = oe["uv"][oe["uv"][:, 0] >= 0]
    wroi = np.zeros_like(roi); wroi[seen[:, 0], seen[:, 1]] = 1
    assert np.array_equal(roi, wroi)

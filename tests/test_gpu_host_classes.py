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


def _dump_sequence(tmp_path, gray, depth, levels, start):
    """framemono_%04d.xml dumps as the publisher writes them (src/camTopic2PublisherPyD.cpp:315-365): all pyramid levels."""
    for t in range(gray.shape[0]):
        monos = [O.pyr_nearest(gray[t], l) for l in range(levels)]
        deps = [O.pyr_nearest(depth[t], l) for l in range(levels)]
        assert Hh.store_frame_xml(tmp_path / f"framemono_{start + t:04d}.xml", monos, deps) == 0


def test_solvedvo_public_loops_over_xml_dumps_and_callbacks(tmp_path):
    """The reference's public surface (include/SolveDVO.h:154-168): loop() and loopDry() polling a frame source -- the
    __DATA_FROM_XML_FILES__ replay of src/SolveDVO.cpp:1826-1840 / :1954-1968 and an in-memory callback with empty turns --,
    casualTestFunction() (:2377-2442) and loopFromFile() (:2448-2600)."""
    from test_gpu_sequence import oracle_sequence
    Wd, Hd, Ld, Kd = 160, 120, 3, (131.25, 131.25, 79.5, 59.5)
    nframes, start = 7, 10
    gray, depth, _, _ = O.synth_sequence(93, nframes, Wd, Hd, Kd, max_angle_deg=0.4, max_trans_m=0.008)
    _dump_sequence(tmp_path, gray, depth, Ld, start)
    iters = (8, 8, 8)
    want, wkey, wreason = Hh.solvedvo_sequence(gray, depth, Ld, iters, Kd)
    # loop(): the source stops at the first missing file ("No More files, Quitting..")
    n, out, is_key, reason = Hh.solvedvo_loop_xml(tmp_path, start, 10_000, Wd, Hd, Ld, iters, Kd)
    assert n == nframes and np.array_equal(out, want) and np.array_equal(is_key, wkey) and np.array_equal(reason, wreason)
    # ... or at __DATA_FROM_XML_FILES__END
    n, out, _, _ = Hh.solvedvo_loop_xml(tmp_path, start, start + 3, Wd, Hd, Ld, iters, Kd)
    assert n == 4 and np.array_equal(out, want[:4])
    # loop() with turns on which no frame arrived: `continue`, same result
    m, out_cb = Hh.solvedvo_loop_callback(gray, depth, Ld, iters, Kd, skip_every=3)
    assert m == nframes and np.array_equal(out_cb[:, :12], want[:, :12])
    # loopDry(): every frame consumed as a now frame, nothing pushed to the pose graph
    nd, _, _, _ = Hh.solvedvo_loop_xml(tmp_path, start, 10_000, Wd, Hd, Ld, iters, Kd, dry=True)
    assert nd == nframes                                # dryFrames; the shim adds 100000 if anything reached the pose graph
    # casualTestFunction(): ref = file `start`, now = file `start + 3`, one runIterations(0, 40) from identity
    ne, R, T, en = Hh.solvedvo_casual(tmp_path, start, start + 3, 40, Wd, Hd, Ld, Kd)
    ref = O.preprocess_level(gray[0], depth[0], 0); now = O.preprocess_level(gray[3], None, 0)
    X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], 0, Kd)
    o = O.run_iterations(X, Y, Z, now["dtn"], now["gx"], now["gy"], 0, 40, K=Kd)
    assert ne == 40 and np.allclose(en[: o["iterations_run"]], o["energies"][: o["iterations_run"]], rtol=1e-6)
    assert rot_angle(R, o["R"]) < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5
    # a missing dump: logs "Cannot open file1" and returns without touching the pose
    ne, R, T, _ = Hh.solvedvo_casual(tmp_path, 999, start, 5, Wd, Hd, Ld, Kd)
    assert ne == 0 and np.array_equal(R, np.eye(3))
    # loopFromFile(): the reference's rule is index % 5 == 0 -> files 10 and 15 become references (15 after its own alignment)
    k, poses = Hh.solvedvo_loop_from_file(tmp_path, start, start + nframes, 30, Wd, Hd, Ld, Kd)
    assert k == nframes
    # oracle-driven replay of the same schedule
    keyR, keyT, nR, nT = np.eye(3), np.zeros(3), np.eye(3), np.zeros(3)
    cR, cT = np.eye(3), np.zeros(3)
    refi = None
    def run(ri, ni, R0, T0):
        r = O.preprocess_level(gray[ri], depth[ri], 0); nw = O.preprocess_level(gray[ni], None, 0)
        Xr, Yr, Zr, _, _ = O.select_points(r["edge"], r["depth"], 0, Kd)
        oo = O.run_iterations(Xr, Yr, Zr, nw["dtn"], nw["gx"], nw["gy"], 0, 30, R0=R0, T0=T0, K=Kd)
        return oo["R"], oo["T"]
    for t in range(nframes):
        idx = start + t
        if idx % 5 == 0:
            if idx > start:
                cR, cT = run(refi, t, cR, cT)
            keyR, keyT = nR, nT
            refi = t
            cR, cT = np.eye(3), np.zeros(3)
        cR, cT = run(refi, t, cR, cT)
        nT = keyT + keyR @ cT; nR = keyR @ cR
        assert rot_angle(poses[t, :9].reshape(3, 3), nR) < 1e-5 and np.linalg.norm(poses[t, 9:] - nT) < 1e-5, t


def test_pyramidal_storage_add_level_reference_signature():
    """PyramidalStorageStruct::addLevel(level, im_r_color, im_r, dim_r, X, Y, Z, J, grayVals, redVals, greenVals, blueVals)
    (include/PyramidalStorage.h:42-48, src/PyramidalStorage.cpp:38-65): pure push_back storage with `level` ignored; pushed into
    an estimator's storage, the caller's images are what estimate() aligns against."""
    dA = O.synth_pair(33, W, H, K, bgr=True)
    dB = O.synth_pair(34, W, H, K, bgr=True)
    level = 1
    bad, R, T, A, sizes = Hh.pydstore_add_level(dA["ref_bgr"], dA["ref_depth"], dB["ref_bgr"], dB["ref_depth"], dB["now_bgr"], dB["now_depth"], K, level)
    assert bad == 0 and list(sizes) == [2, 0, 5]
    want = Hh.eposeestimator(dB["ref_bgr"], dB["ref_depth"], dB["now_bgr"], dB["now_depth"], K, False, level, iters=6, huber_k=10.0, lambda0=1e-3)
    assert np.array_equal(A, want["A"]) and np.array_equal(R, want["R"]) and np.array_equal(T, want["T"])
    other = Hh.eposeestimator(dA["ref_bgr"], dA["ref_depth"], dB["now_bgr"], dB["now_depth"], K, False, level, iters=6, huber_k=10.0, lambda0=1e-3)
    assert not np.array_equal(A, other["A"])


def test_cv_mat_and_eigen_overloads_match_the_pod_api():
    """The reference's exact signatures -- setRefFrame(cv::Mat&, cv::Mat&), estimate(Eigen::Matrix3d&, Eigen::Vector3d&),
    runIterations(..., Eigen::Matrix3d&, Eigen::Vector3d&, Eigen::VectorXf&, ...), PyramidalStorageStruct::addLevel / getLevel with
    cv::Mat& and (column-major) Eigen:: arguments -- exist under `#ifdef DVO_HAVE_OPENCV / DVO_HAVE_EIGEN`.  Neither library is
    installed here, so the shim is compiled against minimal stand-in headers (tests/standin_include) and must give exactly what the
    POD API gives."""
    d = O.synth_pair(33, W, H, K, bgr=True)
    level = 2
    want = Hh.eposeestimator(d["ref_bgr"], d["ref_depth"], d["now_bgr"], d["now_depth"], K, False, level, iters=6, huber_k=10.0, lambda0=1e-3)
    got = Hh.cv_eposeestimator(d["ref_bgr"], d["ref_depth"], d["now_bgr"], d["now_depth"], K, level)
    assert got["levels"] == 5 and got["roundtrip_ok"]
    assert np.array_equal(got["R"], want["R"]) and np.array_equal(got["T"], want["T"]) and np.array_equal(got["J"], want["J"])
    w2 = Hh.solvedvo_run_iterations(d["ref_gray"], d["ref_depth"], d["now_gray"], d["now_depth"], L, 1, 12, K)
    g2 = Hh.cv_solvedvo_run_iterations(d["ref_gray"], d["ref_depth"], d["now_gray"], d["now_depth"], L, 1, 12, K)
    assert np.array_equal(g2["R"], w2["R"]) and np.array_equal(g2["T"], w2["T"]) and np.array_equal(g2["energies"], w2["energies"])
    assert np.array_equal(g2["eps"], w2["eps"]) and np.array_equal(g2["u"], w2["u"]) and g2["best_index"] == w2["best_index"]
    assert np.array_equal(g2["gop"][:9].reshape(3, 3), w2["R"]) and np.array_equal(g2["gop"][9:], w2["T"])   # first key frame = identity

"""The C++ oracle against an INDEPENDENT numpy transliteration of the reference source (tests/ref_transliteration.py, written from
/root/reference/src/SolveDVO.cpp:224-264, 306-462 and /root/reference/src/EPoseEstimator.cpp:320-477 rather than from the oracle):
catches transcription slips -- swapped indices, signs, operand order -- in the part of the oracle no OpenCV fixture can pin."""
import numpy as np
import pytest

import oracle_lib as O
import ref_transliteration as T

K = O.K640


def _levels(seed, level):
    d = O.synth_pair(seed)
    ref = O.preprocess_level(d["ref_gray"], d["ref_depth"], level)
    now = O.preprocess_level(d["now_gray"], None, level)
    return d, ref, now


@pytest.mark.parametrize("level", [0, 2, 3])
def test_enlist_ref_edge_points_matches_transliteration(level):
    d, ref, now = _levels(7, level)
    mask = ((ref["edge"] > 0) & (ref["depth"].astype(np.float32) > 100.0)).astype(np.int32)      # selectedPts (:1230-1264)
    p3, p2 = T.enlist_ref_edge_pts(level, mask, ref["depth"].astype(np.float32), *K)
    X, Y, Z, u, v = O.select_points(ref["edge"], ref["depth"], level)
    assert p3.shape[1] == len(X) > 0
    assert np.array_equal(p2[0], u) and np.array_equal(p2[1], v)                                 # same enumeration order: column outer, row inner
    assert np.array_equal(p3[0].view(np.uint32), X.view(np.uint32)) and np.array_equal(p3[1].view(np.uint32), Y.view(np.uint32))
    assert np.array_equal(p3[2].view(np.uint32), Z.view(np.uint32))


@pytest.mark.parametrize("level,seed", [(0, 7), (1, 8), (3, 9)])
def test_jacobian_residual_weight_match_transliteration(level, seed):
    d, ref, now = _levels(seed, level)
    X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], level)
    p3 = np.stack([X, Y, Z]).astype(np.float32)
    # at the identity pose every point reprojects onto integer pixel coordinates, so float noise (BLAS vs the oracle's operation order)
    # flips floor() for a few per cent of them; away from it the two evaluations pick the same texel everywhere
    for (R, t, agree) in ((np.eye(3), np.zeros(3), 0.95), (d["R"], d["T"], 0.995)):
        J, rep, vis = T.compute_jacobian_of_now_frame(level, R.astype(np.float32), t.astype(np.float32), p3, now["dtn"], now["gx"], now["gy"], *K)
        eps, w, ratio = T.get_reprojected_epsilons(rep, now["dtn"], vis)
        o = O.evaluate(X, Y, Z, now["dtn"], now["gx"], now["gy"], level, R, t, per_point=True)
        assert np.allclose(rep[0], o["u"], rtol=2e-6, atol=2e-4) and np.allclose(rep[1], o["v"], rtol=2e-6, atol=2e-4)
        # a reprojection within float noise of a pixel boundary may land in the neighbouring texel in one of the two evaluations
        same = (np.floor(rep[0]) == np.floor(o["u"])) & (np.floor(rep[1]) == np.floor(o["v"])) & vis & (o["w"] > 0)
        assert same.sum() >= agree * vis.sum() and abs(int(vis.sum()) - o["nvis"]) <= 2
        assert np.array_equal(eps[same].view(np.uint32), o["eps"][same].view(np.uint32))
        assert np.array_equal(w[same].view(np.uint32), o["w"][same].view(np.uint32))
        scale = np.abs(o["J"][same]).max(axis=1, keepdims=True) + 1e-6
        assert (np.abs(J[same] - o["J"][same]) <= 2e-5 * scale).all()
        # columns: 0-2 translation, 3-5 rotation, and not identically zero
        assert np.abs(o["J"][same][:, :3]).max() > 0 and np.abs(o["J"][same][:, 3:]).max() > 0
        g = T.weighted_subgradient(J, w, eps)
        keep = same | ~vis
        if keep.all():
            assert np.allclose(g, o["g"], rtol=1e-4, atol=1e-4 * np.abs(o["g"]).max())
        else:   # compare on the common set
            g_common = T.weighted_subgradient(J[same], w[same], eps[same])
            og = (o["J"][same] * o["w"][same][:, None]).astype(np.float32).astype(np.float64).T @ o["eps"][same].astype(np.float64)
            assert np.allclose(g_common, og, rtol=1e-4, atol=1e-4 * np.abs(og).max())
        assert abs(float(ratio) - o["nvis"] / len(X)) < 1e-3


@pytest.mark.parametrize("level", [0, 2, 4])
def test_eposeestimator_jacobian_and_3d_match_transliteration(level):
    d = O.synth_pair(11, 320, 240, (262.5, 262.5, 159.5, 119.5), bgr=True)
    Kp = (262.5, 262.5, 159.5, 119.5)
    o = O.photo_build_ref_level(d["ref_bgr"], d["ref_depth"], level, Kp, compat=True)
    J, (X, Y, Z), pJ5 = T.evaluate_jacobian(o["gray"], o["depth"], *Kp, scaleFactor=2.0 ** (-level))
    assert np.array_equal(X, o["X"]) and np.array_equal(Y, o["Y"]) and np.array_equal(Z, o["Z"])
    assert np.array_equal(T.filter2d_forward(o["gray"], 1), o["gx"]) and np.array_equal(T.filter2d_forward(o["gray"], 0), o["gy"])
    assert J.shape == o["J"].shape
    assert np.allclose(J, o["J"], rtol=1e-14, atol=0) and np.array_equal(J[:, 3], J[:, 4])       # quirk: column 5 is a copy of column 4 (:415)
    oc = O.photo_build_ref_level(d["ref_bgr"], d["ref_depth"], level, Kp, compat=False)
    assert not np.array_equal(oc["J"][:, 4], oc["J"][:, 3])                                      # the corrected mode really differs
    try:
        import cv2
    except Exception:
        return
    kx = np.array([[0, 0, 0], [0, -1.0, 1.0], [0, 0, 0]], np.float32)
    assert np.array_equal(cv2.filter2D(o["gray"], cv2.CV_64F, kx), T.filter2d_forward(o["gray"], 1))
    assert np.array_equal(cv2.filter2D(o["gray"], cv2.CV_64F, kx.T), T.filter2d_forward(o["gray"], 0))

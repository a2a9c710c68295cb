"""GPU parity tests of the RGBDOdometry path (semi-dense photometric Gauss-Newton, src/RGBDOdometry.cpp)
against the CPU oracle, through the C-ABI.

Bit-exact: BGR2GRAY + INTER_NEAREST pyramids, the semi-dense selection (set and enumeration order), Jacobian rows
(fp64, same operation order; rgbd.cu is built with -fmad=false), residuals and the pixels they were read from.
Tolerance: fp64 reductions A = J^T J, ||eps|| (summation order differs): relative 1e-10; poses after the full
coarse-to-fine run: 1e-9 relative (device sin / cos / sqrt differ from libm in the last ulp).
"""
import numpy as np
import pytest

import oracle_lib as O
import rgbd_odometry_b200 as dvo

pytestmark = pytest.mark.gpu

W, H, L = 320, 240, 4
K = (262.5, 262.5, 159.5, 119.5)
NP = 3


@pytest.fixture(scope="module")
def data():
    return O.synth_batch(700, NP, W, H, K, bgr=True, now_depth=True)


@pytest.fixture(scope="module")
def rg(data):
    r = dvo.RGBDAligner(W, H, L, max_batch=NP + 1, intrinsics=K)
    r.set_frames(dvo.FRAME_REF, data["ref_bgr"], data["ref_depth"], first=1)
    r.set_frames(dvo.FRAME_NOW, data["now_bgr"], data["now_depth"], first=1)
    r.compute_jacobians(NP, first=1)
    yield r
    r.close()


def _pose(seed):
    rng = np.random.default_rng(seed)
    T = O.rgbd_exponential_map(np.concatenate([rng.standard_normal(3) * 3.0, rng.standard_normal(3) * 2e-3]))   # translation in the depth's raw units (mm)
    return T


def test_pyramids_bit_exact(rg, data):
    for i in range(NP):
        for l in range(L):
            for f, (bgr, dep) in enumerate(((data["ref_bgr"][i], data["ref_depth"][i]), (data["now_bgr"][i], data["now_depth"][i]))):
                g, d = rg.get_level(i + 1, f, l)
                og, od = O.rgbd_level(bgr, dep, l)
                assert np.array_equal(g, og) and np.array_equal(d, od), f"pair {i} frame {f} L{l}"


def test_jacobian_selection_and_rows_bit_exact(rg, data):
    for i in range(NP):
        for l in range(1, L):
            o = O.rgbd_jacobian(data["ref_bgr"][i], data["ref_depth"][i], l, K)
            g = rg.eval(i + 1, l, np.eye(4))
            assert len(o["ij"]) > 100
            assert np.array_equal(g["ij"], o["ij"]), f"selection pair {i} L{l}"
            assert np.array_equal(g["J"], o["J"]), f"J pair {i} L{l}: max diff {np.abs(g['J'] - o['J']).max()}"
            A, n = rg.get_A(i + 1, l)
            assert n == len(o["ij"])
            assert np.allclose(A, o["A"], rtol=1e-10, atol=1e-10 * np.abs(o["A"]).max()), f"A pair {i} L{l}"
            assert np.array_equal(o["J"][:, 0], K[0] * K[0] * (1.0 / data_depth(rg, i, l, o["ij"])))      # quirk: column 0 = fx*fx*invZ (:487)


def data_depth(rg, i, l, ij):
    _, d = rg.get_level(i + 1, 0, l)
    return d[ij[:, 0], ij[:, 1]].astype(np.float64)


def test_epsilon_bit_exact(rg, data):
    for i in range(NP):
        for l in range(1, L):
            for T in (np.eye(4), _pose(10 * i + l)):
                o = O.rgbd_epsilon(data["ref_bgr"][i], data["ref_depth"][i], data["now_bgr"][i], data["now_depth"][i], l, K, T)
                g = rg.eval(i + 1, l, T)
                assert g["nvis"] == o["nvis"] and o["nvis"] > 50
                assert np.array_equal(g["uv"], o["uv"]), f"uv pair {i} L{l}"
                assert np.array_equal(g["eps"], o["eps"]), f"eps pair {i} L{l}"
                assert np.allclose(g["b"], o["b"], rtol=1e-10, atol=1e-10 * np.abs(o["b"]).max())
                assert abs(g["sumsq"] - o["sumsq"]) <= 1e-10 * max(1.0, o["sumsq"])


def test_gauss_newton_matches_oracle(rg, data):
    """eventLoop's schedule (:157-158): gaussNewtonIterations(3, T) then (2, T), T carried, 3 iterations, ||eps|| < 200 exit."""
    rg.set_pose(NP, None, first=1)
    for l in (3, 2):
        rg.gauss_newton(NP, l, first=1)
    T, info = rg.get_poses(NP, first=1)
    moved = 0
    for i in range(NP):
        oT, oi = O.rgbd_gauss_newton(data["ref_bgr"][i], data["ref_depth"][i], data["now_bgr"][i], data["now_depth"][i], K, levels=(3, 2))
        assert np.allclose(T[i], oT, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(oT).max())), f"pair {i}: {np.abs(T[i] - oT).max()}"
        for k, l in enumerate((3, 2)):
            assert info[i].npts[l] == int(oi[k, 0]) and info[i].iters_run[l] == int(oi[k, 1]) and info[i].updates[l] == int(oi[k, 2])
            assert info[i].nvis_last[l] == int(oi[k, 3])
            assert abs(info[i].eps_norm_first[l] - oi[k, 4]) <= 1e-10 * oi[k, 4]
            assert abs(info[i].eps_norm_last[l] - oi[k, 5]) <= 1e-9 * oi[k, 5]
        assert info[i].status == 0
        moved += int(np.abs(T[i] - np.eye(4)).max() > 1e-6)
    assert moved == NP            # every pair really took Gauss-Newton steps


def test_early_exit_and_status_bits(rg, data):
    # identical frames: ||eps|| < 200 at the first evaluation, no update (:556).  (eps is not identically 0: X*fx/Z + cx can
    # round just below the integer it came from and floor() then reads the neighbouring pixel -- in the oracle as well.)
    r = dvo.RGBDAligner(W, H, L, max_batch=1, intrinsics=K)
    r.set_frames(dvo.FRAME_REF, data["ref_bgr"][:1], data["ref_depth"][:1])
    r.set_frames(dvo.FRAME_NOW, data["ref_bgr"][:1], data["ref_depth"][:1])
    r.compute_jacobians(1)
    r.set_pose(1)
    r.gauss_newton(1, 2)
    T, info = r.get_poses(1)
    oT, oi = O.rgbd_gauss_newton(data["ref_bgr"][0], data["ref_depth"][0], data["ref_bgr"][0], data["ref_depth"][0], K, levels=(2,))
    assert np.array_equal(T[0], np.eye(4)) and np.array_equal(oT, np.eye(4))
    assert info[0].iters_run[2] == 1 and info[0].updates[2] == 0 and info[0].eps_norm_first[2] < 200.0
    assert abs(info[0].eps_norm_first[2] - oi[0, 4]) <= 1e-10 * max(1.0, oi[0, 4])
    # the reference's asserts become status bits: too few (:497) / too many (:463) selected points
    r.gauss_newton(1, 2, dvo.rgbd_params(min_points=10**6))
    assert r.get_poses(1)[1][0].status & 1
    r.set_pose(1)
    r.gauss_newton(1, 2, dvo.rgbd_params(max_points=10))
    assert r.get_poses(1)[1][0].status & 2
    with pytest.raises(dvo.DvoError):          # level 0 has no Jacobian (:518)
        r.gauss_newton(1, 0)
    r.close()

"""Parity at the sizes BASELINE.json's configs name (the other GPU test files use small images where they can):

  config 2  batched edge alignment, 640x480x4, at a batch large enough that the solver takes the shape the benchmark
            times (max_batch > SM count -> 256 threads per pair, several CTAs per SM) -- GN-10 and SUBGRAD-50;
  config 3  photometric estimator (Huber + LM) at 640x480;
  config 4  1280x720, 5 levels, sub-gradient solver, consecutive-pair sequence with a key-frame switch.

Tolerances as north_star states them: poses within 1e-5 rad / 1e-5 m of the CPU oracle.
"""
import numpy as np
import pytest

import oracle_lib as O
import rgbd_odometry_b200 as dvo

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


# ------------------------------------------------------------------------------------------------ config 2
NFULL = 160            # > 148 SMs: the context picks the 256-thread solver shape, as the 1024-pair benchmark does
SAMPLE = (0, 1, 37, 80, 121, 159)


@pytest.fixture(scope="module")
def full_load():
    d = O.synth_batch(5000, NFULL)
    al = dvo.BatchAligner(640, 480, 4, max_batch=NFULL)
    al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"])
    al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
    al.build_pyramids(NFULL)
    al.prepare(NFULL)
    yield al, d
    al.close()


@pytest.mark.parametrize("solver,iters", [(dvo.GN, (10, 10, 10, 10)), (dvo.SUBGRAD_REF, (50, 50, 50, 50))])
def test_config2_full_load_solver_shape_matches_oracle(full_load, solver, iters):
    """solve_kernel<EXACT, REFERENCE, NEED_H = true / false, 256 threads> -- the instantiations bench.py times -- against
    the oracle on sampled pairs of a 160-pair launch at 640x480x4."""
    al, d = full_load
    prm = dvo.solver_params(solver=solver, iters=iters)
    al.set_initial_pose(NFULL, None)
    al.run(NFULL, prm)
    poses, info = al.get_poses(NFULL)
    assert all(inf.status == 0 for inf in info)
    for i in SAMPLE:
        o = O.align_pair(d["ref_gray"][i], d["ref_depth"][i], d["now_gray"][i], 4, iters, scfg=O.cfg(solver))
        R, T = poses[i, :9].reshape(3, 3), poses[i, 9:]
        assert rot_angle(R, o["R"]) < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5, \
            f"pair {i}: dR {rot_angle(R, o['R'])} dT {np.linalg.norm(T - o['T'])}"
        assert list(info[i].npts[:4]) == list(o["npts"])
        assert list(info[i].iterations_run[:4]) == list(o["iterations_run"])
        assert list(info[i].best_index[:4]) == list(o["best_index"])
        assert np.allclose(list(info[i].best_energy[:4]), o["best_energy"], rtol=1e-6)


def test_config2_pose_does_not_depend_on_launch_size(full_load):
    """The solver shape is a property of the context, so a pair's pose is bit-identical whether it is solved in a launch
    of 160 pairs, of 5 pairs or alone."""
    al, d = full_load
    prm = dvo.solver_params(solver=dvo.GN, iters=(10, 10, 10, 10))
    al.set_initial_pose(NFULL, None)
    al.run(NFULL, prm)
    want, _ = al.get_poses(NFULL)
    al.run(5, prm, first=33)
    assert np.array_equal(al.get_poses(5, first=33)[0], want[33:38])
    al.run(1, prm, first=159)
    assert np.array_equal(al.get_poses(1, first=159)[0], want[159:160])


# ------------------------------------------------------------------------------------------------ config 3
def test_config3_photometric_huber_lm_at_640x480():
    """EPoseEstimator path, corrected formulation with Huber weights and LM damping, levels 4 -> 0 at 640x480: normal
    equations at level 0 entry for entry (1e-9 relative) and the coarse-to-fine poses against the oracle."""
    W, H, L, K, NP = 640, 480, 5, O.K640, 2
    data = O.synth_batch(77, NP, W, H, K, bgr=True, now_depth=True)
    est = dvo.PhotoEstimator(W, H, L, max_batch=NP, intrinsics=K)
    est.set_frames(dvo.FRAME_REF, data["ref_bgr"], data["ref_depth"])
    est.set_frames(dvo.FRAME_NOW, data["now_bgr"], data["now_depth"])
    est.prepare_ref(NP, compat=False)
    for i in range(NP):                                  # one evaluation at full resolution: A, b, canvas, counts
        O.photo_build_ref_level(data["ref_bgr"][i], data["ref_depth"][i], 0, K, compat=False)
        now = O.photo_now_level(data["now_bgr"][i], 0)
        o = O.photo_evaluate(now, data["R"][i], data["T"][i], K, compat=False, huber_k=10.0, want_canvas=True)
        g = est.eval(i, 0, data["R"][i], data["T"][i], compat=False, huber_k=10.0, want_canvas=True)
        assert g["nreproj"] == o["nreproj"] and g["nused"] == o["nused"]
        assert np.array_equal(g["canvas"], o["canvas"])
        assert np.allclose(g["b"], o["b"], rtol=1e-9, atol=1e-9 * np.abs(o["b"]).max())
        assert np.allclose(g["A"], o["A"], rtol=1e-9, atol=1e-9 * np.abs(o["A"]).max())
    est.set_pose(NP, None)
    Ro = [np.eye(3) for _ in range(NP)]
    To = [np.zeros(3) for _ in range(NP)]
    for l in (4, 3, 2, 1, 0):
        est.estimate(NP, l, iters=6, compat=False, huber_k=10.0, lambda0=1e-3)
        poses, info = est.get_poses(NP)
        for i in range(NP):
            O.photo_build_ref_level(data["ref_bgr"][i], data["ref_depth"][i], l, K, compat=False)
            o = O.photo_estimate(O.photo_now_level(data["now_bgr"][i], l), Ro[i], To[i], 6, K, compat=False, huber_k=10.0, lambda0=1e-3)
            Ro[i], To[i] = o["R"], o["T"]
            R, T = poses[i, :9].reshape(3, 3), poses[i, 9:]
            assert info[i].status == o["status"] and info[i].iters_run == o["iters_run"], f"L{l} pair {i}"
            assert rot_angle(R, o["R"]) < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5, f"L{l} pair {i}"
            assert abs(info[i].sumsq_last - o["sumsq_last"]) <= 1e-9 * o["sumsq_last"]
    est.close()


# ------------------------------------------------------------------------------------------------ config 4
def test_config4_sequence_1280x720_5_levels_subgradient():
    """SolveDVO::loop at 1280x720, 5 levels, the shipped sub-gradient solver with 50 iterations per level: two sequences of
    7 frames (key-frame switch at frame 5), device result against the oracle-driven loop frame by frame.  The level-0
    Canny takes the global-bitmap path at this size."""
    from test_gpu_sequence import oracle_sequence
    W, H, L, K = 1280, 720, 5, O.K1280
    nseq, nframes = 2, 7
    seqs = [O.synth_sequence(170 + s, nframes, W, H, K, max_angle_deg=0.4, max_trans_m=0.008) for s in range(nseq)]
    gray = np.stack([s[0] for s in seqs]); depth = np.stack([s[1] for s in seqs])
    iters = (50, 50, 50, 50, 50)
    prm = dvo.solver_params(iters=iters)
    al = dvo.BatchAligner(W, H, L, max_batch=nseq, keep_now_depth=True, intrinsics=K)
    rel, kind, glob = al.run_sequences(gray, depth, prm)
    rel_g, kind_g, reason_g, glob_g = al.run_sequences_gated(gray, depth, prm, dvo.keyframe_policy())
    assert np.array_equal(kind, kind_g) and np.array_equal(rel, rel_g) and np.array_equal(glob, glob_g)
    for s in range(nseq):
        orel, okind, oglob = oracle_sequence(gray[s], depth[s], L, iters, K)
        assert np.array_equal(kind[s], okind), (kind[s], okind)
        assert list(np.nonzero(okind == 2)[0]) == [4]
        for t in range(nframes):
            assert rot_angle(rel[s, t, :9].reshape(3, 3), orel[t, :9].reshape(3, 3)) < 1e-5, (s, t)
            assert np.linalg.norm(rel[s, t, 9:] - orel[t, 9:]) < 1e-5, (s, t)
            assert rot_angle(glob[s, t, :9].reshape(3, 3), oglob[t, :9].reshape(3, 3)) < 1e-5
            assert np.linalg.norm(glob[s, t, 9:12] - oglob[t, 9:12]) < 1e-5
    al.close()


def test_config4_back_to_back_process_calls_at_1280x720_are_bit_identical():
    """Images too large for shared-memory hysteresis bitmaps use a per-slot global scratch; back-to-back dvo_process calls run
    their half batches on two internal streams without joining, so two Canny launches of different calls can be in flight
    at once.  Results must equal the staged calls bit for bit."""
    W, H, L, K = 1280, 720, 5, O.K1280
    nb = 8
    d = O.synth_batch(300, nb, W, H, K)
    n = 4 * 148                                        # the overlapped path needs count >= 4 x SM count
    idx = np.arange(n) % nb
    rg, rd, ng = d["ref_gray"][idx], d["ref_depth"][idx], d["now_gray"][idx]
    al = dvo.BatchAligner(W, H, L, max_batch=n, intrinsics=K)
    al.set_frames(dvo.FRAME_REF, rg, rd); al.set_frames(dvo.FRAME_NOW, ng, None)
    prm = dvo.solver_params(solver=dvo.GN, iters=(4, 4, 4, 4, 4))
    al.build_pyramids(n); al.prepare(n); al.run(n, prm)
    want, _ = al.get_poses(n)
    edge_want = al.get_level_buffer(n - 1, 1, 0, "edge")
    for _ in range(3):
        al.process(n, prm)
    al.join(); al.synchronize()
    got, _ = al.get_poses(n)
    assert np.array_equal(got, want)
    assert np.array_equal(al.get_level_buffer(n - 1, 1, 0, "edge"), edge_want)
    for i in range(nb, n):                             # replicated inputs -> replicated results
        assert np.array_equal(got[i], got[i % nb])
    al.close()

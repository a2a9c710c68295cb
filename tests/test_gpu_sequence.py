"""Consecutive-pair odometry (SolveDVO::loop, src/SolveDVO.cpp:1896-2373): warm start, key frame every 5 frames = the
previous frame, pose reset + re-solve, GOP composition -- GPU (dvo_run_sequences) against the same schedule driven
through the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
import rgbd_odometry_b200 as dvo

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def oracle_sequence(gray, depth, levels, iters, K, keyframe_every=5):
    n = gray.shape[0]
    rel = np.zeros((n, 12)); kind = np.zeros(n, np.int32)
    rel[0, [0, 4, 8]] = 1.0; kind[0] = 1
    ref = 0; last_ref = 0
    R, T = np.eye(3), np.zeros(3)
    for t in range(1, n):
        if (t - last_ref) == keyframe_every and last_ref != t - 1:
            last_ref = t - 1; ref = t - 1; kind[t - 1] = 2
            R, T = np.eye(3), np.zeros(3)
        o = O.align_pair(gray[ref], depth[ref], gray[t], levels, iters, K, R0=R, T0=T)
        R, T = o["R"], o["T"]
        rel[t, :9] = R.reshape(9); rel[t, 9:] = T
    glob, is_key, _ = O.gop_replay(kind, np.full(n, 5, np.int32), rel)
    return rel, kind, glob


@pytest.mark.parametrize("W,H,L,K", [(320, 240, 4, (262.5, 262.5, 159.5, 119.5))])
def test_sequences_match_oracle_schedule(W, H, L, K):
    nseq, nframes = 3, 12
    seqs = [O.synth_sequence(70 + s, nframes, W, H, K, max_angle_deg=0.4, max_trans_m=0.008) for s in range(nseq)]
    gray = np.stack([s[0] for s in seqs]); depth = np.stack([s[1] for s in seqs])
    iters = (8, 8, 8, 8)
    al = dvo.BatchAligner(W, H, L, max_batch=nseq, keep_now_depth=True, intrinsics=K)
    rel, kind, glob = al.run_sequences(gray, depth, dvo.solver_params(iters=iters))
    for s in range(nseq):
        orel, okind, oglob = oracle_sequence(gray[s], depth[s], L, iters, K)
        assert np.array_equal(kind[s], okind), (kind[s], okind)
        assert list(np.nonzero(okind == 2)[0]) == [4, 8]
        for t in range(nframes):
            assert rot_angle(rel[s, t, :9].reshape(3, 3), orel[t, :9].reshape(3, 3)) < 1e-5
            assert np.linalg.norm(rel[s, t, 9:] - orel[t, 9:]) < 1e-5
            assert rot_angle(glob[s, t, :9].reshape(3, 3), oglob[t, :9].reshape(3, 3)) < 1e-5
            assert np.linalg.norm(glob[s, t, 9:12] - oglob[t, 9:12]) < 1e-5
    # global poses follow the true trajectory (loose: the estimator's own accuracy, not a parity bound)
    Rw, Tw = seqs[0][2], seqs[0][3]
    err = np.linalg.norm(glob[0, -1, 9:12] - Tw[-1])
    assert err < np.linalg.norm(Tw[-1]) + 0.05
    al.close()


def test_promote_requires_a_previous_now_frame():
    al = dvo.BatchAligner(160, 120, 2, max_batch=1, keep_now_depth=True, intrinsics=(131.25, 131.25, 79.5, 59.5))
    with pytest.raises(dvo.DvoError, match="previous now frame"):
        al.promote_now_to_ref(1)
    al2 = dvo.BatchAligner(160, 120, 2, max_batch=1, keep_now_depth=False)
    with pytest.raises(dvo.DvoError, match="keep_now_depth"):
        al2.promote_now_to_ref(1)
    al.close(); al2.close()


def oracle_sequence_gated(gray, depth, levels, iters, K, every, gates, b_thr, vis_thr, min_reproj):
    """SolveDVO::loop with the key-frame policy of src/SolveDVO.cpp:2117-2233 (quality gates as designed, OR periodic)."""
    n = gray.shape[0]
    rel = np.zeros((n, 12)); kind = np.zeros(n, np.int32); reason = np.zeros(n, np.int32)
    rel[0, [0, 4, 8]] = 1.0; kind[0] = 1; reason[0] = 1
    ref = 0; last_ref = 0
    finest = [l for l in range(levels) if iters[l] > 0][0]
    R, T = np.eye(3), np.zeros(3)
    stats = []
    for t in range(1, n):
        o = O.align_pair(gray[ref], depth[ref], gray[t], levels, iters, K, R0=R, T0=T, stats=True)
        stats.append((o["b_cap"], float(o["visible_ratio"][finest]), int(o["npts"][finest])))
        signal, why = False, 0
        if gates:
            if o["b_cap"] > b_thr: signal, why = True, 2
            if o["visible_ratio"][finest] < vis_thr: signal, why = True, 3
            if o["npts"][finest] < min_reproj: signal, why = True, 4
        if every > 0 and (t - last_ref) == every: signal, why = True, 5
        if signal and last_ref != t - 1:
            last_ref = t - 1; ref = t - 1; kind[t - 1] = 2; reason[t - 1] = why
            o = O.align_pair(gray[ref], depth[ref], gray[t], levels, iters, K, R0=np.eye(3), T0=np.zeros(3))
        R, T = o["R"], o["T"]
        rel[t, :9] = R.reshape(9); rel[t, 9:] = T
    glob, _, _ = O.gop_replay(kind, np.maximum(reason, 1), rel)
    return rel, kind, reason, glob, stats


def test_gated_keyframe_policy_on_device():
    """dvo_run_sequences_gated: (a) with the shipped policy it reproduces dvo_run_sequences; (b) with the quality gates
    enabled every sequence switches key frames where ITS OWN residual statistic / visible ratio says so, matching the
    oracle-driven loop frame by frame (switch frames, reasons, relative and global poses)."""
    W, H, L, K = 320, 240, 4, (262.5, 262.5, 159.5, 119.5)
    nseq, nframes = 3, 10
    seqs = [O.synth_sequence(70 + s, nframes, W, H, K, max_angle_deg=0.4, max_trans_m=0.008) for s in range(nseq)]
    gray = np.stack([s[0] for s in seqs]); depth = np.stack([s[1] for s in seqs])
    iters = (8, 8, 8, 8)
    prm = dvo.solver_params(iters=iters)
    al = dvo.BatchAligner(W, H, L, max_batch=nseq, keep_now_depth=True, intrinsics=K)
    rel0, kind0, glob0 = al.run_sequences(gray, depth, prm)
    rel1, kind1, reason1, glob1 = al.run_sequences_gated(gray, depth, prm, dvo.keyframe_policy())
    assert np.array_equal(kind0, kind1) and np.array_equal(rel0, rel1) and np.array_equal(glob0, glob1)
    assert sorted(set(reason1[kind1 == 2].tolist())) == [5] and (reason1[:, 0] == 1).all()
    b_thr, vis_thr = 9.5, 0.95
    pol = dvo.keyframe_policy(keyframe_every=0, use_quality_gates=True, laplacian_thresh=b_thr, visible_ratio_thresh=vis_thr)
    rel, kind, reason, glob = al.run_sequences_gated(gray, depth, prm, pol)
    switches = []
    for s in range(nseq):
        orel, okind, oreason, oglob, stats = oracle_sequence_gated(gray[s], depth[s], L, iters, K, 0, True, b_thr, vis_thr, 50)
        # the thresholds sit well away from every statistic the decision saw, so fp32-vs-fp64 means cannot flip a gate
        assert min(abs(b - b_thr) for b, _, _ in stats) > 1e-3 and min(abs(v - vis_thr) for _, v, _ in stats) > 1e-4
        assert np.array_equal(kind[s], okind), (s, kind[s], okind)
        assert np.array_equal(reason[s], oreason), (s, reason[s], oreason)
        for t in range(nframes):
            assert rot_angle(rel[s, t, :9].reshape(3, 3), orel[t, :9].reshape(3, 3)) < 1e-5 and np.linalg.norm(rel[s, t, 9:] - orel[t, 9:]) < 1e-5
            assert rot_angle(glob[s, t, :9].reshape(3, 3), oglob[t, :9].reshape(3, 3)) < 1e-5 and np.linalg.norm(glob[s, t, 9:12] - oglob[t, 9:12]) < 1e-5
        switches.append(tuple(np.nonzero(okind == 2)[0]))
    assert any(len(sw) for sw in switches) and len(set(switches)) > 1, switches      # the sequences really diverged
    al.close()


def test_sequences_from_device_memory_and_pinned_host_match():
    """dvo_run_sequences_mem: the same sequences given as pageable host arrays, pinned host arrays (asynchronous, pipelined uploads)
    and device-resident buffers give bit-identical records."""
    import torch
    W, H, L, K = 320, 240, 3, (262.5, 262.5, 159.5, 119.5)
    nseq, nframes = 4, 8
    seqs = [O.synth_sequence(90 + s, nframes, W, H, K, max_angle_deg=0.4, max_trans_m=0.008) for s in range(nseq)]
    gray = np.stack([s[0] for s in seqs]); depth = np.stack([s[1] for s in seqs])
    prm = dvo.solver_params(iters=(8, 8, 8))
    al = dvo.BatchAligner(W, H, L, max_batch=nseq, keep_now_depth=True, intrinsics=K)
    for pol in (dvo.keyframe_policy(keyframe_every=3), dvo.keyframe_policy(keyframe_every=0, use_quality_gates=True, laplacian_thresh=9.5, visible_ratio_thresh=0.95)):
        want = al.run_sequences_mem(gray, depth, nseq, nframes, prm, pol)
        pg, pd = torch.from_numpy(gray).pin_memory(), torch.from_numpy(depth.view(np.int16)).pin_memory()
        got_pin = al.run_sequences_mem(pg.data_ptr(), pd.data_ptr(), nseq, nframes, prm, pol)
        dg, dd = pg.cuda(), pd.cuda()
        torch.cuda.synchronize()
        got_dev = al.run_sequences_mem(dg.data_ptr(), dd.data_ptr(), nseq, nframes, prm, pol, device=True)
        for a, b, c in zip(want, got_pin, got_dev):
            assert np.array_equal(a, b) and np.array_equal(a, c)
    rel, kind, glob = al.run_sequences(gray, depth, prm, keyframe_every=3)
    w = al.run_sequences_mem(gray, depth, nseq, nframes, prm, dvo.keyframe_policy(keyframe_every=3))
    assert np.array_equal(rel, w[0]) and np.array_equal(kind, w[1]) and np.array_equal(glob, w[3])
    assert list(np.nonzero(kind[0] == 2)[0]) == [2, 4, 6]
    al.close()

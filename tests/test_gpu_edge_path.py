"""GPU parity tests of the edge-alignment path (SolveDVO) against the CPU oracle, through the C-ABI.

Bit-exact: pyramid indexing, Canny edge maps, integer d2, normalised DT + gradients (IEEE fp32, same operation
order), reference point lists, per-point Jacobians / residuals / weights in EXACT arithmetic.
Tolerance (stated per test): fp64 reductions (summation order differs), poses, FAST arithmetic.
"""
import numpy as np
import pytest

import oracle_lib as O
import rgbd_odometry_b200 as dvo

pytestmark = pytest.mark.gpu

NPAIR = 3
LEVELS = 4
ITERS = (50, 50, 50, 50)


def rot_angle(Ra, Rb):
    R = Ra.T @ Rb
    return float(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def mismatch(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    bad = np.argwhere(a != b)
    return f"{len(bad)} of {a.size} differ; first at {bad[:5].tolist()} gpu={a[tuple(bad[0])] if len(bad) else None} cpu={b[tuple(bad[0])] if len(bad) else None}"


@pytest.fixture(scope="module")
def batch():
    return O.synth_batch(100, NPAIR, now_depth=True)


@pytest.fixture(scope="module")
def prepared(batch):
    al = dvo.BatchAligner(640, 480, LEVELS, max_batch=NPAIR + 1, keep_now_depth=True, trace_iters=50)
    # use a non-zero slot offset to exercise `first`
    al.set_frames(dvo.FRAME_REF, batch["ref_gray"], batch["ref_depth"], first=1)
    al.set_frames(dvo.FRAME_NOW, batch["now_gray"], batch["now_depth"], first=1)
    al.build_pyramids(NPAIR, first=1)
    al.prepare(NPAIR, first=1)
    al.synchronize()
    yield al
    al.close()


@pytest.fixture(scope="module")
def oracle_levels(batch):
    out = []
    for i in range(NPAIR):
        ref = [O.preprocess_level(batch["ref_gray"][i], batch["ref_depth"][i], l) for l in range(LEVELS)]
        now = [O.preprocess_level(batch["now_gray"][i], batch["now_depth"][i], l) for l in range(LEVELS)]
        out.append((ref, now))
    return out


def test_pyramid_bit_exact(prepared, oracle_levels):
    for i in range(NPAIR):
        for f, name in ((0, "ref"), (1, "now")):
            for l in range(LEVELS):
                o = oracle_levels[i][f][l]
                g = prepared.get_level_buffer(i + 1, f, l, "gray")
                assert np.array_equal(g, o["gray"]), f"gray {name} pair {i} L{l}: {mismatch(g, o['gray'])}"
                d = prepared.get_level_buffer(i + 1, f, l, "depth")
                assert np.array_equal(d, o["depth"]), f"depth {name} pair {i} L{l}: {mismatch(d, o['depth'])}"


def test_canny_bit_exact(prepared, oracle_levels):
    for i in range(NPAIR):
        for f, name in ((0, "ref"), (1, "now")):
            for l in range(LEVELS):
                e = prepared.get_level_buffer(i + 1, f, l, "edge")
                o = oracle_levels[i][f][l]["edge"]
                assert np.array_equal(e, o), f"edge {name} pair {i} L{l}: {mismatch(e, o)}"


def test_edt_d2_bit_exact(prepared, oracle_levels):
    for i in range(NPAIR):
        for l in range(LEVELS):
            d2 = prepared.get_level_buffer(i + 1, 1, l, "d2")
            o = oracle_levels[i][1][l]["d2"]
            assert np.array_equal(d2, o), f"d2 pair {i} L{l}: {mismatch(d2, o)}"


def test_dt_and_gradient_bit_exact(prepared, oracle_levels):
    for i in range(NPAIR):
        for l in range(LEVELS):
            for k in ("dtn", "gx", "gy"):
                a = prepared.get_level_buffer(i + 1, 1, l, k)
                o = oracle_levels[i][1][l][k]
                assert bits_equal(a, o), f"{k} pair {i} L{l}: {mismatch(a, o)} maxabs {np.abs(a - o).max()}"


def test_reference_points_bit_exact(prepared, oracle_levels):
    for i in range(NPAIR):
        for l in range(LEVELS):
            ref = oracle_levels[i][0][l]
            X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], l)
            gX, gY, gZ = prepared.get_points(i + 1, l)
            assert len(gX) == len(X), f"npts pair {i} L{l}: {len(gX)} vs {len(X)}"
            assert len(X) > 0
            assert bits_equal(gX, X) and bits_equal(gY, Y) and bits_equal(gZ, Z), f"points pair {i} L{l}"


def _true_pose(batch, i):
    return batch["R"][i], batch["T"][i]


@pytest.mark.parametrize("jac", [dvo.JAC_REFERENCE, dvo.JAC_EXACT])
@pytest.mark.parametrize("weight", [dvo.W_REF_CAUCHY, dvo.W_HUBER, dvo.W_NONE])
def test_normal_equations_exact_arithmetic(prepared, oracle_levels, batch, jac, weight):
    """Per-point J / eps / w bit-exact; H = J^T W J and g = J^T W eps to 1e-12 relative (spec: 1e-5)."""
    for i in range(NPAIR):
        for l in range(LEVELS):
            ref, now = oracle_levels[i][0][l], oracle_levels[i][1][l]
            X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], l)
            for (R, T) in ((np.eye(3), np.zeros(3)), _true_pose(batch, i)):
                o = O.evaluate(X, Y, Z, now["dtn"], now["gx"], now["gy"], l, R, T, jac=jac, weight=weight, per_point=True)
                g = prepared.eval_normal_equations(i + 1, l, R, T, jacobian=jac, weight=weight, per_point=True, npts=len(X))
                assert g["nvis"] == o["nvis"]
                assert bits_equal(g["u"], o["u"]) and bits_equal(g["v"], o["v"]), f"reprojections pair {i} L{l}"
                assert bits_equal(g["eps"], o["eps"]), f"eps pair {i} L{l}"
                assert bits_equal(g["w"], o["w"]), f"w pair {i} L{l}: {mismatch(g['w'], o['w'])}"
                assert bits_equal(g["J"], o["J"]), f"J pair {i} L{l}: {mismatch(g['J'], o['J'])}"
                scale_H = np.abs(o["H"]).max()
                # the GPU mirrors the upper triangle; the oracle's lower triangle sums fl(J_c w) J_r instead of
                # fl(J_r w) J_c, a ~1e-8 relative asymmetry -- far inside the 1e-5 spec
                assert np.abs(g["H"] - o["H"]).max() <= 1e-6 * scale_H, f"H pair {i} L{l}"
                # upper triangle is accumulated exactly as the oracle does
                iu = np.triu_indices(6)
                assert np.allclose(g["H"][iu], o["H"][iu], rtol=1e-12, atol=1e-12 * scale_H)
                assert np.allclose(g["g"], o["g"], rtol=1e-12, atol=1e-12 * np.abs(o["g"]).max())
                assert abs(g["sumsq"] - o["sumsq"]) <= 1e-12 * o["sumsq"]


def test_normal_equations_fast_arithmetic_within_spec(prepared, oracle_levels, batch):
    """FAST arithmetic (FMA contraction, approximate reciprocals in the Jacobian): J^T W J within relative 1e-5 (north_star)."""
    for i in range(NPAIR):
        for l in range(LEVELS):
            ref, now = oracle_levels[i][0][l], oracle_levels[i][1][l]
            X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], l)
            R, T = _true_pose(batch, i)
            o = O.evaluate(X, Y, Z, now["dtn"], now["gx"], now["gy"], l, R, T)
            g = prepared.eval_normal_equations(i + 1, l, R, T, arithmetic=dvo.ARITH_FAST)
            # FAST keeps the reprojection IEEE-exact, so the same texels are gathered; only the Jacobian math is contracted
            assert g["nvis"] == o["nvis"]
            assert np.abs(g["H"] - o["H"]).max() <= 1e-5 * np.abs(o["H"]).max(), f"H fast pair {i} L{l}"
            assert np.abs(g["g"] - o["g"]).max() <= 1e-5 * np.abs(o["g"]).max(), f"g fast pair {i} L{l}"


def _run_gpu(prepared, params):
    prepared.set_initial_pose(NPAIR, None, first=1)
    prepared.run(NPAIR, params, first=1)
    return prepared.get_poses(NPAIR, first=1)


def test_subgradient_solver_pose_parity(prepared, batch):
    """SUBGRAD_REF (the shipped solver), EXACT arithmetic: pose within 1e-5 rad / 1e-5 m of the oracle (north_star),
    per-iteration g / H within relative 1e-5, same best-iterate bookkeeping."""
    params = dvo.solver_params(iters=ITERS)
    poses, info = _run_gpu(prepared, params)
    for i in range(NPAIR):
        o = O.align_pair(batch["ref_gray"][i], batch["ref_depth"][i], batch["now_gray"][i], LEVELS, ITERS, trace=True)
        R, T = poses[i, :9].reshape(3, 3), poses[i, 9:]
        assert rot_angle(R, o["R"]) < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5, \
            f"pair {i}: dR {rot_angle(R, o['R'])} dT {np.linalg.norm(T - o['T'])}"
        assert list(info[i].npts[:LEVELS]) == list(o["npts"])
        assert list(info[i].iterations_run[:LEVELS]) == list(o["iterations_run"])
        assert list(info[i].best_index[:LEVELS]) == list(o["best_index"]), f"best idx {list(info[i].best_index[:LEVELS])} vs {o['best_index']}"
        assert np.allclose(list(info[i].best_energy[:LEVELS]), o["best_energy"], rtol=1e-6)
        assert np.allclose(list(info[i].visible_ratio[:LEVELS]), o["visible_ratio"], rtol=1e-6)
        for l in range(LEVELS):
            tr = prepared.get_trace(i + 1, l)
            n = o["iterations_run"][l]
            ot = {k: v[l, :n] for k, v in o["trace"].items()}
            assert np.allclose(tr["energy"][:n], ot["energy"], rtol=1e-6), f"energy trace pair {i} L{l}"
            assert np.array_equal(tr["nvis"][:n], ot["nvis"]), f"nvis trace pair {i} L{l}"
            gs = np.abs(ot["g"]).max(axis=1, keepdims=True)
            assert (np.abs(tr["g"][:n] - ot["g"]) <= 1e-5 * gs).all(), f"g trace pair {i} L{l}"
            Hs = np.abs(ot["H"]).max(axis=(1, 2), keepdims=True)
            assert (np.abs(tr["H"][:n] - ot["H"]) <= 1e-5 * Hs).all(), f"H trace pair {i} L{l}"
            assert np.abs(tr["T"][:n] - ot["T"]).max() < 1e-5


@pytest.mark.parametrize("solver,jac,weight", [(dvo.GN, dvo.JAC_REFERENCE, dvo.W_REF_CAUCHY), (dvo.GN, dvo.JAC_EXACT, dvo.W_HUBER),
                                               (dvo.LM, dvo.JAC_EXACT, dvo.W_REF_CAUCHY), (dvo.LM, dvo.JAC_REFERENCE, dvo.W_NONE)])
def test_gauss_newton_and_lm_pose_parity(prepared, batch, solver, jac, weight):
    iters = (10, 10, 10, 10)
    params = dvo.solver_params(solver=solver, jacobian=jac, weight=weight, iters=iters)
    poses, info = _run_gpu(prepared, params)
    for i in range(NPAIR):
        o = O.align_pair(batch["ref_gray"][i], batch["ref_depth"][i], batch["now_gray"][i], LEVELS, iters,
                         scfg=O.cfg(solver, jac, weight))
        R, T = poses[i, :9].reshape(3, 3), poses[i, 9:]
        assert rot_angle(R, o["R"]) < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5, \
            f"pair {i}: dR {rot_angle(R, o['R'])} dT {np.linalg.norm(T - o['T'])}"
        assert list(info[i].iterations_run[:LEVELS]) == list(o["iterations_run"])
        assert list(info[i].best_index[:LEVELS]) == list(o["best_index"])


def test_interpolated_dt_residual_bit_exact(prepared, oracle_levels, batch):
    """The reference's compiled-out __INTERPOLATE_DISTANCE_TRANSFORM variant (src/SolveDVO.cpp:443-444, :1285-1308):
    per-point eps / w / J bit-exact against the oracle's restatement, sums to 1e-12 relative."""
    for i in range(NPAIR):
        for l in range(LEVELS):
            ref, now = oracle_levels[i][0][l], oracle_levels[i][1][l]
            X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], l)
            for (R, T) in ((np.eye(3), np.zeros(3)), _true_pose(batch, i)):
                o = O.evaluate(X, Y, Z, now["dtn"], now["gx"], now["gy"], l, R, T, per_point=True, residual=O.RES_DT_INTERP)
                g = prepared.eval_normal_equations(i + 1, l, R, T, per_point=True, npts=len(X), residual=dvo.RES_DT_INTERP)
                f = prepared.eval_normal_equations(i + 1, l, R, T, per_point=True, npts=len(X))
                assert g["nvis"] == o["nvis"]
                assert bits_equal(g["eps"], o["eps"]), f"eps pair {i} L{l}: {mismatch(g['eps'], o['eps'])}"
                assert bits_equal(g["w"], o["w"]), f"w pair {i} L{l}: {mismatch(g['w'], o['w'])}"
                assert bits_equal(g["J"], o["J"]) and bits_equal(g["J"], f["J"])       # the Jacobian keeps the floor texel (:376-385)
                assert not np.array_equal(g["eps"], f["eps"])                          # and the residual really is a different one
                iu = np.triu_indices(6)
                assert np.allclose(g["H"][iu], o["H"][iu], rtol=1e-12, atol=1e-12 * np.abs(o["H"]).max())
                assert np.allclose(g["g"], o["g"], rtol=1e-12, atol=1e-12 * np.abs(o["g"]).max())


@pytest.mark.parametrize("solver,iters", [(dvo.SUBGRAD_REF, (30, 30, 30, 30)), (dvo.GN, (8, 8, 8, 8))])
def test_interpolated_dt_residual_pose_parity(prepared, batch, solver, iters):
    params = dvo.solver_params(solver=solver, iters=iters, residual=dvo.RES_DT_INTERP)
    poses, info = _run_gpu(prepared, params)
    for i in range(NPAIR):
        o = O.align_pair(batch["ref_gray"][i], batch["ref_depth"][i], batch["now_gray"][i], LEVELS, iters,
                         scfg=O.cfg(solver, residual=O.RES_DT_INTERP))
        R, T = poses[i, :9].reshape(3, 3), poses[i, 9:]
        assert rot_angle(R, o["R"]) < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5, \
            f"pair {i}: dR {rot_angle(R, o['R'])} dT {np.linalg.norm(T - o['T'])}"
        assert list(info[i].best_index[:LEVELS]) == list(o["best_index"])
    with pytest.raises(dvo.DvoError):                                                  # EXACT arithmetic only
        prepared.run(NPAIR, dvo.solver_params(iters=iters, residual=dvo.RES_DT_INTERP, arithmetic=dvo.ARITH_FAST), first=1)


def test_align_batch_end_to_end_matches_staged_calls(batch):
    """dvo_align_batch (host buffers, chunked over max_batch) == staged calls == oracle."""
    al = dvo.BatchAligner(640, 480, LEVELS, max_batch=2)
    params = dvo.solver_params(iters=(8, 8, 8, 8))
    poses, info = al.align_batch(batch["ref_gray"], batch["ref_depth"], batch["now_gray"], params)
    assert al.launch_count() > 0
    for i in range(NPAIR):
        o = O.align_pair(batch["ref_gray"][i], batch["ref_depth"][i], batch["now_gray"][i], LEVELS, (8, 8, 8, 8))
        R, T = poses[i, :9].reshape(3, 3), poses[i, 9:]
        assert rot_angle(R, o["R"]) < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5
        assert info[i].status == 0
    al.close()


def test_process_two_stream_schedule_is_bit_identical():
    """dvo_process (two staggered half batches on internal streams, poses copied per half, deferred join) gives exactly
    the staged calls' results, also across back-to-back calls and when other entry points are interleaved."""
    import torch
    n = 600                                   # >= 4 x 148 SMs, so the overlapped path is taken
    d = O.synth_batch(900, 8, 160, 120, (131.25, 131.25, 79.5, 59.5))
    rep = lambda a: np.ascontiguousarray(np.tile(a, (n // 8, 1, 1)))
    rg, rd, ng = rep(d["ref_gray"]), rep(d["ref_depth"]), rep(d["now_gray"])
    ng[1::2] = np.roll(ng[1::2], 1, axis=2)                       # make the odd pairs different problems
    al = dvo.BatchAligner(160, 120, 3, max_batch=n, intrinsics=(131.25, 131.25, 79.5, 59.5))
    al.set_frames(dvo.FRAME_REF, rg, rd); al.set_frames(dvo.FRAME_NOW, ng, None)
    prm = dvo.solver_params(solver=dvo.GN, iters=(6, 6, 6))
    al.build_pyramids(n); al.prepare(n); al.run(n, prm)
    want, winfo = al.get_poses(n)
    for i in (0, 1, 2, 3):                     # max_batch > SM count: the 256-thread solver shape (also pinned at 640x480 in test_gpu_config_sizes.py)
        o = O.align_pair(rg[i], rd[i], ng[i], 3, (6, 6, 6), (131.25, 131.25, 79.5, 59.5), scfg=O.cfg(O.GN))
        assert rot_angle(want[i, :9].reshape(3, 3), o["R"]) < 1e-5 and np.linalg.norm(want[i, 9:] - o["T"]) < 1e-5
    out = torch.zeros((n, 12), dtype=torch.float64, device="cuda")
    al.set_frames(dvo.FRAME_NOW, ng[::-1].copy(), None)            # scramble, then restore: process must rebuild everything
    al.process(n, prm)
    al.set_frames(dvo.FRAME_NOW, ng, None)                         # joins before it touches the slots
    for _ in range(3):
        al.process(n, prm, poses_out=out.data_ptr())               # back to back, no join in between
    al.join(); al.synchronize()
    got, ginfo = al.get_poses(n)
    assert np.array_equal(got, want) and np.array_equal(out.cpu().numpy(), want)
    assert all(list(ginfo[i].iterations_run[:3]) == list(winfo[i].iterations_run[:3]) for i in range(n))
    al.process(100, prm, first=7)                                  # small range: staged fallback; the solver shape is a property of the
    assert np.array_equal(al.get_poses(100, first=7)[0], want[7:107])                    # context (max_batch), so a sub-range is bit-identical too
    al.process(n, prm)                                             # a different split while work is in flight: joins first
    al.process(n - 50, prm, first=50)
    al.join(); al.synchronize()
    assert np.array_equal(al.get_poses(n)[0], want)
    al.close()


def test_blank_and_degenerate_inputs():
    """No edges anywhere: the reference asserts nSelectedPts > 0 (src/SolveDVO.cpp:282); here status bit 0 is set, the
    pose stays at its initial value, d2 is the EDT_INF sentinel and DTn is 0 -- exactly what the oracle produces."""
    al = dvo.BatchAligner(160, 120, 3, max_batch=2)
    gray = np.full((2, 120, 160), 77, np.uint8)
    depth = np.full((2, 120, 160), 1500, np.uint16)
    # slot 1: one bright square on the reference only -> ref has points, now has no edges
    gray_ref = gray.copy()
    gray_ref[1, 40:80, 50:100] = 200
    poses, info = al.align_batch(gray_ref, depth, gray, dvo.solver_params(iters=(5, 5, 5)))
    for i in range(2):
        assert info[i].status & 1
        assert np.allclose(poses[i, :9].reshape(3, 3), np.eye(3)) and np.allclose(poses[i, 9:], 0)
    d2 = al.get_level_buffer(0, 1, 0, "d2")
    assert (d2 == (1 << 28)).all()
    assert (al.get_level_buffer(0, 1, 0, "dtn") == 0).all()
    assert (al.get_level_buffer(0, 1, 0, "edge") == 0).all()
    e = al.get_level_buffer(1, 0, 0, "edge")
    assert np.array_equal(e, O.canny(gray_ref[1]))
    al.close()


@pytest.mark.parametrize("W,H,L", [(100, 75, 3), (82, 61, 2), (1280, 720, 5), (320, 240, 4), (2112, 96, 2), (200, 152, 2), (132, 100, 1)])
def test_ragged_and_large_sizes(W, H, L):
    """Odd sizes (cvRound half-to-even level dims, non-multiple-of-32 widths), 1280x720x5 (BASELINE config 4;
    hysteresis bitmaps exceed shared memory -> global scratch path), a very wide image (EDT bisection fallback) and widths
    that are a multiple of 4 but not of 32 / 128 (the vector path of sobel_nms_kernel with a partial last chunk and word)."""
    K = (525.0 * W / 640, 525.0 * W / 640, (W - 1) / 2.0, (H - 1) / 2.0)
    d = O.synth_pair(7, W, H, K)
    al = dvo.BatchAligner(W, H, L, max_batch=1, intrinsics=K)
    al.set_frames(dvo.FRAME_REF, d["ref_gray"][None], d["ref_depth"][None])
    al.set_frames(dvo.FRAME_NOW, d["now_gray"][None], None)
    al.build_pyramids(1)
    al.prepare(1)
    for l in range(L):
        w, h = al.level_dims(l)
        assert (w, h) == (O.level_dim(W, l), O.level_dim(H, l))
        ref = O.preprocess_level(d["ref_gray"], d["ref_depth"], l)
        now = O.preprocess_level(d["now_gray"], None, l)
        assert np.array_equal(al.get_level_buffer(0, 0, l, "edge"), ref["edge"]), f"ref edge L{l}"
        assert np.array_equal(al.get_level_buffer(0, 1, l, "edge"), now["edge"]), f"now edge L{l}"
        assert np.array_equal(al.get_level_buffer(0, 1, l, "d2"), now["d2"]), f"d2 L{l}"
        for k in ("dtn", "gx", "gy"):
            assert bits_equal(al.get_level_buffer(0, 1, l, k), now[k]), f"{k} L{l}"
        X, Y, Z, _, _ = O.select_points(ref["edge"], ref["depth"], l, K)
        gX, gY, gZ = al.get_points(0, l)
        assert bits_equal(gX, X) and bits_equal(gY, Y) and bits_equal(gZ, Z), f"points L{l}"
    iters = tuple([6] * L)
    al.run(1, dvo.solver_params(iters=iters))
    poses, info = al.get_poses(1)
    o = O.align_pair(d["ref_gray"], d["ref_depth"], d["now_gray"], L, iters, K=K)
    assert rot_angle(poses[0, :9].reshape(3, 3), o["R"]) < 1e-5 and np.linalg.norm(poses[0, 9:] - o["T"]) < 1e-5
    al.close()


def test_edt_adversarial_patterns():
    """EDT on sparse / structured edge sets (single pixel, one column, corners) checked against the brute-force oracle.
    The edge map is produced by the real Canny stage from crafted images, so this also exercises hysteresis chains."""
    rng = np.random.default_rng(3)
    W, H = 96, 64
    imgs = []
    a = np.full((H, W), 20, np.uint8); a[30:34, 40:44] = 250; imgs.append(a)                  # tiny blob
    b = np.full((H, W), 20, np.uint8); b[:, 47:] = 220; imgs.append(b)                        # one vertical step
    c = np.full((H, W), 20, np.uint8); c[:3, :3] = 255; c[-3:, -3:] = 255; imgs.append(c)     # opposite corners
    d = (rng.integers(0, 2, (H // 8, W // 8)) * 200 + 20).astype(np.uint8).repeat(8, 0).repeat(8, 1); imgs.append(d)
    # a long weak ramp attached to a strong corner: hysteresis must walk the whole chain
    e = np.full((H, W), 20, np.uint8)
    for x in range(W):
        e[40:, x] = 20 + 30 + (x * 10) // W
    e[40:, :6] = 120
    imgs.append(e)
    n = len(imgs)
    gray = np.stack(imgs)
    al = dvo.BatchAligner(W, H, 1, max_batch=n)
    al.set_frames(dvo.FRAME_REF, gray, np.full((n, H, W), 1000, np.uint16))
    al.set_frames(dvo.FRAME_NOW, gray, None)
    al.build_pyramids(n)
    al.prepare(n)
    for i in range(n):
        edge = al.get_level_buffer(i, 1, 0, "edge")
        assert np.array_equal(edge, O.canny(gray[i])), f"image {i}: {mismatch(edge, O.canny(gray[i]))}"
        d2 = al.get_level_buffer(i, 1, 0, "d2")
        assert np.array_equal(d2, O.edt_d2(edge, brute=True)), f"image {i}"
    al.close()


def test_edt_sparse_images_take_the_bisection_kernel():
    """Images with very few edge pixels (nedge * 2048 < P) are routed to the bisection EDT kernel; a lone blob and a
    short segment far from everything give distances of hundreds of pixels."""
    W, H = 320, 240
    a = np.full((H, W), 30, np.uint8); a[100:103, 200:203] = 255
    b = np.full((H, W), 30, np.uint8); b[5:8, 5:12] = 230
    gray = np.stack([a, b])
    al = dvo.BatchAligner(W, H, 1, max_batch=2, intrinsics=(262.5, 262.5, 159.5, 119.5))
    al.set_frames(dvo.FRAME_REF, gray, np.full((2, H, W), 1000, np.uint16))
    al.set_frames(dvo.FRAME_NOW, gray, None)
    al.build_pyramids(2)
    al.prepare(2)
    for i in range(2):
        edge = al.get_level_buffer(i, 1, 0, "edge")
        assert np.array_equal(edge, O.canny(gray[i]))
        n = int((edge > 0).sum())
        assert 0 < n and n * 2048 < W * H, n
        d2 = al.get_level_buffer(i, 1, 0, "d2")
        assert np.array_equal(d2, O.edt_d2(edge)), mismatch(d2, O.edt_d2(edge))
        dtn, _ = O.dt_normalize(d2)
        assert bits_equal(al.get_level_buffer(i, 1, 0, "dtn"), dtn)
    al.close()


def test_edt_packed_rows_redo_far_pixels_exactly():
    """The dense-image EDT rows run on packed 16-bit lanes with column distances clamped to 170 pixels; a row holding a
    pixel further than that from every edge must be redone by the 32-bit routine.  A textured patch in one corner (dense
    enough for the window kernels, nedge * 2048 >= P) leaves most of a 640x480 image 170..600 pixels from the nearest edge;
    a second image puts a single column of edges at x = 171 / 170 / 169 so that distances sit right on the switch."""
    rng = np.random.default_rng(11)
    W, H = 640, 480
    a = np.full((H, W), 40, np.uint8)
    a[:96, :128] = (rng.integers(0, 2, (12, 16)) * 180 + 40).astype(np.uint8).repeat(8, 0).repeat(8, 1)
    b = np.full((H, W), 40, np.uint8); b[:, :171] = 220; b[200:, :170] = 220; b[400:, :169] = 220
    b[:, 600:] = (rng.integers(0, 2, (60, 5)) * 180 + 40).astype(np.uint8).repeat(8, 0).repeat(8, 1)
    gray = np.stack([a, b])
    al = dvo.BatchAligner(W, H, 2, max_batch=2)
    al.set_frames(dvo.FRAME_REF, gray, np.full((2, H, W), 1000, np.uint16))
    al.set_frames(dvo.FRAME_NOW, gray, None)
    al.build_pyramids(2)
    al.prepare(2)
    for i in range(2):
        for l in range(2):
            edge = al.get_level_buffer(i, 1, l, "edge")
            assert np.array_equal(edge, O.canny(O.pyr_nearest(gray[i], l)))
            n = int((edge > 0).sum())
            assert n * 2048 >= edge.size, (i, l, n)                     # dense: the window kernels, not the bisection
            d2 = al.get_level_buffer(i, 1, l, "d2")
            o = O.edt_d2(edge)
            assert np.array_equal(d2, o), mismatch(d2, o)
            if l == 0:
                assert o.max() > 170 * 170 and (o < 170 * 170).any()
            dtn, _ = O.dt_normalize(d2)
            assert bits_equal(al.get_level_buffer(i, 1, l, "dtn"), dtn)
    al.close()


def test_gop_compose_matches_oracle():
    rng = np.random.default_rng(5)
    nseq, nframes = 5, 23
    kind = np.zeros((nseq, nframes), np.int32)
    kind[:, 0] = 1
    kind[:, 5::5] = 2
    rel = np.zeros((nseq, nframes, 12))
    for s in range(nseq):
        for f in range(nframes):
            R, t = O.se3_exp(rng.normal(size=6) * 0.05)
            rel[s, f, :9] = R.reshape(9)
            rel[s, f, 9:] = t
    al = dvo.BatchAligner(64, 64, 1, max_batch=1)
    out = al.gop_compose(kind, rel)
    for s in range(nseq):
        o, is_key, _ = O.gop_replay(kind[s], np.full(nframes, 5, np.int32), rel[s])
        assert np.allclose(out[s], o, rtol=0, atol=1e-12)
    al.close()


def test_undistort_front_end_bit_exact():
    """dvo_undistort (cv::undistort of the publisher, src/camTopic2PublisherPyD.cpp:86-117) against the oracle, which is itself
    pinned against cv2.undistort: BGR u8, gray u8 and 16-bit depth, batched, including an odd size and strong distortion."""
    rng = np.random.default_rng(8)
    for (W, H), K, D in (((640, 480), (525.0, 525.0, 319.5, 239.5), (0.2312, -0.7849, -0.0033, -0.0001, 0.9172)),
                         ((101, 77), (90.0, 88.0, 50.0, 38.0), (-0.35, 0.15, 0.002, -0.001, 0.0))):
        for img in (rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8), rng.integers(0, 256, (2, H, W), dtype=np.uint8),
                    rng.integers(0, 65536, (3, H, W), dtype=np.uint16)):
            got = dvo.undistort(img, K, D)
            for n in range(img.shape[0]):
                assert np.array_equal(got[n], O.undistort(img[n], K, D)), (W, H, img.dtype, img.shape, n)
    # identity lens: an exact copy
    img = rng.integers(0, 256, (1, 48, 64), dtype=np.uint8)
    assert np.array_equal(dvo.undistort(img, (60.0, 60.0, 31.5, 23.5), (0, 0, 0, 0, 0)), img)


def test_raw_frame_ingest_bit_exact():
    """dvo_set_frames_raw (src/camTopic2PublisherPyD.cpp:65-80, :338-348): fused BGR2GRAY + metres -> u16 millimetres
    (cvRound half-even, saturate, 0 -> 1) into the level-0 regions, against the oracle (itself pinned to cv2), then the whole
    path on top of it equals the path fed with the converted images.  Odd size = scalar tail path."""
    rng = np.random.default_rng(21)
    for (W, H, L) in ((640, 480, 4), (101, 77, 2)):
        K = (525.0 * W / 640, 525.0 * W / 640, (W - 1) / 2.0, (H - 1) / 2.0)
        d = O.synth_batch(60, 2, W, H, K, bgr=True, now_depth=True)
        dm_ref = d["ref_depth"].astype(np.float32) / np.float32(1000.0) + rng.uniform(-4e-4, 4e-4, d["ref_depth"].shape).astype(np.float32)
        dm_now = d["now_depth"].astype(np.float32) / np.float32(1000.0)
        # values the conversion must saturate / repair: zero, negative, > 65.535 m, half-way cases, NaN, +-inf
        dm_ref[0, 0, :9] = [0.0, -1.0, 70.0, 0.0005, 0.0015, 0.0025, np.nan, np.inf, -np.inf]
        al = dvo.BatchAligner(W, H, L, max_batch=3, keep_now_depth=True, intrinsics=K)
        al.set_frames_raw(dvo.FRAME_REF, d["ref_bgr"], dm_ref, first=1)
        al.set_frames_raw(dvo.FRAME_NOW, d["now_bgr"], dm_now, first=1)
        al.build_pyramids(2, first=1)
        want_ref = np.stack([O.depth_m_to_mm(dm_ref[i]) for i in range(2)])
        assert list(want_ref[0, 0, :9]) == [1, 1, 65535, 1, 2, 2, 1, 1, 1], want_ref[0, 0, :9]
        for i in range(2):
            assert np.array_equal(al.get_level_buffer(i + 1, 0, 0, "gray"), O.bgr2gray(d["ref_bgr"][i]))
            assert np.array_equal(al.get_level_buffer(i + 1, 1, 0, "gray"), O.bgr2gray(d["now_bgr"][i]))
            assert np.array_equal(al.get_level_buffer(i + 1, 0, 0, "depth"), want_ref[i])
            assert np.array_equal(al.get_level_buffer(i + 1, 1, 0, "depth"), O.depth_m_to_mm(dm_now[i]))
            for l in range(1, L):                 # gray(level) == BGR2GRAY(NEAREST(bgr)) as the publisher computes it
                assert np.array_equal(al.get_level_buffer(i + 1, 0, l, "gray"), O.bgr2gray(O.pyr_nearest(d["ref_bgr"][i], l)))
        iters = tuple([6] * L)
        prm = dvo.solver_params(solver=dvo.GN, iters=iters)
        al.prepare(2, first=1); al.run(2, prm, first=1)
        got, _ = al.get_poses(2, first=1)
        al2 = dvo.BatchAligner(W, H, L, max_batch=3, intrinsics=K)
        g_ref = np.stack([O.bgr2gray(x) for x in d["ref_bgr"]]); g_now = np.stack([O.bgr2gray(x) for x in d["now_bgr"]])
        want, _ = al2.align_batch(g_ref, want_ref, g_now, prm)
        assert np.array_equal(got, want)
        with pytest.raises(dvo.DvoError):          # the reference frame needs depth; nothing may have been modified by the failed call
            al.set_frames_raw(dvo.FRAME_REF, d["ref_bgr"], None, first=1)
        assert np.array_equal(al.get_level_buffer(1, 0, 0, "depth"), want_ref[0])
        al.close(); al2.close()


def test_set_frames_error_leaves_state_untouched():
    """A NOW upload without depth on a keep_now_depth context is rejected BEFORE the previous-frame rotation happens."""
    W, H, K = 160, 120, (131.25, 131.25, 79.5, 59.5)
    d = O.synth_batch(5, 1, W, H, K, now_depth=True)
    al = dvo.BatchAligner(W, H, 2, max_batch=1, keep_now_depth=True, intrinsics=K)
    al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"])
    al.set_frames(dvo.FRAME_NOW, d["now_gray"], d["now_depth"])
    with pytest.raises(dvo.DvoError, match="needs the now frame's depth"):
        al.set_frames(dvo.FRAME_NOW, d["ref_gray"], None)
    with pytest.raises(dvo.DvoError, match="previous now frame"):          # the failed call must not have produced a "previous" frame
        al.promote_now_to_ref(1)
    assert np.array_equal(al.get_level_buffer(0, 1, 0, "gray"), d["now_gray"][0])
    al.close()

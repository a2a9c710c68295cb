"""Probe: solve-stage time per level / solver / arithmetic, and FAST-arithmetic pose deviation from the oracle."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import rgbd_odometry_b200 as dvo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = O.synth_batch(0, B)
al = dvo.BatchAligner(640, 480, 4, max_batch=B)
al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
al.build_pyramids(B); al.prepare(B); al.synchronize()

def timed(params, reps=3):
    al.run(B, params); al.synchronize()
    al.enable_timing(True)
    for _ in range(reps): al.run(B, params)
    ms = al.stage_ms()["solve"] / reps
    al.enable_timing(False)
    return ms

for name, solver in (("gn", dvo.GN), ("subgrad", dvo.SUBGRAD_REF)):
    for arith in (dvo.ARITH_EXACT, dvo.ARITH_FAST):
        for iters in ((10, 10, 10, 10), (10, 0, 0, 0), (0, 10, 0, 0), (0, 0, 10, 0), (0, 0, 0, 10), (0, 0, 0, 50)):
            ms = timed(dvo.solver_params(solver=solver, arithmetic=arith, iters=iters))
            print(f"{name:8s} arith={arith} iters={iters} solve_ms={ms:8.3f}  per-iter-wave_us={ms*1e3/max(1,sum(iters))/max(1.0,B/296):8.2f}")

def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))

n = min(B, 12)
for name, solver, iters in (("subgrad", dvo.SUBGRAD_REF, (50, 50, 50, 50)), ("gn", dvo.GN, (10, 10, 10, 10)), ("lm", dvo.LM, (10, 10, 10, 10))):
    res = {}
    for arith in (dvo.ARITH_EXACT, dvo.ARITH_FAST):
        al.set_initial_pose(B, None); al.run(B, dvo.solver_params(solver=solver, arithmetic=arith, iters=iters))
        res[arith] = al.get_poses(B)[0]
    dR = [rot_angle(res[0][i, :9].reshape(3, 3), res[1][i, :9].reshape(3, 3)) for i in range(B)]
    dT = np.linalg.norm(res[0][:, 9:] - res[1][:, 9:], axis=1)
    print(f"{name}: FAST vs EXACT over {B} pairs: max dR {max(dR):.3e} max dT {dT.max():.3e}; pairs over 1e-5: {int(sum((np.array(dR) > 1e-5) | (dT > 1e-5)))}")
    worst = []
    for i in range(n):
        o = O.align_pair(d["ref_gray"][i], d["ref_depth"][i], d["now_gray"][i], 4, iters, scfg=O.cfg(solver))
        worst.append((rot_angle(res[0][i, :9].reshape(3, 3), o["R"]), np.linalg.norm(res[0][i, 9:] - o["T"])))
    print(f"   EXACT vs oracle over {n} pairs: max dR {max(w[0] for w in worst):.3e} max dT {max(w[1] for w in worst):.3e}")

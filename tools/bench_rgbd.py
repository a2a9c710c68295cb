"""RGBDOdometry path: the reference's eventLoop schedule -- computeJacobianAllLevels once per reference
frame, then gaussNewtonIterations(3, T), gaussNewtonIterations(2, T) -- on a batch of synthetic 640x480 pairs.  Prints
pairs/s (CUDA events, frames resident) and the ingest time (H2D + BGR2GRAY + NEAREST pyramids)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle_lib as O
import rgbd_odometry_b200 as dvo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
levels = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (3, 2)
W, H, L = 640, 480, 4
K = (525.0, 525.0, 319.5, 239.5)
d = O.synth_batch(0, B, W, H, K, bgr=True, now_depth=True)
rg = dvo.RGBDAligner(W, H, L, max_batch=B, intrinsics=K)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); rg.set_stream(stream.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def prep():
    rg.set_frames(dvo.FRAME_REF, d["ref_bgr"], d["ref_depth"]); rg.set_frames(dvo.FRAME_NOW, d["now_bgr"], d["now_depth"])

def solve():
    rg.compute_jacobians(B)
    rg.set_pose(B, None)
    for l in levels:
        rg.gauss_newton(B, l)

prep(); solve(); torch.cuda.synchronize()
t = []
for _ in range(3):
    e0.record(stream); solve(); e1.record(stream); torch.cuda.synchronize(); t.append(e0.elapsed_time(e1))
e0.record(stream); prep(); e1.record(stream); torch.cuda.synchronize(); tprep = e0.elapsed_time(e1)
T, info = rg.get_poses(B)
pix = sum((W >> l) * (H >> l) for l in levels)
sweeps = sum(info[i].iters_run[l] for i in range(B) for l in levels) / B
print(json.dumps({"workload": f"RGBDOdometry semi-dense photometric GN, {B} pairs 640x480, levels {levels} x 3 iterations",
                  "solve_ms": min(t), "pairs_per_s_solve": B / (min(t) * 1e-3), "ingest_ms_incl_h2d": tprep,
                  "mean_selected_points_L2": float(np.mean([info[i].npts[2] for i in range(B)])), "mean_sweeps_per_pair": sweeps,
                  "algorithmic_GBs": B * sweeps / len(levels) * pix * 5 / (min(t) * 1e-3) / 1e9, "launches": rg.launch_count()}))

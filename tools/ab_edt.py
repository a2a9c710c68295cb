"""A/B of the EDT row pass + texel packing at the bench workload: unfused kernels (DVO_EDT_BAND=-1: edt_rows_* + pack_texel_kernel)
against the fused band kernel with 12 / 20 / 28 / 60-row bands.  Prints per-stage CUDA-event times; poses must be bit-identical."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import rgbd_odometry_b200 as dvo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W, H, L = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (640, 480, 4)
K = O.K640 if W == 640 else O.K1280
d = O.synth_batch(0, B, W, H, K)
res, poses = {}, {}
for band in (-1, 12, 160, 20, 200, 24, 28):
    os.environ["DVO_EDT_BAND"] = str(band)
    al = dvo.BatchAligner(W, H, L, max_batch=B, intrinsics=K)
    al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
    prm = dvo.solver_params(solver=dvo.GN, iters=(10,) * L)
    al.build_pyramids(B); al.prepare(B); al.run(B, prm); al.synchronize()
    al.enable_timing(True)
    reps = 3
    for _ in range(reps):
        al.build_pyramids(B); al.prepare(B); al.run(B, prm)
    st = {k: round(v / reps, 3) for k, v in al.stage_ms().items() if v > 0}
    al.enable_timing(False)
    st["total"] = round(sum(st.values()), 3)
    res[f"band{band}"] = st
    poses[band] = al.get_poses(B)[0].copy()
    al.close()
res["identical"] = {str(b): bool(np.array_equal(poses[b], poses[-1])) for b in poses}
print(json.dumps(res, indent=1))

"""One prepared batch, a few solve launches: the command ncu wraps (tools/ one launch per solver variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import rgbd_odometry_b200 as dvo
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
solver = {"gn": dvo.GN, "subgrad": dvo.SUBGRAD_REF}[sys.argv[2] if len(sys.argv) > 2 else "gn"]
it = int(sys.argv[3]) if len(sys.argv) > 3 else 10
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
d = O.synth_batch(0, B)
al = dvo.BatchAligner(640, 480, 4, max_batch=B)
al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
al.build_pyramids(B); al.prepare(B)
prm = dvo.solver_params(solver=solver, iters=(it,) * 4)
for _ in range(reps):
    al.run(B, prm)
al.synchronize()
al.close()

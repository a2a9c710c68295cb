"""One solve launch restricted to the coarsest level (serial-step dominated) -- for ncu source-level sampling."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import rgbd_odometry_b200 as dvo
B = 296
d = O.synth_batch(0, B)
al = dvo.BatchAligner(640, 480, 4, max_batch=B)
al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
al.build_pyramids(B); al.prepare(B); al.synchronize()
solver = dvo.GN if (len(sys.argv) < 2 or sys.argv[1] == "gn") else dvo.SUBGRAD_REF
al.run(B, dvo.solver_params(solver=solver, iters=(0, 0, 0, 50))); al.synchronize()

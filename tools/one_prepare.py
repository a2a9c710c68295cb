"""One resident batch, a few preprocessing passes: the command ncu wraps to profile canny_kernel / edt_pack_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import rgbd_odometry_b200 as dvo
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
d = O.synth_batch(0, B)
al = dvo.BatchAligner(640, 480, 4, max_batch=B)
al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
for _ in range(2):
    al.build_pyramids(B); al.prepare(B)
al.synchronize()
al.close()

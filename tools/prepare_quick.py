"""Quick preprocessing-stage timing at the bench workload (1024 pairs of 640x480): pyramid / canny / edt(+texels) in ms."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import rgbd_odometry_b200 as dvo
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = O.synth_batch(0, B)
al = dvo.BatchAligner(640, 480, 4, max_batch=B)
al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
al.build_pyramids(B); al.prepare(B); al.synchronize()
al.enable_timing(True)
reps = 3
for _ in range(reps):
    al.build_pyramids(B); al.prepare(B)
st = {k: v / reps for k, v in al.stage_ms().items()}
al.enable_timing(False)
print(" ".join(f"{k}={v:.3f}ms" for k, v in st.items() if v > 0))

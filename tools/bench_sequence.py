"""BASELINE config 4: 1280x720, 5-level pyramid, consecutive-pair odometry with key frames every 5 frames, batched over
sequences (dvo_run_sequences).  Prints frame-pairs/s through the host-buffer API."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import rgbd_odometry_b200 as dvo

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
W, H, L = 1280, 720, 5
K = O.K1280
seqs = [O.synth_sequence(500 + s, nframes, W, H, K, max_angle_deg=0.4, max_trans_m=0.008) for s in range(nseq)]
gray = np.stack([s[0] for s in seqs]); depth = np.stack([s[1] for s in seqs])
al = dvo.BatchAligner(W, H, L, max_batch=nseq, keep_now_depth=True, intrinsics=K)
params = dvo.solver_params(iters=(50,) * L)          # the shipped sub-gradient solver, 50 iterations per level
al.run_sequences(gray[:, :3], depth[:, :3], params)   # warm-up
t0 = time.time(); rel, kind, glob = al.run_sequences(gray, depth, params); dt = time.time() - t0
Tw = np.stack([s[3] for s in seqs])
err = np.linalg.norm(glob[:, -1, 9:12] - Tw[:, -1], axis=1)
print(json.dumps({"workload": f"{nseq} sequences x {nframes} frames 1280x720, 5 levels, SUBGRAD_REF 50 it/level, key frame every 5",
                  "seconds": dt, "frame_pairs_per_s": nseq * (nframes - 1) / dt, "keyframes_per_seq": int((kind[0] == 2).sum()) + 1,
                  "final_position_error_m_mean": float(err.mean()), "path_length_m_mean": float(np.linalg.norm(Tw[:, -1], axis=1).mean())}))

"""BASELINE config 4: 1280x720, 5-level pyramid, consecutive-pair odometry with key frames every 5 frames, batched over
sequences (dvo_run_sequences_mem).  Prints frame-pairs/s with pinned host inputs (e2e) and device-resident inputs."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from concurrent.futures import ThreadPoolExecutor
import oracle_lib as O
import rgbd_odometry_b200 as dvo

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
distinct = int(sys.argv[3]) if len(sys.argv) > 3 else nseq
W, H, L = 1280, 720, 5
K = O.K1280
t0 = time.time()
with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
    seqs = list(ex.map(lambda s: O.synth_sequence(500 + s, nframes, W, H, K, max_angle_deg=0.4, max_trans_m=0.008), range(distinct)))
idx = np.arange(nseq) % distinct
gray = torch.from_numpy(np.stack([s[0] for s in seqs])[idx]).pin_memory()
depth = torch.from_numpy(np.stack([s[1] for s in seqs])[idx].view(np.int16)).pin_memory()
tsynth = time.time() - t0
al = dvo.BatchAligner(W, H, L, max_batch=nseq, keep_now_depth=True, intrinsics=K)
params = dvo.solver_params(iters=(50,) * L)          # the shipped sub-gradient solver, 50 iterations per level
pol = dvo.keyframe_policy()
al.run_sequences_mem(gray.data_ptr(), depth.data_ptr(), nseq, nframes, params, pol, want_global=False)   # warm-up
res = {}
for name, dev in (("pinned_host", False), ("device", True)):
    g, d = (gray.cuda(), depth.cuda()) if dev else (gray, depth)
    torch.cuda.synchronize()
    best = None
    for _ in range(2):
        t0 = time.time(); rel, kind, reason, glob = al.run_sequences_mem(g.data_ptr(), d.data_ptr(), nseq, nframes, params, pol, device=dev); dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    res[name] = {"seconds": best, "frame_pairs_per_s": nseq * (nframes - 1) / best}
Tw = np.stack([s[3] for s in seqs])[idx]
err = np.linalg.norm(glob[:, -1, 9:12] - Tw[:, -1], axis=1)
print(json.dumps({"workload": f"{nseq} sequences ({distinct} distinct) x {nframes} frames 1280x720, 5 levels, SUBGRAD_REF 50 it/level, key frame every 5",
                  "synth_seconds": tsynth, **res, "keyframes_per_seq": int((kind[0] == 2).sum()) + 1,
                  "final_position_error_m_mean": float(err.mean()), "path_length_m_mean": float(np.linalg.norm(Tw[:, -1], axis=1).mean())}))

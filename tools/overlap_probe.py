"""Probe: two half batches on two streams, staggered so that one half's solve overlaps the other half's preprocessing."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle_lib as O
import rgbd_odometry_b200 as dvo
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nsplit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = O.synth_batch(0, B)
al = dvo.BatchAligner(640, 480, 4, max_batch=B)
al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
al.synchronize()
prm = dvo.solver_params(solver=dvo.GN, iters=(10, 10, 10, 10))
streams = [torch.cuda.Stream() for _ in range(nsplit)]
evs = [torch.cuda.Event() for _ in range(nsplit)]
main = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
bounds = [(k * B // nsplit, (k + 1) * B // nsplit) for k in range(nsplit)]

def step_serial():
    al.set_stream(main.cuda_stream)
    al.build_pyramids(B); al.prepare(B); al.run(B, prm)

def step_overlap():
    # half k preprocesses after half k-1 finished preprocessing (event), so its preprocessing overlaps half k-1's solve
    for k, (a, b) in enumerate(bounds):
        s = streams[k]
        al.set_stream(s.cuda_stream)
        if k > 0:
            s.wait_event(evs[k - 1])
        al.build_pyramids(b - a, first=a); al.prepare(b - a, first=a)
        evs[k].record(s)
        al.run(b - a, prm, first=a)

def timed(fn, K=6):
    fn(); torch.cuda.synchronize()
    e0.record(main)
    for s in streams: s.wait_event(e0)
    for _ in range(K): fn()
    for s in streams:
        ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

ts = timed(step_serial); to = timed(step_overlap)
al.set_stream(main.cuda_stream)
print(f"B={B} split={nsplit} serial {ts:.3f} ms/step ({B/ts:.1f} k pairs/s)  overlapped {to:.3f} ms/step ({B/to:.1f} k pairs/s)")

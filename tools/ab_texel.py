"""A/B of the solver's texel source at the bench workload: legacy float4 texels (DVO_TEXEL_MODE=0, normgrad_kernel) against
packed 8-byte texels + lookup tables (DVO_TEXEL_MODE=1, pack_texel_kernel), for both solver shapes.  Prints per-stage CUDA-event
times and checks that the poses of the two modes are bit-identical (same per-point values, same summation order)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import rgbd_odometry_b200 as dvo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = O.synth_batch(0, B)
res = {}
poses = {}
for shape in (256, 512):
    for mode in (0, 1):
        os.environ["DVO_TEXEL_MODE"] = str(mode)
        os.environ["DVO_SOLVE_SHAPE"] = str(shape)
        al = dvo.BatchAligner(640, 480, 4, max_batch=B)
        al.set_frames(dvo.FRAME_REF, d["ref_gray"], d["ref_depth"]); al.set_frames(dvo.FRAME_NOW, d["now_gray"], None)
        out = {}
        for name, solver, it in (("gn10", dvo.GN, 10), ("subgrad50", dvo.SUBGRAD_REF, 50)):
            prm = dvo.solver_params(solver=solver, iters=(it,) * 4)
            al.build_pyramids(B); al.prepare(B); al.run(B, prm); al.synchronize()
            al.enable_timing(True)
            reps = 3
            for _ in range(reps):
                al.build_pyramids(B); al.prepare(B); al.run(B, prm)
            st = {k: v / reps for k, v in al.stage_ms().items()}
            al.enable_timing(False)
            out[name] = {k: round(v, 3) for k, v in st.items() if v > 0}
            out[name]["total"] = round(sum(st.values()), 3)
            p, info = al.get_poses(B)
            poses[(shape, mode, name)] = p.copy()
        res[f"shape{shape}_mode{mode}"] = out
        al.close()
for shape in (256, 512):
    for name in ("gn10", "subgrad50"):
        a, b = poses[(shape, 0, name)], poses[(shape, 1, name)]
        res[f"identical_shape{shape}_{name}"] = bool(np.array_equal(a, b))
        if not np.array_equal(a, b):
            res[f"maxdiff_shape{shape}_{name}"] = float(np.abs(a - b).max())
print(json.dumps(res, indent=1))

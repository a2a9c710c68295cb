#!/usr/bin/env python
"""bench.py -- frame-pairs aligned/s (640x480, 4-level pyramid) on N B200s, with roofline and CPU baseline.

Workload (BASELINE.json configs[1]): a batch of synthetic 640x480 RGB-D frame pairs per GPU, NEAREST 4-level
pyramid, Canny + exact EDT + normalise + gradient (now frame), Canny + edge back-projection (reference frame),
Gauss-Newton on H = J^T W J / g = J^T W eps, 10 iterations per level, levels 3 -> 0.  One "step" = one pass of the
whole hot path over the batch.  Inputs are resident in HBM before the timed region (`value`); the `e2e` figure runs
the same batch through dvo_align_batch with pinned HOST buffers (H2D of the images and D2H of the poses inside the
timed region).  Multi-GPU: frame pairs are partitioned across ranks (no data-path collective); NCCL all-gathers the
12-double pose records once per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--solver gn|subgrad|lm]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

W, H, LEVELS = 640, 480, 4
K = (525.0, 525.0, 319.5, 239.5)
PIX = sum((W >> l) * (H >> l) for l in range(LEVELS))        # 408000 px per frame
METRIC = "frame-pairs aligned/s (640x480, 4-lvl)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1024, help="frame pairs per GPU per step")
    ap.add_argument("--solver", default="gn", choices=["gn", "subgrad", "lm"])
    ap.add_argument("--iters", type=int, default=None, help="iterations per level (default 10 for gn/lm, 50 for subgrad)")
    ap.add_argument("--arith", default="exact", choices=["exact", "fast"])
    ap.add_argument("--cpu-sample", type=int, default=48, help="pairs in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def solver_setup(args):
    import oracle_lib as O
    it = args.iters or (50 if args.solver == "subgrad" else 10)
    iters = (it,) * LEVELS
    code = {"gn": 1, "subgrad": 0, "lm": 2}[args.solver]
    return iters, code, O.cfg(solver=code)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, False, []

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def _run_nvml(self):
        """Same fields through NVML (a sample every 5 ms instead of one nvidia-smi process per ~100 ms); False -> fall back."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = (0x8, 0x40, 0x20, 0x4)      # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); get_reasons(h)
        except Exception:
            return False
        while not self.stop_flag:
            try:
                r = int(get_reasons(h))
                self.samples.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            time.sleep(0.005)
        return True

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
            except Exception:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_baseline(args, iters, ocfg, sample, nthreads, seed0=0):
    import oracle_lib as O
    d = O.synth_batch(seed0, sample, W, H, K, nthreads=nthreads)
    best = None
    for _ in range(3):
        _, _, secs = O.align_batch(d["ref_gray"], d["ref_depth"], d["now_gray"], LEVELS, iters, K, ocfg, nthreads=nthreads)
        best = secs if best is None else min(best, secs)
    return sample / best


def run_reference(args):
    """The reference arm: the CPU restatement of the reference's own path (it cannot be compiled here -- SURVEY §8c) on all
    host threads, same metric/config.  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    iters, code, ocfg = solver_setup(args)
    nthreads = os.cpu_count() or 1
    sample = max(nthreads, min(args.cpu_sample, 4 * nthreads))
    d = O.synth_batch(0, sample, W, H, K, nthreads=nthreads)
    times = []
    for s in range(args.warmup + args.steps):
        _, _, secs = O.align_batch(d["ref_gray"], d["ref_depth"], d["now_gray"], LEVELS, iters, K, ocfg, nthreads=nthreads)
        if s >= args.warmup:
            times.append(secs)
    per_step = float(np.mean(times))
    v = sample / per_step
    line = {"metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic", "impl": "reference",
            "config": config_dict(args, iters, sample_note=f"each step = {sample} pairs (bounded sample of the {args.pairs}-pair batch)"),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": nthreads, "kind": "port",
                             "sample": f"{sample} synthetic pairs per step, oracle port of SolveDVO (reference not compilable here)"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(args, iters, sample_note=None):
    c = {"workload": f"batched edge alignment: {args.pairs} synthetic 640x480 frame pairs per GPU, 4-level NEAREST pyramid, "
                     f"{args.solver} solver, {iters[0]} iterations/level (BASELINE configs[1])",
         "pairs_per_gpu": args.pairs, "width": W, "height": H, "levels": LEVELS, "solver": args.solver, "iters_per_level": iters[0],
         "arithmetic": args.arith, "l2": "inputs (1.2 MB/pair) larger than L2, no explicit flush",
         "schedule": "dvo_process: two staggered half batches on internal streams, steps issued back to back"}
    if sample_note:
        c["sample"] = sample_note
    return c


def run_ours(args):
    import torch
    import torch.distributed as dist
    import oracle_lib as O            # synthetic renderer + (rank 0) the cpu_baseline leg only
    import rgbd_odometry_b200 as dvo

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None      # host buffers of the e2e leg next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    iters, code, ocfg = solver_setup(args)
    B = args.pairs
    nthreads = max(1, (os.cpu_count() or 1) // max(1, world))
    # ---- synthetic inputs: this rank's contiguous block of seeds (frame-pair partitioning) ----
    t0 = time.time()
    data = O.synth_batch(rank * B, B, W, H, K, nthreads=nthreads)
    t_synth = time.time() - t0
    al = dvo.BatchAligner(W, H, LEVELS, max_batch=B, device=local, intrinsics=K)
    # a dedicated non-default stream shared by torch (events, NCCL) and the C-ABI context, so that the CUDA events
    # below are recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    al.set_stream(stream.cuda_stream)
    params = dvo.solver_params(solver=code, arithmetic=1 if args.arith == "fast" else 0, iters=iters)
    # inputs resident in HBM (the context's level-0 regions) before any timed region
    al.set_frames(dvo.FRAME_REF, data["ref_gray"], data["ref_depth"])
    al.set_frames(dvo.FRAME_NOW, data["now_gray"], None)
    # dvo_process runs the batch as two staggered half batches on internal streams and copies the poses into a device
    # buffer without joining, so back-to-back steps overlap (solve of one half beside the preprocessing of the other).
    # The pose buffers alternate per step; the NCCL gather of step k runs on its own stream after that step's work and
    # step k+2 (which overwrites the same buffer) first waits for it.
    poses_buf = [torch.empty((B, 12), dtype=torch.float64, device="cuda") for _ in range(2)]
    poses_dev = poses_buf[0]
    gathered = torch.empty((world * B, 12), dtype=torch.float64, device="cuda") if world > 1 else None
    comm = torch.cuda.Stream() if world > 1 else None
    gather_done = [None, None]
    step_no = [0]

    def step():
        k = step_no[0] % 2
        step_no[0] += 1
        if gather_done[k] is not None:
            stream.wait_event(gather_done[k])
        al.process(B, params, poses_out=poses_buf[k].data_ptr())
        if world > 1:
            al.join_stream(comm.cuda_stream)
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(gathered, poses_buf[k])
                ev = torch.cuda.Event()
                ev.record(comm)
            gather_done[k] = ev

    def finish_steps():
        al.join()
        if comm is not None:
            stream.wait_stream(comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = al.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    finish_steps()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = al.launch_count() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * B / (ms_max * 1e-3)

    # ---- per-stage device time (CUDA events on the launching stream), for the roofline of the dominant kernel ----
    poses, info = al.get_poses(B)
    al.enable_timing(True)
    nstage = 3
    for _ in range(nstage):
        al.build_pyramids(B); al.prepare(B); al.run(B, params)
    stage = {k: v / nstage for k, v in al.stage_ms().items()}
    al.enable_timing(False)
    npts = np.array([[inf.npts[l] for l in range(LEVELS)] for inf in info], dtype=np.int64)
    itrun = np.array([[inf.iterations_run[l] for l in range(LEVELS)] for inf in info], dtype=np.int64)
    n_total = int(npts.sum())
    iter_pts = int((npts * itrun).sum())
    # algorithmic bytes: S1 pyramid 4P, S2 canny 4P, S3 EDT 5P, S4 norm+grad 16P, S5 points 3P + 12N, S6 24*sum(I_L N_L)
    alg = {"pyramid": 4 * PIX * B, "canny": (4 + 3) * PIX * B + 12 * n_total, "edt_rows": 5 * PIX * B, "normgrad": 16 * PIX * B,
           "solve": 24 * iter_pts}
    bytes_pair = (32 * PIX * B + 12 * n_total + 24 * iter_pts) / B
    peak, peak_src = peaks()
    # The dominant single kernel is solve_kernel (one launch per step; the other stages are split over 4-8 launches).
    # achieved = algorithmic bytes of that launch (24 B per point-iteration) / its CUDA-event duration.
    dom = "solve"
    dom_gbs = alg[dom] / (stage[dom] * 1e-3) / 1e9
    traffic = None
    try:   # dram__bytes_read+write of solve_kernel per frame pair from the committed `ncu --set full` capture of this workload
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        key = f"solve_kernel_{args.solver}{iters[0]}_bytes_per_pair"
        if key in t:
            traffic = float(t[key]) * B
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "solve_kernel", "achieved": dom_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": dom_gbs / peak, "traffic": traffic,
                "timing": "CUDA events around each stage on the launching stream, 3 staged steps run right after the timed region "
                          "(inside it the two half batches overlap on internal streams, so a kernel cannot be bracketed there)",
                "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": stage[dom],
                "whole_path": {"bytes_per_pair": bytes_pair, "achieved_gbs": bytes_pair * B / (ms_max * 1e-3) / 1e9,
                               "frac": bytes_pair * B / (ms_max * 1e-3) / 1e9 / peak},
                "stages_ms": stage,
                "stages_gbs": {k: alg[k] / (stage[k] * 1e-3) / 1e9 for k in alg if stage[k] > 0},
                "mean_points_per_pair": n_total / B, "mean_point_iterations_per_pair": iter_pts / B}

    # ---- end to end through the public C-ABI with pinned HOST buffers ----
    e2e = None
    if not args.no_e2e:
        pin = {k: torch.from_numpy(data[k]).pin_memory() for k in ("ref_gray", "ref_depth", "now_gray")}
        hp = {k: v.numpy() for k, v in pin.items()}
        al.align_batch(hp["ref_gray"], hp["ref_depth"], hp["now_gray"], params, want_info=False)
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        e0.record(stream)
        for _ in range(n_e2e):
            al.align_batch(hp["ref_gray"], hp["ref_depth"], hp["now_gray"], params, want_info=False)
        e1.record(stream)
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1) / n_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B / (float(ems.item()) * 1e-3), "unit": "pairs/s",
               "h2d_bytes_per_step": int(B * W * H * 4), "d2h_bytes_per_step": int(B * 96), "ms_per_step": float(ems.item())}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the box's host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nth = os.cpu_count() or 1
        sample = max(nth, min(args.cpu_sample, 4 * nth))
        v_all = cpu_baseline(args, iters, ocfg, sample, nth)
        v_one = cpu_baseline(args, iters, ocfg, max(4, sample // max(1, nth)), 1)
        cpu = {"value": v_all, "unit": "pairs/s", "cores": nth, "kind": "port",
               "sample": f"{sample} synthetic pairs, best of 3, oracle port of SolveDVO built -O3 -march=native (reference not compilable here)",
               "single_core_pairs_per_s": v_one}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
                "data": "synthetic", "config": config_dict(args, iters), "clocks": sampler.summary(), "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "numa": numa, "synth_seconds": t_synth,
                "status_nonzero_pairs": int(sum(1 for inf in info if inf.status != 0))}
        print(json.dumps(line))
    al.close()
    if world > 1:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads (and so the first touch of its pinned host buffers) to the NUMA node its GPU hangs off:
    with 8 ranks uploading 1.26 GB per step each, buffers on the far socket halve the aggregate H2D rate.  Best effort:
    returns None when the topology is not visible (containers without /sys NUMA info)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

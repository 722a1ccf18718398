#!/usr/bin/env python
"""bench.py -- RGB-D frames/s fused into the semantic voxel map (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the mapping hot path over one walkthrough of synthetic input:
BASELINE config 2 -- 500 box-room frames, 224x224 depth + 54-class per-pixel
probabilities, fused in order into a 384x384x96 map at 0.05 m.
N > 1 (torchrun, one rank per GPU): every rank fuses its own independent episode, no
data-path collective (weak scaling); value = frames of all ranks / max-over-ranks time.

One JSON line on stdout (rank 0).  `value` has the inputs resident in HBM; `e2e` goes
through the reference-facing layer API with pinned HOST buffers (H2D + D2H inside the
timed region); `roofline` is for the dominant kernel; `cpu_baseline` is the CPU oracle
(a port of the reference's algorithm) timed on this box's host cores on a bounded sample.
`--impl reference` times that CPU implementation alone, with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "RGB-D frames/sec fused into semantic voxel map"
UNIT = "frames/s"

# BASELINE config 2 (agent.py:825-832 defaults + SURVEY.md 8d box-room)
C2 = dict(camera_height=224, camera_width=224, vertical_fov=90.0, map_height=384, map_width=384,
          map_depth=96, feature_size=54, grid_resolution=0.05, interpolation_weight=0.5)
C2_FRAMES = 500
# mean number of distinct voxels touched per frame over the 500 box-room frames, from the oracle's
# indices (DESIGN.md "algorithmic bytes"; regenerate with tools/count_touched.py)
C2_TOUCHED_PER_FRAME = 22243.0


# dram__bytes_read.sum + dram__bytes_write.sum of the k_cell_accumulate launches of one step on this workload (one
# working launch + the empty overflow rounds), from the committed ncu launch list named below
ACCUMULATE_DRAM_BYTES_PER_LAUNCH = 7622.8e6 + 263.4e6
ACCUMULATE_TRAFFIC_SOURCE = "profiles/r01zz_step_traffic.txt"
# the same two metrics summed over all 36 launches of one step (same capture)
STEP_DRAM_BYTES = 9881.0e6 + 1576.6e6


def algorithmic_bytes_per_frame(h, w, fh, fw, F, touched):
    """depth + features as fed + one read and one write of every touched
    voxel row + pose."""
    return 4.0 * h * w + 4.0 * fh * fw * F + 8.0 * F * touched + 48.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms while the timed region (and the load phase before it) runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_walkthrough(num_frames, seed_offset=0, H=224, W=224, F=54):
    """Host-side synthetic walkthrough: poses, depth [T,H,W,1] and LOW-RES class probabilities
    [T,H/8,W/8,F] (nearest up-sampling x8 gives the per-pixel probabilities)."""
    from mass_b200.utils import synthetic
    rays = synthetic.camera_rays(H, W)
    pos, yaw, elev, depth, low = [], [], [], [], []
    for t in range(num_frames):
        p, y, e = synthetic.boxroom_pose(t, num_frames)
        d, _ = synthetic.render_depth(rays, p, y, e)
        pos.append(p), yaw.append(y), elev.append(e), depth.append(d[..., None])
        low.append(synthetic.boxroom_probs(t + seed_offset, H, W, F))
    return dict(position=np.stack(pos), yaw=np.array(yaw, np.float32), elevation=np.array(elev, np.float32),
                depth=np.stack(depth), probs_low=np.stack(low))


def cpu_port_frames_per_s(walk, frame_ids, nthreads):
    """Times the CPU oracle (port of the reference's algorithm) on the given frames."""
    from mass_b200.utils import synthetic
    from oracle import oracle
    layer = oracle.OracleLayer(nthreads=nthreads, **C2, **synthetic.MAP_ORIGIN)
    frames = [dict(position=walk["position"][t], yaw=walk["yaw"][t], elevation=walk["elevation"][t],
                   depth=walk["depth"][t], features=synthetic.upsample(walk["probs_low"][t], 8)) for t in frame_ids]
    layer.update(frames[0])                      # warm-up (page faults of the 2.85 GiB map rows)
    touched = []
    t0 = time.perf_counter()
    for f in frames:
        layer.update(f)
        touched.append(layer.n_touched)
    dt = time.perf_counter() - t0
    return len(frames) / dt, float(np.mean(touched))


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (CPU oracle, all threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 12
    walk = make_walkthrough(C2_FRAMES)
    ids = [int(i) for i in np.linspace(0, C2_FRAMES - 1, sample)]
    for _ in range(max(args.warmup, 0)):
        cpu_port_frames_per_s(walk, ids[:2], cores)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_port_frames_per_s(walk, ids, cores)
        vals.append(v)
    elapsed = time.perf_counter() - t0
    value = float(np.mean(vals))
    desc = "%d of the %d frames (evenly spaced) per step" % (sample, C2_FRAMES)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(n_gpus):
    return {"workload": "c2: walkthrough of 500 synthetic box-room 224x224 RGB-D frames + 54-class per-pixel "
                        "probabilities into a 384x384x96 map at 0.05 m, frames fused in order",
            "frames_per_step": C2_FRAMES, "episodes": n_gpus, "sharding": "one independent episode per GPU, "
            "no data-path collective", "l2": "inputs (5.5 GB per step) are larger than L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"],
                    help="voxel-reduce arithmetic: affine form (<=1e-5 rel) or the reference's operation order")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels from the host instead of a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from mass_b200 import _lib
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    T, H, W, F = C2_FRAMES, 224, 224, 54
    walk = make_walkthrough(T, seed_offset=10000 * rank)          # each rank: its own episode
    layer = BaseProjectionLayer(exact=(args.mode == "exact"), **C2, **synthetic.MAP_ORIGIN).to(dev)
    L = _lib.lib()

    # ---- device-resident inputs (value) ---------------------------------------------------------
    depth_d = torch.from_numpy(walk["depth"]).to(dev)
    low_d = torch.from_numpy(walk["probs_low"]).to(dev)
    probs_d = low_d.repeat_interleave(8, dim=1).repeat_interleave(8, dim=2).contiguous()   # [T,H,W,F]
    del low_d
    obs_d = dict(position=walk["position"], yaw=walk["yaw"], elevation=walk["elevation"], depth=depth_d,
                 features=probs_d)

    prep = layer.prepare_batch(obs_d)                 # poses on the device too: the timed region has no host work
    graph = None

    def step_device():
        if graph is not None:
            graph.replay()
        else:
            layer.update_prepared(prep)

    for _ in range(args.warmup):
        step_device()
    # one extra untimed step with stage events: duration of each stage of the batched pipeline
    import ctypes
    L.mb_profile_stages(1)
    step_device()
    stage_ms = (ctypes.c_float * 8)()
    n_stage = L.mb_profile_read(stage_ms, 8)
    L.mb_profile_stages(0)
    stages = dict(zip(["voxelise", "sort", "index", "scalar_pass", "accumulate", "apply"],
                      [float(stage_ms[i]) for i in range(n_stage)]))
    # the step as a CUDA graph: same kernels, launched back to back without host latency between them
    launches_per_step = None
    if not args.no_graph:
        torch.cuda.synchronize()
        before = L.mb_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            layer.update_prepared(prep)
        launches_per_step = int(L.mb_launch_count() - before)
        graph = g
        for _ in range(2):
            step_device()
    barrier()
    launches0 = L.mb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        # ~0.5 s of untimed steps first, so that nvidia-smi (50 ms period) samples the clocks under this very load
        # right up to and during the timed region, which itself lasts only a few milliseconds
        t_load = time.perf_counter() + 0.5
        while time.perf_counter() < t_load:
            step_device()
            torch.cuda.synchronize()
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step_device()
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = int(L.mb_launch_count() - launches0)
    if launches_per_step is not None:
        launches = launches_per_step * args.steps      # replayed from the graph: counted at capture
    value = world * T * args.steps / (ms_total * 1e-3)

    # ---- end to end: pinned host buffers -> layer API -> D2H of a result ------------------------------
    e2e = None
    if not args.no_e2e:
        depth_h = torch.from_numpy(walk["depth"]).pin_memory()
        probs_h = torch.empty(probs_d.shape, dtype=torch.float32, pin_memory=True)
        probs_h.copy_(probs_d)
        chunk = 100
        # two device staging buffers: the copy of chunk i+1 (copy stream) overlaps the fusion of chunk i
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [(torch.empty((chunk,) + tuple(depth_h.shape[1:]), dtype=torch.float32, device=dev),
                  torch.empty((chunk,) + tuple(probs_h.shape[1:]), dtype=torch.float32, device=dev)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def step_host():
            main = torch.cuda.current_stream(dev)
            for ev in free:
                ev.record(main)
            for i, s in enumerate(range(0, T, chunk)):
                e = min(s + chunk, T)
                d_buf, p_buf = stage[i % 2]
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[i % 2])
                    d_buf[:e - s].copy_(depth_h[s:e], non_blocking=True)
                    p_buf[:e - s].copy_(probs_h[s:e], non_blocking=True)
                    ready[i % 2].record(copy_stream)
                main.wait_event(ready[i % 2])
                layer.update_batch(dict(position=walk["position"][s:e], yaw=walk["yaw"][s:e],
                                        elevation=walk["elevation"][s:e], depth=d_buf[:e - s], features=p_buf[:e - s]))
                free[i % 2].record(main)
            occupied = (layer.data[:, :, :, 0] != 0).sum()
            return int(occupied.item())                                   # D2H read of the result

        step_host()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(1, min(args.steps, 2))):
            step_host()
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        n_e2e = max(1, min(args.steps, 2))
        e2e = {"value": world * T * n_e2e / (float(ems.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(depth_h.numel() * 4 + probs_h.numel() * 4 + T * 48),
               "d2h_bytes_per_step": 8}
        del depth_h, probs_h, stage

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the step's kernels (one GPU's share) -----------------------------------------------
    peak, peak_src = measured_peaks()
    bpf = algorithmic_bytes_per_frame(H, W, H, W, F, C2_TOUCHED_PER_FRAME)
    per_gpu_fps = T * args.steps / (ms_total * 1e-3)
    achieved = per_gpu_fps * bpf / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": STEP_DRAM_BYTES, "traffic_source": ACCUMULATE_TRAFFIC_SOURCE, "peak_source": peak_src,
                "kernel": "whole step (all kernels of update_batch)",
                "algorithmic_bytes_per_frame": bpf, "stage_ms": stages}
    if "accumulate" in stages and stages["accumulate"] > 0:
        # the dominant kernel alone: it streams every feature row once (4*h*w*F per frame) and writes the run rows
        acc_bytes = 4.0 * H * W * F * T
        roofline["dominant_kernel"] = {
            "kernel": "k_cell_accumulate", "ms": stages["accumulate"], "algorithmic_bytes": acc_bytes,
            "achieved": acc_bytes / (stages["accumulate"] * 1e-3) / 1e9, "unit": "GB/s",
            "frac": acc_bytes / (stages["accumulate"] * 1e-3) / 1e9 / peak,
            "traffic": ACCUMULATE_DRAM_BYTES_PER_LAUNCH, "traffic_source": ACCUMULATE_TRAFFIC_SOURCE}

    cpu = None
    if not args.no_cpu_baseline and world == 1 or (not args.no_cpu_baseline and rank == 0 and world == 1):
        cores = os.cpu_count() or 1
        ids = [int(i) for i in np.linspace(0, T - 1, 24)]
        v, touched = cpu_port_frames_per_s(walk, ids, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "24 of the 500 frames (evenly spaced), CPU oracle with %d threads" % cores,
               "touched_voxels_per_frame_in_sample": touched}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(workload_config(world), mode=args.mode, launch="host" if args.no_graph else "cuda graph"), "clocks": clocks.summary(),
           "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

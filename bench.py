#!/usr/bin/env python
"""bench.py -- RGB-D frames/s fused into the semantic voxel map (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4]

Default (`--config c2`, the configuration the metric is quoted on): a step is one pass of the mapping hot
path over one walkthrough of synthetic input -- BASELINE config 2: 500 box-room frames, 224x224 depth +
54-class per-pixel probabilities, fused in order into a 384x384x96 map at 0.05 m.
N > 1 (torchrun, one rank per GPU): every rank fuses its own independent episode, no data-path collective
(weak scaling); value = frames of all ranks / max-over-ranks time.

One JSON line on stdout (rank 0).  `value` has the inputs resident in HBM; `e2e` goes through the
reference-facing layer API (`layer.update_batch`) with pinned HOST buffers (H2D + D2H inside the timed region);
`roofline` is the step against SURVEY.md 8d's bytes per frame (U_f counted by the kernels in this very run), with
the dominant kernel and the input-only bound beside it; `class_id_path` is the same walkthrough through
`SemanticProjectionLayer.update_batch` (arg-max class ids: the path agent.py drives); `cpu_baseline` is the
reference's own torch CPU path (oracle/_ref: the UNMODIFIED reference, mass/nn/base_projection_layer.py:282-343)
timed on this box's host cores on a bounded sample, with the C port's rate beside it.  `--impl reference` times
that reference CPU path alone, with all host threads.

`--config c3` (bench_configs.py): BASELINE config 3, the episode pair incl. the MATCH stage -- two semantic + two
256-d maps built, then the agent's predict_scene_differences loop; `--config c4`: BASELINE config 4, the frames of
one large scene sharded over the ranks with the ordered affine combine over NVLink (strong scaling).  With N > 1
the default line also carries a short `c4` object so that the sharded path shows up in the scaling record.
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

# NOT measured by this script: dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu pass named
# below (same command, same workload, the build named there).  bench.py only repeats them so that the JSON line
# carries the traffic next to the algorithmic bytes; every other figure of the line is measured live.
TRAFFIC_SOURCE = ("profiles/r03z_step_traffic.txt (ncu pass of the round-2 build just before k_seg_sums_long was added: "
                  "22 of the final 23 launches, the 23rd moves no data on this workload; NOT measured in this run)")
ACCUMULATE_DRAM_BYTES_PER_LAUNCH = 6842.1e6 + 171.4e6     # the k_cell_accumulate launch of one step
STEP_DRAM_BYTES = 9042.6e6 + 1459.6e6                       # all 22 launches of one step of that build


def algorithmic_bytes_per_frame(h, w, fh, fw, F, touched_row_floats, feature_bytes=4.0):
    """depth + features as fed + one read and one write of every voxel row the frame touches
    (touched_row_floats = map channels x U_f) + pose.  feature_bytes = bytes per fed feature element (4 for float
    probabilities; for class ids pass F = 1 and 8: one int64 per pixel)."""
    return 4.0 * h * w + feature_bytes * fh * fw * F + 8.0 * touched_row_floats + 48.0


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


def _frame(walk, t):
    from mass_b200.utils import synthetic
    return dict(position=walk["position"][t], yaw=walk["yaw"][t], elevation=walk["elevation"][t],
                depth=walk["depth"][t], features=synthetic.upsample(walk["probs_low"][t], 8))


def cpu_port_frames_per_s(walk, frame_ids, nthreads):
    """Times the CPU oracle (C port of the reference's algorithm) on the given frames."""
    from mass_b200.utils import synthetic
    from oracle import oracle
    layer = oracle.OracleLayer(nthreads=nthreads, **C2, **synthetic.MAP_ORIGIN)
    frames = [_frame(walk, t) for t in frame_ids]
    layer.update(frames[0])                      # warm-up (page faults of the 2.85 GiB map rows)
    t0 = time.perf_counter()
    for f in frames:
        layer.update(f)
    dt = time.perf_counter() - t0
    return len(frames) / dt


_REF_LAYER = {}


def cpu_reference_frames_per_s(walk, frame_ids, nthreads, warm=1):
    """Times the UNMODIFIED reference (oracle/_ref or /root/reference, through oracle/reference.py):
    BaseProjectionLayer.update on CPU tensors, mass/nn/base_projection_layer.py:282-343, nthreads torch threads.
    Returns (frames/s, seconds per frame list)."""
    from mass_b200.utils import synthetic
    from oracle import reference
    R = reference.load()
    torch.set_num_threads(nthreads)
    if "layer" not in _REF_LAYER:
        _REF_LAYER["layer"] = R.base.BaseProjectionLayer(**C2, **synthetic.MAP_ORIGIN)
    layer = _REF_LAYER["layer"]
    frames = [_frame(walk, t) for t in frame_ids]
    for f in frames[:warm]:
        layer.update(f)
    per = []
    for f in frames:
        t0 = time.perf_counter()
        layer.update(f)
        per.append(time.perf_counter() - t0)
    return len(frames) / sum(per), per


def workload_config(n_gpus):
    return {"workload": "c2: walkthrough of 500 synthetic box-room 224x224 RGB-D frames + 54-class per-pixel "
                        "probabilities into a 384x384x96 map at 0.05 m, frames fused in order",
            "frames_per_step": C2_FRAMES, "episodes": "one per GPU", "sharding": "one independent episode per GPU, "
            "no data-path collective", "l2": "inputs (5.5 GB per step) are larger than L2; no flush needed"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (all threads),
    each step a bounded sample of the C2 walkthrough."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference
    cores = os.cpu_count() or 1
    walk = make_walkthrough(C2_FRAMES)
    steps = max(args.steps, 1)
    if reference.available():
        kind = "reference"
        # warm-up doubles as the probe that sizes the sample: >= 24 frames per step unless K steps of that
        # would run past ~4 minutes
        _, per = cpu_reference_frames_per_s(walk, [0, 250], cores, warm=1)
        for _ in range(max(args.warmup - 1, 0)):
            cpu_reference_frames_per_s(walk, [125], cores, warm=0)
        sample = int(max(6, min(24, 240.0 / (steps * max(per)))))

        def one_step(ids):
            return cpu_reference_frames_per_s(walk, ids, cores, warm=0)[0]
        what = "the UNMODIFIED reference (mass.nn.base_projection_layer.BaseProjectionLayer.update, torch CPU, %d threads)" % cores
    else:
        kind = "port"
        sample = 12
        for _ in range(max(args.warmup, 0)):
            cpu_port_frames_per_s(walk, [0, 250], cores)

        def one_step(ids):
            return cpu_port_frames_per_s(walk, ids, cores)
        what = "reference not installed in oracle/_ref: C port of its algorithm (oracle/mass_oracle.c, %d threads)" % cores
    ids = [int(i) for i in np.linspace(0, C2_FRAMES - 1, sample)]
    vals = []
    t0 = time.perf_counter()
    for _ in range(steps):
        vals.append(one_step(ids))
    elapsed = time.perf_counter() - t0
    value = float(np.mean(vals))
    port = cpu_port_frames_per_s(walk, ids, cores)
    desc = "%d of the %d frames (evenly spaced) per step; %s" % (sample, C2_FRAMES, what)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
                         "port_value": port, "port": "oracle/mass_oracle.c on the same frames, %d threads" % cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
class Ranks:
    """torch.distributed plumbing of one bench process (one rank per GPU)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def timed_steps(R, step, steps):
    """EXACTLY `steps` calls of step() between barrier + synchronize, CUDA events on the launching stream,
    max over ranks.  Returns total ms."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R.barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    R.barrier()
    return R.max_ms(ev0.elapsed_time(ev1))


def graph_of(L, fn):
    """Captures fn() (kernel launches only) in a CUDA graph; returns (graph, launches inside)."""
    torch.cuda.synchronize()
    before = L.mb_launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g, int(L.mb_launch_count() - before)


def h2d_ceiling(R, nbytes=256 << 20, reps=8):
    """Bare host->device copy rate with every rank copying at once (one cudaMemcpyAsync per repetition from pinned
    memory, nothing else running): the box's host-side ceiling for the e2e figure.  Returns (this rank's GB/s as
    the max over ranks of the elapsed time, aggregate GB/s)."""
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=R.dev)
    dst.copy_(src, non_blocking=True)
    ms = timed_steps(R, lambda: dst.copy_(src, non_blocking=True), reps)
    per_rank = nbytes * reps / (ms * 1e-3) / 1e9
    return per_rank, per_rank * R.world


def run_c2(args):
    from mass_b200 import _lib
    from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import synthetic
    import ctypes

    R = Ranks()
    dev, world, rank = R.dev, R.world, R.rank
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
    # one extra untimed step with stage events: duration of each stage of the batched pipeline; its device counters
    # give the touched-voxel figures the roofline bookkeeping needs (nothing is hard-coded)
    L.mb_profile_stages(1)
    step_device()
    stage_ms = (ctypes.c_float * 8)()
    n_stage = L.mb_profile_read(stage_ms, 8)
    L.mb_profile_stages(0)
    stages = dict(zip(["voxelise", "sort", "index", "scalar_pass", "accumulate", "apply"],
                      [float(stage_ms[i]) for i in range(n_stage)]))
    counters = layer.counters() if args.mode == "fast" else None
    launches_per_step = None
    if not args.no_graph:
        graph, launches_per_step = graph_of(L, lambda: layer.update_prepared(prep))
        for _ in range(2):
            step_device()
    R.barrier()
    launches0 = L.mb_launch_count()
    with ClockSampler(R.local) as clocks:
        # ~0.5 s of untimed steps first, so that nvidia-smi (50 ms period) samples the clocks under this very load
        # right up to and during the timed region, which itself lasts only a few milliseconds
        t_load = time.perf_counter() + 0.5
        while time.perf_counter() < t_load:
            step_device()
            torch.cuda.synchronize()
        ms_total = timed_steps(R, step_device, args.steps)
    launches = int(L.mb_launch_count() - launches0)
    if launches_per_step is not None:
        launches = launches_per_step * args.steps      # replayed from the graph: counted at capture
    value = world * T * args.steps / (ms_total * 1e-3)

    # ---- end to end: pinned host buffers -> layer.update_batch (the public call) -> D2H of a result ------------
    e2e = None
    n_e2e = max(1, min(args.steps, 2))
    if not args.no_e2e:
        depth_h = torch.from_numpy(walk["depth"]).pin_memory()
        probs_h = torch.empty(probs_d.shape, dtype=torch.float32, pin_memory=True)
        probs_h.copy_(probs_d)
        obs_h = dict(position=walk["position"], yaw=walk["yaw"], elevation=walk["elevation"], depth=depth_h,
                     features=probs_h)

        def step_host():
            # update_batch cuts the host frames into chunks and overlaps their copies with the fusion itself
            layer.update_batch(obs_h)
            occupied = (layer.data[:, :, :, 0] != 0).sum()
            return int(occupied.item())                                   # D2H read of the result

        step_host()
        ems = timed_steps(R, step_host, n_e2e)
        ceil_rank, ceil_all = h2d_ceiling(R)
        h2d = int(depth_h.numel() * 4 + probs_h.numel() * 4 + T * 48)
        e2e = {"value": world * T * n_e2e / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8, "api": "BaseProjectionLayer.update_batch(host tensors): chunked, copies "
               "overlapped with the fusion of the previous chunk inside the call",
               "h2d_gbs_achieved": world * h2d * n_e2e / (ems * 1e-3) / 1e9,
               "h2d_ceiling_gbs": ceil_all, "h2d_ceiling": "bare concurrent cudaMemcpyAsync from pinned memory, "
               "256 MiB x 8 per rank, all %d ranks at once (measured in this run)" % world,
               "frac_of_h2d_ceiling": (world * h2d * n_e2e / (ems * 1e-3) / 1e9) / ceil_all}
        del depth_h, probs_h, obs_h

    # ---- the class-id path: SemanticProjectionLayer.update_batch(semantic ids), the call agent.py makes ----------
    class_id = None
    if not args.no_class_ids and args.mode == "fast":
        ids_d = probs_d.argmax(-1).unsqueeze(-1).contiguous()                       # [T,H,W,1] int64
        del probs_d, obs_d, prep, graph
        graph = None
        sem = SemanticProjectionLayer(exact=False, **C2, **synthetic.MAP_ORIGIN).to(dev)
        sprep = sem.prepare_batch(dict(position=walk["position"], yaw=walk["yaw"], elevation=walk["elevation"],
                                       depth=depth_d, class_ids=ids_d[..., 0]))
        for _ in range(max(args.warmup, 1)):
            sem.update_prepared(sprep)
        sc = sem.counters()
        sgraph, _ = graph_of(L, lambda: sem.update_prepared(sprep))
        sgraph.replay()
        sms = timed_steps(R, sgraph.replay, args.steps)
        class_id = {"value": world * T * args.steps / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms / args.steps,
                    "api": "SemanticProjectionLayer (class ids, no one-hot tensor materialised; "
                           "mass/nn/applications/semantic_projection_layer.py:203-214)"}
        if sc:
            b = algorithmic_bytes_per_frame(H, W, H, W, 1, F * sc["voxel_frames"] / T, feature_bytes=8.0)
            class_id["algorithmic_bytes_per_frame"] = b
            class_id["roofline_frac"] = (T * args.steps / (sms * 1e-3)) * b / 1e9 / measured_peaks()[0]
        if not args.no_e2e:
            ids_h = ids_d.cpu().pin_memory()
            depth_h = torch.from_numpy(walk["depth"]).pin_memory()
            sobs = dict(position=walk["position"], yaw=walk["yaw"], elevation=walk["elevation"], depth=depth_h,
                        semantic=ids_h)

            def sem_host():
                sem.update_batch(sobs)
                return int((sem.data[:, :, :, 0] != 0).sum().item())

            sem_host()
            sems = timed_steps(R, sem_host, n_e2e)
            class_id["e2e"] = {"value": world * T * n_e2e / (sems * 1e-3), "unit": UNIT,
                               "h2d_bytes_per_step": int(depth_h.numel() * 4 + ids_h.numel() * 8 + T * 48),
                               "d2h_bytes_per_step": 8}
        del sem, sgraph, sprep, ids_d

    # ---- the frame-sharded scene (BASELINE config 4) in short, at every N, so that the scaling record also holds ------
    # ---- a path WITH a data-path exchange (strong scaling), not only replicas ---------------------------------------
    c4 = None
    if not args.no_c4:
        try:
            import bench_configs
            layer = prep = graph = None
            torch.cuda.empty_cache()
            # the same 1024-frame scene at every N (strong scaling); `--config c4` runs all 4096 frames
            c4 = bench_configs.c4_sharded(R, frames_total=1024 if args.c4_frames <= 0 else args.c4_frames, brief=True)
        except Exception as exc:                                  # never lose the main line to the extra one
            c4 = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank != 0:
        R.close()
        return

    # ---- roofline of the step's kernels (one GPU's share) -----------------------------------------------
    peak, peak_src = measured_peaks()
    per_gpu_fps = T * args.steps / (ms_total * 1e-3)
    roofline = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                "kernel": "whole step (all kernels of update_batch)", "stage_ms": stages,
                "traffic": STEP_DRAM_BYTES, "traffic_source": TRAFFIC_SOURCE}
    if counters:
        # U_f = voxels one frame touches, averaged over the 500 frames: counted by k_voxel_scalars in this run
        u_f = counters["voxel_frames"] / T
        bpf = algorithmic_bytes_per_frame(H, W, H, W, F, F * u_f)
        achieved = per_gpu_fps * bpf / 1e9
        # what a T-frame batch must move at the very least: every input byte once + every distinct touched row once
        must = (4.0 * H * W * (1 + F) + 48.0) * T + 8.0 * F * counters["voxels"]
        roofline.update({"achieved": achieved, "frac": achieved / peak, "frac_of_nominal_8tbs": achieved / 8000.0,
                         "algorithmic_bytes_per_frame": bpf, "touched_voxels_per_frame": u_f,
                         "touched_voxels_per_step": counters["voxels"], "counters": counters,
                         "input_only_bytes_per_step": must,
                         "input_only_frac": must / (ms_total / args.steps * 1e-3) / 1e9 / peak})
    if "accumulate" in stages and stages["accumulate"] > 0:
        # the dominant kernel alone: it streams every feature row once (4*h*w*F per frame) and writes the run rows
        acc_bytes = 4.0 * H * W * F * T
        roofline["dominant_kernel"] = {
            "kernel": "k_cell_accumulate", "ms": stages["accumulate"], "algorithmic_bytes": acc_bytes,
            "achieved": acc_bytes / (stages["accumulate"] * 1e-3) / 1e9, "unit": "GB/s",
            "frac": acc_bytes / (stages["accumulate"] * 1e-3) / 1e9 / peak,
            "traffic": ACCUMULATE_DRAM_BYTES_PER_LAUNCH, "traffic_source": TRAFFIC_SOURCE}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import reference
        cores = os.cpu_count() or 1
        ids = [int(i) for i in np.linspace(0, T - 1, 24)]
        port = cpu_port_frames_per_s(walk, ids, cores)
        if reference.available():
            v, _ = cpu_reference_frames_per_s(walk, ids, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "24 of the 500 frames (evenly spaced): the UNMODIFIED reference's "
                             "BaseProjectionLayer.update on CPU tensors, torch with %d threads" % cores,
                   "port_value": port, "port": "oracle/mass_oracle.c on the same 24 frames, %d threads" % cores}
        else:
            cpu = {"value": port, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "24 of the 500 frames (evenly spaced), C port of the reference's algorithm "
                             "(oracle/_ref not installed), %d threads" % cores}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(world),
           "run": {"mode": args.mode, "launch": "host" if args.no_graph else "cuda graph"}, "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
           "cpu_baseline": cpu, "class_id_path": class_id}
    if c4 is not None:
        out["c4"] = c4
    print(json.dumps(out))
    R.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4"])
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"],
                    help="voxel-reduce arithmetic: affine form (<=1e-5 rel) or the reference's operation order")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-class-ids", action="store_true")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--c4-frames", type=int, default=0, help="frames of the c4 scene (default 4096; brief run: 256 per GPU)")
    ap.add_argument("--check", action="store_true", help="c3 / c4: also compare a frame subset with the CPU oracle")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels from the host instead of a CUDA graph")
    args = ap.parse_args()
    if args.config != "c2":
        import bench_configs
        return bench_configs.main(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_c2(args)


if __name__ == "__main__":
    main()
